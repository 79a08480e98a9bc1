#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cmath>
struct float2{float x,y;};
#define M 1e-4f
struct HullLine { float l0, l1, l2; };
static HullLine hull_line(float2 a, float2 b) { return {a.y * b.x - a.x * b.y, b.y - a.y, a.x - b.x}; }
static float hull_side(const HullLine& l, float2 p) { return (l.l0 + p.x * l.l1) + p.y * l.l2; }
std::vector<float2> refchain(const std::vector<float2>& pts){
  std::vector<float2> st;
  for(auto p: pts){
    while(st.size()>=2){ HullLine l=hull_line(st[st.size()-2],st[st.size()-1]); if(!(hull_side(l,p)<=M))break; st.pop_back(); }
    st.push_back(p);
  }
  return st;
}
long g_iter=0,g_pts=0;
std::vector<float2> look(const std::vector<float2>& pts, uint32_t WIN){
  uint32_t n=pts.size();
  if(n<3) return pts;
  std::vector<float2> stack(n+4);
  auto lds=[&](uint32_t addr){return stack[addr/8];};
  auto sts=[&](uint32_t addr,float2 v){stack[addr/8]=v;};
  float2 a=pts[0], b=pts[1], c=a; sts(0,a); sts(8,b);
  HullLine lab=hull_line(a,b), lca=lab;
  uint32_t top=16; const uint32_t floor2=16;
  uint32_t nw=(n+WIN-1)/WIN;
  for(uint32_t w=0;w<nw;w++){
    uint32_t k0= w==0?2:0, k1=std::min(WIN, n-w*WIN);
    for(uint32_t k=k0;k<k1;){
      g_iter++;
      uint32_t rem=k1-k; uint32_t base=w*WIN+k;
      float2 p0=pts[base], p1=pts[base+(rem>1?1:0)], p2=pts[base+(rem>2?2:0)], p3=pts[base+(rem>3?3:0)];
      HullLine l1=hull_line(a,p0), l2=hull_line(a,p1), l3=hull_line(a,p2), l4=hull_line(a,p3);
      float u0=hull_side(lab,p0), u1=hull_side(l1,p1), u2=hull_side(l2,p2), u3=hull_side(l3,p3);
      bool two= top==floor2;
      bool r0 = u0<=M && (two || !(hull_side(lca,p0)<=M));
      bool r1 = r0 && rem>1 && u1<=M && (two || !(hull_side(lca,p1)<=M));
      bool r2 = r1 && rem>2 && u2<=M && (two || !(hull_side(lca,p2)<=M));
      bool r3 = r2 && rem>3 && u3<=M && (two || !(hull_side(lca,p3)<=M));
      if(r0){ uint32_t run=1+r1+r2+r3; b= r3?p3:(r2?p2:(r1?p1:p0)); lab= r3?l4:(r2?l3:(r1?l2:l1)); sts(top-8,b); k+=run; g_pts+=run; continue; }
      float2 p=p0;
      if(!(u0<=M)){ sts(top,p); top+=8; c=a;a=b;b=p; lca=lab; lab=hull_line(a,b); }
      else { top-=16; b=c;
        while(top>=floor2){ a=lds(top-16); lab=hull_line(a,b); if(!(hull_side(lab,p)<=M))break; top-=8; b=a; }
        sts(top,p); top+=8; a=b; b=p; lab=hull_line(a,b);
        if(top>floor2){ c=lds(top-24); lca=hull_line(c,a);} }
      k+=1; g_pts++;
    }
  }
  stack.resize(top/8); return stack;
}
int main(){
  std::mt19937 rng(7);
  for(int trial=0;trial<60000;++trial){
    int n=1+rng()%300; std::vector<float2> pts(n);
    int mode=rng()%5;
    for(auto&p:pts){ if(mode==0){p.x=(rng()%1000)/100.f;p.y=(rng()%1000)/100.f;} else if(mode==1){p.x=(rng()%8);p.y=(rng()%8);} else if(mode==2){p.x=(rng()%100)/10.f; p.y=p.x*0.5f+(rng()%3)*1e-4f;} else if(mode==3){p.x=(rng()%20)*0.01f;p.y=(rng()%20)*0.01f;} else { float t=(rng()%10000)/10000.f*6.2831853f; p.x=100*cosf(t); p.y=100*sinf(t);} }
    std::sort(pts.begin(),pts.end(),[](float2 a,float2 b){return a.x<b.x||(a.x==b.x&&a.y<b.y);});
    uint32_t WIN = (trial&1)? 128: 16;
    for(int dir=0;dir<2;++dir){
      auto q=pts; if(dir) std::reverse(q.begin(),q.end());
      auto r=refchain(q), s=look(q,WIN);
      bool ok=r.size()==s.size(); for(size_t i=0;ok&&i<r.size();++i) ok=r[i].x==s[i].x&&r[i].y==s[i].y;
      if(!ok){printf("MISMATCH trial %d n %d dir %d mode %d: %zu vs %zu\n",trial,n,dir,mode,r.size(),s.size());return 1;}
    }
  }
  printf("ok pts/iter %.2f\n",(double)g_pts/g_iter);
}
