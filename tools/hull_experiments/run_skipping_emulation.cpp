// Host emulation of the warp-cooperative run-skipping hull chain vs the reference loop.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
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
static long g_steps=0, g_points=0;
std::vector<float2> warpchain(const std::vector<float2>& pts){
  uint32_t N=pts.size();
  std::vector<float2> S(N+64);
  if(N<3){ return pts; }
  S[0]=pts[0]; S[1]=pts[1]; uint32_t n=2, k=2;
  float2 a=S[0], b=S[1], c=a; HullLine lca=hull_line(c,a);
  while(k<N){
    g_steps++;
    float2 q[32]; bool valid[32];
    for(int l=0;l<32;l++){ uint32_t idx=k+l; valid[l]=idx<N; q[l]=valid[l]?pts[idx]:float2{0,0}; }
    uint32_t mP=0,mK=0;
    for(int l=0;l<32;l++){
      float2 prev = l==0? b : q[l-1];
      float2 prev2 = l==0? a : (l==1? b : q[l-2]);
      float tA=hull_side(hull_line(a,prev),q[l]); bool A=tA<=M;
      float tB=hull_side(lca,q[l]); bool B=(n==2)||!(tB<=M);
      float tK=hull_side(hull_line(prev2,prev),q[l]); bool K=!(tK<=M);
      if(valid[l]&&A&&B) mP|=1u<<l;
      if(valid[l]&&K) mK|=1u<<l;
    }
    auto cto=[](uint32_t m){ return m==0xffffffffu?32:__builtin_ctz(~m); };
    uint32_t L1=cto(mP), L2=cto(mK);
    if(L1>0){ b=q[L1-1]; S[n-1]=b; k+=L1; g_points+=L1; }
    else if(L2>0){
      for(uint32_t l=0;l<L2;l++) S[n+l]=q[l];
      float2 nb=q[L2-1];
      float2 na = L2>=2? q[L2-2] : b;
      float2 nc = L2>=3? q[L2-3] : (L2==2? b : a);
      a=na; b=nb; c=nc; lca=hull_line(c,a); n+=L2; k+=L2; g_points+=L2;
    } else {
      float2 q0=q[0];
      // multi-pop: at least b and a go
      uint32_t base=2; uint32_t rstar=0; float2 na{0,0}, nc{0,0}; bool found=false;
      while(!found){
        uint32_t stopmask=0; float2 e1[32],e2[32];
        for(int l=0;l<32;l++){
          int r=(int)n-(int)base-l;   // remaining size after popping base+l entries
          bool stop=false;
          if(r>=1){
            if(r==1) stop=true;
            else { e1[l]=S[r-2]; e2[l]=S[r-1]; stop=!(hull_side(hull_line(e1[l],e2[l]),q0)<=M); }
            if(r>=1 && r<2) e2[l]=S[r-1];
          }
          if(stop) stopmask|=1u<<l;
        }
        if(stopmask){ int l=__builtin_ctz(stopmask); rstar=n-base-l; na=e2[l]; nc=e1[l]; found=true; }
        else base+=32;
      }
      S[rstar]=q0; n=rstar+1; b=q0; a=na; c=nc; if(n>=3) lca=hull_line(c,a); k+=1; g_points+=1;
    }
  }
  S.resize(n); return S;
}
int main(){
  std::mt19937 rng(1);
  for(int trial=0;trial<40000;++trial){
    int n=1+rng()%200; std::vector<float2> pts(n);
    int mode=rng()%5;
    for(auto&p:pts){ if(mode==0){p.x=(rng()%1000)/100.f;p.y=(rng()%1000)/100.f;} else if(mode==1){p.x=(rng()%8);p.y=(rng()%8);} else if(mode==2){p.x=(rng()%100)/10.f; p.y=p.x*0.5f+(rng()%3)*1e-4f;} else if(mode==3){p.x=(rng()%20)*0.01f;p.y=(rng()%20)*0.01f;} else { float t=(rng()%10000)/10000.f*6.2831853f; p.x=100*cosf(t); p.y=100*sinf(t);} }
    std::sort(pts.begin(),pts.end(),[](float2 a,float2 b){return a.x<b.x||(a.x==b.x&&a.y<b.y);});
    for(int dir=0;dir<2;++dir){
      auto q=pts; if(dir) std::reverse(q.begin(),q.end());
      auto r=refchain(q), s=warpchain(q);
      bool ok=r.size()==s.size(); for(size_t i=0;ok&&i<r.size();++i) ok=r[i].x==s[i].x&&r[i].y==s[i].y;
      if(!ok){printf("MISMATCH trial %d n %d dir %d mode %d: %zu vs %zu\n",trial,n,dir,mode,r.size(),s.size());return 1;}
    }
  }
  printf("ok points/step %.2f\n",(double)g_points/g_steps);
}
