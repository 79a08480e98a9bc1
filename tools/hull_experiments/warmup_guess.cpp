// CPU experiment: can a block of the chain start from a WARM-UP guess of the stack (machine run from empty over the W points
// before the block) and provably make the true decisions? Validity: the top d entries of the guess (d = entries the block's
// run looked at) equal the top d entries of the true stack at the block start.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
struct P{float x,y;};
#define M 1e-4f
struct L{float l0,l1,l2;};
static L line(P a,P b){return {a.y*b.x-a.x*b.y,b.y-a.y,a.x-b.x};}
static float side(const L&l,P p){return (l.l0+p.x*l.l1)+p.y*l.l2;}
// runs points [s,e) on stack st; returns lowest index looked at
static size_t run(const std::vector<P>&pts,size_t s,size_t e,std::vector<P>&st){
  size_t low=st.size();
  for(size_t k=s;k<e;++k){P p=pts[k];
    while(true){ if(st.size()<2){low=0;break;} low=std::min(low,st.size()-2); if(!(side(line(st[st.size()-2],st[st.size()-1]),p)<=M))break; st.pop_back(); }
    st.push_back(p);}
  return low;
}
static bool eq(P a,P b){return a.x==b.x&&a.y==b.y;}
int main(){
  FILE*f=fopen("/tmp/hx/pts.bin","rb"); uint32_t ns; fread(&ns,4,1,f);
  const int B=16; const size_t Ws[]={64,128,256,512,1024,100000};
  for(size_t W:Ws){ long ok=0,tot=0; double davg=0; 
  fseek(f,4,SEEK_SET);
  for(uint32_t s=0;s<ns;++s){ uint32_t n; fread(&n,4,1,f); std::vector<P> pts(n); fread(pts.data(),8,n,f);
    for(int dir=0;dir<2;++dir){ if(dir) std::reverse(pts.begin(),pts.end());
      size_t bs=(n+B-1)/B;
      std::vector<P> truth; 
      for(int j=0;j<B;++j){ size_t b0=j*bs, b1=std::min<size_t>(n,(j+1)*bs); if(b0>=n)break;
        if(j>0){
          std::vector<P> g; size_t w0 = b0>W? b0-W:0; run(pts,w0,b0,g);
          size_t gsz=g.size(); size_t low=run(pts,b0,b1,g); size_t d=gsz-std::min(low,gsz);
          bool valid = low>0 && truth.size()>=d;   // low==0: needs the whole stack and its size -> only valid if identical
          if(low==0){ valid = truth.size()==gsz; d=gsz; }
          std::vector<P> g0; run(pts,w0,b0,g0);
          for(size_t i=0;valid&&i<d;++i) valid = eq(truth[truth.size()-1-i], g0[g0.size()-1-i]);
          ok+=valid; tot++; davg+=d;
        }
        run(pts,b0,b1,truth);
      }
    }
  }
  printf("W=%zu: %ld of %ld blocks valid (%.1f%%), mean depth looked at %.1f\n",W,ok,tot,100.0*ok/tot,davg/tot);
  }
}
