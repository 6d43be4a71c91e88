// Experiment (not product, not test): statistics of how much of the probe stream of config 3 can be cleared by
// coarse dilated occupancy levels.  Includes the oracle source to reuse its exact ray generation.
#include "../../oracle/vxo.cpp"
#include <vector>
#include <cstdio>
#include <algorithm>

namespace {
struct Lvl { int sh, cx, cy, cz; std::vector<uint8_t> occ, dil; 
  bool at(const std::vector<uint8_t>& a, int x,int y,int z) const { if (x<0||y<0||z<0||x>=cx||y>=cy||z>=cz) return false; return a[(size_t)x + (size_t)y*cx + (size_t)z*cx*cy]!=0; } };
Lvl build(const uint8_t* vol, int sx, int sy, int sz, int sh) {
  Lvl L; L.sh = sh; int tpc = 1 << (sh-1);
  L.cx=(sx+tpc-1)/tpc; L.cy=(sy+tpc-1)/tpc; L.cz=(sz+tpc-1)/tpc;
  L.occ.assign((size_t)L.cx*L.cy*L.cz,0);
  #pragma omp parallel for
  for (int z=0;z<sz;++z) for (int y=0;y<sy;++y) for (int x=0;x<sx;++x) if (vol[(size_t)x+(size_t)y*sx+(size_t)z*sx*sy]) L.occ[(size_t)(x/tpc)+(size_t)(y/tpc)*L.cx+(size_t)(z/tpc)*L.cx*L.cy]=1;
  L.dil.assign(L.occ.size(),0);
  #pragma omp parallel for
  for (int z=0;z<L.cz;++z) for (int y=0;y<L.cy;++y) for (int x=0;x<L.cx;++x) { bool b=false;
    for (int dz=-1;dz<=1&&!b;++dz) for(int dy=-1;dy<=1&&!b;++dy) for(int dx=-1;dx<=1&&!b;++dx) b = L.at(L.occ,x+dx,y+dy,z+dz);
    L.dil[(size_t)x+(size_t)y*L.cx+(size_t)z*L.cx*L.cy]=b; }
  return L;
}
}

extern "C" void exp_run(const uint8_t* vol, int sx, int sy, int sz, const vxo_view* view, const vxo_gbuffer* gb, int n_ao, int bstep, double* out) {
  const Luts& L = luts();
  vxo_volume V{vol, sx, sy, sz};
  std::vector<Lvl> lv; for (int sh=2; sh<=5; ++sh) lv.push_back(build(vol,sx,sy,sz,sh));   // cells 4,8,16,32
  const int W=gb->width,H=gb->height; const V3 SUN=sun_dir();
  // counters
  // [kind 0 sun/1 ao][phase][..]: probes, plain clear@4,8,16,32, dil clear@4,8,16,32
  double cnt[2][2][16]; memset(cnt,0,sizeof cnt);
  double nrays[2]={0,0}; double totsteps[2]={0,0};
  // strategy sims
  double lookA[2]={0,0};   // sphere-trace with a clearance-class lookup (1 lookup gives best class), + fine lookups
  double lookB[2]={0,0};   // candidates (fine L4 set)
  double fullclear2[2]={0,0}; // rays whose phase 2 is entirely dil-clear at some level per probe group
  double histA[2][64]; memset(histA,0,sizeof histA);
  for (int by=0; by<H/16; by+=bstep) for (int bx=0; bx<W/32; bx+=bstep) {
    for (int ly=0; ly<16; ++ly) for (int lx=0; lx<32; ++lx) {
      int px=bx*32+lx, py=by*16+ly; size_t idx=(size_t)py*W+px;
      float depth=unorm24(gb->depth24[idx]); if(!(depth<0.999f)) continue;
      Pixel p=pixel_setup(*view,W,H,px,py);
      V3 pos=p.farvec*(depth*(1.0f+1.0f/FAR_)); V3 normal=decode_normal(gb->normal[idx]); V3 wd=SUN;
      V3 wcp=xyz(mat_mul(view->InverseViewMatrix,V4{pos.x,pos.y,pos.z,1.0f}))*10.0f;
      uint32_t n=get_noise(*gb,*view,p,-1);
      V3 randomVec=cosine_sample_hemisphere(L,n,n>>8)*0.1f; randomVec.z*=gsign(unorm8(n>>16)-0.5f);
      wd=mix3(wd,randomVec,0.5f); wd=normalize3(wd); wcp=wcp+wd*(unorm8(n>>24)*1.0f); wcp=wcp+randomVec*2.5f;
      float bias=gsmoothstep(0.0f,0.2f,depth)*50.0f+1.5f; V3 origin=wcp+normal*bias;
      V3 tangent=fabsf(normal.z)>0.5f? v3(0.0f,-normal.z,normal.y):v3(-normal.y,normal.x,0.0f); V3 bitangent=cross3(normal,tangent);
      for (int r=0;r<=n_ao;++r) {
        int kind = r==0?0:1; V3 dir; float step0;
        if (r==0){dir=wd;step0=0.5f;} else { uint32_t ni=(r==1)?n:get_noise(*gb,*view,p,r-1); V3 rv=cosine_sample_hemisphere(L,ni,ni>>8); dir=tangent*rv.x+bitangent*rv.y+normal*rv.z; step0=2.5f; }
        nrays[kind]++;
        // enumerate probes (no early exit: stats of the full-length ray; ~95% of rays are misses anyway)
        std::vector<V3> P; std::vector<int> ph; V3 sd=dir*step0; V3 q=origin; float d=step0, sf=step0;
        while(d<16.0f){P.push_back(q);ph.push_back(0);q=q+sd;d+=sf;} sf*=2; sd=sd*2.0f; while(d<128.0f){P.push_back(q);ph.push_back(1);q=q+sd;d+=sf;}
        uint64_t ns=0; march(V,origin,dir,128.0f,step0,ns,nullptr); int np=std::min((int)P.size(),(int)ns); totsteps[kind]+=np;
        std::vector<int> cls(np);  // max dilated-clear class: 0 none, 1: c=4, 2: c=8, 3: c=16, 4: c=32
        std::vector<int> fine(np);
        for (int k=0;k<np;++k){ double* c=cnt[kind][ph[k]]; c[0]++; int best=0;
          for (int li=0;li<4;++li){ float cs=(float)(4<<li); int X=(int)floorf(P[k].x/cs),Y=(int)floorf(P[k].y/cs),Z=(int)floorf(P[k].z/cs);
            if(!lv[li].at(lv[li].occ,X,Y,Z)) c[1+li]++; if(!lv[li].at(lv[li].dil,X,Y,Z)) {c[5+li]++; best=li+1;} 
            if (li==0) fine[k]=lv[0].at(lv[0].occ,X,Y,Z); }
          cls[k]=best; }
        // strategy A: per phase, sphere trace
        int look=0, cand=0;
        for (int phs=0; phs<2; ++phs) {
          float stepv = step0*(phs?2.0f:1.0f); float dm = fmaxf(fmaxf(fabsf(dir.x),fabsf(dir.y)),fabsf(dir.z))*stepv;
          int k=0; while(k<np && ph[k]!=phs) ++k; int kend=k; while(kend<np && ph[kend]==phs) ++kend;
          while(k<kend){ look++; int c=cls[k]; if(c==0){ if(fine[k]) cand++; k++; } else { float R=(float)(4<<(c-1))-0.25f; int nskip=(int)floorf(R/dm); if (nskip>1000) nskip=1000; k+=1+nskip; } }
        }
        lookA[kind]+=look; lookB[kind]+=cand; histA[kind][std::min(look,63)]++;
      }
    }
  }
  int o=0; for(int a=0;a<2;++a)for(int b=0;b<2;++b)for(int c=0;c<9;++c) out[o++]=cnt[a][b][c];
  out[o++]=nrays[0]; out[o++]=nrays[1]; out[o++]=totsteps[0]; out[o++]=totsteps[1]; out[o++]=lookA[0]; out[o++]=lookA[1]; out[o++]=lookB[0]; out[o++]=lookB[1];
  for(int a=0;a<2;++a)for(int c=0;c<64;++c) out[o++]=histA[a][c];
}
