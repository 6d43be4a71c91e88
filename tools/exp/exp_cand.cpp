// Experiment (not product, not test): per-probe-index statistics of the AO rays of config 3 -- how many probes land in
// an occupied 4-voxel cell / a non-zero texel / hit, and how far from the block's origin centre they are.
#include "../../oracle/vxo.cpp"
#include <vector>
#include <cstdio>
#include <algorithm>

extern "C" void exp_cand(const uint8_t* vol, int sx, int sy, int sz, const vxo_view* view, const vxo_gbuffer* gb, int n_ao, int bstep, double* out) {
  const Luts& L = luts();
  vxo_volume V{vol, sx, sy, sz};
  const int W=gb->width,H=gb->height; const V3 SUN=sun_dir();
  auto texel = [&](V3 p)->unsigned { int x=(int)floorf(p.x/2), y=(int)floorf(p.y/2), z=(int)floorf(p.z/2); if(x<0||y<0||z<0||x>=sx||y>=sy||z>=sz) return 0u; return vol[(size_t)x+(size_t)y*sx+(size_t)z*sx*sy]; };
  auto cell4 = [&](V3 p)->bool { int x=(int)floorf(p.x/4)*2, y=(int)floorf(p.y/4)*2, z=(int)floorf(p.z/4)*2; for(int c=0;c<8;++c){int X=x+(c&1),Y=y+((c>>1)&1),Z=z+(c>>2); if(X<0||Y<0||Z<0||X>=sx||Y>=sy||Z>=sz) continue; if(vol[(size_t)X+(size_t)Y*sx+(size_t)Z*sx*sy]) return true;} return false; };
  // out[k*8 + ...]: performed, cell4 set, texel set, hit, cand_near32, cand_near48, cand_near64, (unused)
  // out[29*8 + r]: rays with r candidates (cell4) ; out[29*8+32+..]: spread histogram
  double cnt[29][8]; memset(cnt,0,sizeof cnt);
  double spreadh[32]; memset(spreadh,0,sizeof spreadh);
  double nrays=0, nhit=0;
  for (int by=0; by<H/16; by+=bstep) for (int bx=0; bx<W/32; bx+=bstep) {
    // pass 1: origins bbox
    float lo[3]={1e30f,1e30f,1e30f}, hi[3]={-1e30f,-1e30f,-1e30f}; bool any=false;
    std::vector<V3> origins(512); std::vector<char> lit(512,0); std::vector<V3> normals(512); std::vector<uint32_t> noises(512);
    for (int ly=0; ly<16; ++ly) for (int lx=0; lx<32; ++lx) {
      int px=bx*32+lx, py=by*16+ly; size_t idx=(size_t)py*W+px;
      float depth=unorm24(gb->depth24[idx]); if(!(depth<0.999f)) continue;
      Pixel p=pixel_setup(*view,W,H,px,py);
      V3 pos=p.farvec*(depth*(1.0f+1.0f/FAR_)); V3 normal=decode_normal(gb->normal[idx]); V3 wd=SUN;
      V3 wcp=xyz(mat_mul(view->InverseViewMatrix,V4{pos.x,pos.y,pos.z,1.0f}))*10.0f;
      float bias=gsmoothstep(0.0f,0.2f,depth)*50.0f+1.5f;
      V3 hint=wcp+normal*bias;
      uint32_t n=get_noise(*gb,*view,p,-1);
      V3 randomVec=cosine_sample_hemisphere(L,n,n>>8)*0.1f; randomVec.z*=gsign(unorm8(n>>16)-0.5f);
      wd=mix3(wd,randomVec,0.5f); wd=normalize3(wd); wcp=wcp+wd*(unorm8(n>>24)*1.0f); wcp=wcp+randomVec*2.5f;
      V3 origin=wcp+normal*bias;
      int t=ly*32+lx; origins[t]=origin; lit[t]=1; normals[t]=normal; noises[t]=n; any=true;
      float h[3]={hint.x,hint.y,hint.z}; for(int a=0;a<3;++a){lo[a]=std::min(lo[a],floorf(h[a])); hi[a]=std::max(hi[a],floorf(h[a]));}
    }
    if(!any) continue;
    float ctr[3]; float spread=0; for(int a=0;a<3;++a){ctr[a]=floorf((lo[a]+hi[a])*0.5f); spread=std::max(spread,hi[a]-lo[a]);}
    spreadh[std::min(31,(int)(spread/4))]++;
    for (int ly=0; ly<16; ++ly) for (int lx=0; lx<32; ++lx) { int t=ly*32+lx; if(!lit[t]) continue;
      int px=bx*32+lx, py=by*16+ly; Pixel p=pixel_setup(*view,W,H,px,py);
      V3 normal=normals[t], origin=origins[t]; uint32_t n=noises[t];
      V3 tangent=fabsf(normal.z)>0.5f? v3(0.0f,-normal.z,normal.y):v3(-normal.y,normal.x,0.0f); V3 bitangent=cross3(normal,tangent);
      for (int r=0;r<n_ao;++r) {
        uint32_t ni=(r==0)?n:get_noise(*gb,*view,p,r); V3 rv=cosine_sample_hemisphere(L,ni,ni>>8); V3 dir=tangent*rv.x+bitangent*rv.y+normal*rv.z;
        nrays++;
        V3 sd=dir*2.5f; V3 q=origin; int nc=0; bool hit=false;
        for (int k=0;k<29 && !hit;++k) {
          double* c=cnt[k]; c[0]++;
          bool c4=cell4(q); unsigned tx=texel(q);
          float dist=std::max(std::max(fabsf(q.x-ctr[0]),fabsf(q.y-ctr[1])),fabsf(q.z-ctr[2]));
          if(c4){c[1]++; nc++;}
          if(tx) c[2]++;
          if (k<6) { if(tx){ unsigned bit=0; bit+=gmod(q.x,0.5f)>0.25f?1u:0u; bit+=gmod(q.y,0.5f)>0.25f?2u:0u; bit+=gmod(q.z,0.5f)>0.25f?4u:0u; hit=(tx>>bit)&1u; } }
          else hit = tx!=0;
          if(hit) c[3]++;
          // candidates that still need a fetch with a texel-exact near window of half-size R: inside: phase1 needs fetch iff texel set; phase2 never.  outside: iff cell4
          const float R[3]={31.f,47.f,63.f};
          for(int w=0;w<3;++w){ bool need; if(dist<R[w]) need = (k<6)? (tx!=0) : false; else need=c4; if(need) c[4+w]++; }
          q=q+(k<6?sd:sd*2.0f);
        }
        if(hit) nhit++;
      }
    }
  }
  int o=0; for(int k=0;k<29;++k)for(int c=0;c<8;++c) out[o++]=cnt[k][c];
  out[o++]=nrays; out[o++]=nhit; for(int i=0;i<32;++i) out[o++]=spreadh[i];
}
