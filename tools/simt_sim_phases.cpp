// tools/simt_sim_phases.cpp -- SIMT schedule simulator (phase-structured schedules): replays the per-pixel token streams dumped by
// tools/simt_dump.py (tests/host_harness.cpp::hh_step_sequences) through warp-level schedules -- while-while, voted node/leaf
// turns, time slices, K rays per lane with phase barriers -- and reports model warp-instructions per ray segment.
// Build: g++ -O2 -o simt_sim_phases tools/simt_sim_phases.cpp ; run: ./simt_sim_phases seq_wsah1.bin 8 85
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>
using namespace std;
static int N_NODE=85, LEAF_BASE=25, LEAF_PER=45, SHADE=230, CAMERA=110, FETCH=30, LOOP=4, SWITCH=40, VOTE=4;
static vector<uint8_t> seq; static vector<uint64_t> pix;
struct Slot { uint64_t p; bool live; };
static uint64_t nxt; static int KSD=1;
static bool fetch(Slot& s){ if(nxt<pix.size()){ s.p=pix[nxt++]; s.live=true; return true;} s.live=false; return false; }

// Model A: current kernel (1 slot per lane, while-while)
static double modelA(int n_warps, uint64_t& segs_out, int policy=0, int T=8, int KS=1000000){
  nxt=0; vector<vector<Slot>> W(n_warps, vector<Slot>(32));
  for(auto&w:W) for(auto&s:w) fetch(s);
  double cost=0; uint64_t segs=0; double lane_node=0, it_node=0;
  bool any=true;
  while(any){ any=false;
    for(auto&w:W){
      bool alive=false; for(auto&s:w) alive|=s.live; if(!alive) continue; any=true;
      bool f=false; for(auto&s:w) if(s.live && seq[s.p]==254){ f=true; s.p++; /*dummy*/ fetch(s);} if(f) cost+=FETCH;
      bool c=false; for(auto&s:w) if(s.live && seq[s.p]==253){ c=true; s.p++; } if(c) cost+=CAMERA;
      if(policy==0) for(;;){
        bool anyn=false, anyl=false;
        for(auto&s:w) if(s.live){ uint8_t t=seq[s.p]; if(t==0) anyn=true; else if(t>=1&&t<=8) anyl=true; }
        if(!anyn && !anyl) break;
        for(;;){ int n=0; for(auto&s:w) if(s.live && seq[s.p]==0){ s.p++; n++; } if(!n) break; cost+=N_NODE+2; lane_node+=n; it_node+=1; }
        for(;;){ int mx=0; for(auto&s:w) if(s.live){ uint8_t t=seq[s.p]; if(t>=1&&t<=8){ mx=max(mx,(int)t); s.p++; } } if(!mx) break; cost+=LOOP+LEAF_BASE+LEAF_PER*mx; }
      } else { int slice=0; for(;;){
        // vote: node step unless >= T lanes wait at a leaf (T=1: if-if)
        int nn=0,nl=0,mx=0,nd=0;
        for(auto&s:w) if(s.live){ uint8_t t=seq[s.p]; if(t==0) nn++; else if(t>=1&&t<=8){ nl++; mx=max(mx,(int)t);} else nd++; }
        if(!nn && !nl) break;
        if(++slice>KS && nd>=KSD) break;
        cost+=VOTE;
        if(nn && nl<T){ cost+=N_NODE+2; lane_node+=nn; it_node+=1; for(auto&s:w) if(s.live && seq[s.p]==0) s.p++; }
        else { cost+=LOOP+LEAF_BASE+LEAF_PER*mx; for(auto&s:w) if(s.live){ uint8_t t=seq[s.p]; if(t>=1&&t<=8) s.p++; } }
      } }
      bool sh=false; for(auto&s:w) if(s.live && seq[s.p]==255){ sh=true; s.p++; segs++; } if(sh) cost+=SHADE;
    }
  }
  segs_out=segs; fprintf(stderr,"  [A node-loop lane util %.1f/32]\n", lane_node/it_node);
  return cost/segs;
}

// Model B: k slots per lane, phases: regen all / traverse all (concatenated, vote-scheduled) / shade all.
// T = node step runs while (#lanes blocked on leaf or switch) < T;  foldSwitch: switches execute together with the leaf step
static double modelB(int n_warps, int k, int T, bool foldSwitch, uint64_t& segs_out, bool verbose=false){
  nxt=0; vector<vector<Slot>> W(n_warps, vector<Slot>(32*k));
  for(auto&w:W) for(auto&s:w) fetch(s);
  double cost=0, c_node=0,c_leaf=0,c_sw=0,c_shade=0,c_regen=0; uint64_t segs=0; double lane_node=0,it_node=0,lane_leaf=0,it_leaf=0;
  bool any=true;
  while(any){ any=false;
    for(auto&w:W){
      bool alive=false; for(auto&s:w) alive|=s.live; if(!alive) continue; any=true;
      // regen phase
      for(int j=0;j<k;j++){
        bool f=false,c=false;
        for(int l=0;l<32;l++){ Slot&s=w[l*k+j]; if(s.live && seq[s.p]==254){ f=true; fetch(s);} if(s.live && seq[s.p]==253){ c=true; s.p++; } }
        if(f){ cost+=FETCH; c_regen+=FETCH;} if(c){ cost+=CAMERA; c_regen+=CAMERA; }
      }
      // traverse phase
      int cur[32]; for(int l=0;l<32;l++){ cur[l]=0; while(cur[l]<k && !w[l*k+cur[l]].live) cur[l]++; }
      for(;;){
        int nN=0,nL=0,nW=0, mx=0;
        for(int l=0;l<32;l++){ if(cur[l]>=k) continue; uint8_t t=seq[w[l*k+cur[l]].p]; if(t==0) nN++; else if(t<=8){ nL++; mx=max(mx,(int)t);} else nW++; }
        if(nN+nL+nW==0) break;
        cost+=VOTE;
        bool doNode = nN>0 && (nL+nW)<T;
        if(doNode){ cost+=N_NODE+LOOP; c_node+=N_NODE+LOOP; lane_node+=nN; it_node+=1; for(int l=0;l<32;l++){ if(cur[l]>=k) continue; Slot&s=w[l*k+cur[l]]; if(seq[s.p]==0) s.p++; } }
        else {
          bool doLeaf = nL>0 && (foldSwitch || nL>=nW);
          bool doSw = nW>0 && (foldSwitch || !doLeaf);
          if(doLeaf){ double c=LOOP+LEAF_BASE+LEAF_PER*mx; cost+=c; c_leaf+=c; lane_leaf+=nL; it_leaf+=1; for(int l=0;l<32;l++){ if(cur[l]>=k) continue; Slot&s=w[l*k+cur[l]]; uint8_t t=seq[s.p]; if(t>=1&&t<=8) s.p++; } }
          if(doSw){ cost+=SWITCH; c_sw+=SWITCH; for(int l=0;l<32;l++){ if(cur[l]>=k) continue; Slot&s=w[l*k+cur[l]]; if(seq[s.p]==255){ cur[l]++; while(cur[l]<k && !w[l*k+cur[l]].live) cur[l]++; } } }
        }
      }
      // shade phase
      for(int j=0;j<k;j++){ bool sh=false; for(int l=0;l<32;l++){ Slot&s=w[l*k+j]; if(s.live && seq[s.p]==255){ sh=true; s.p++; segs++; } } if(sh){ cost+=SHADE; c_shade+=SHADE; } }
    }
  }
  segs_out=segs;
  if(verbose) fprintf(stderr,"  [B k=%d T=%d fold=%d: node %.1f (util %.1f/32) leaf %.1f (util %.1f/32) switch %.1f shade %.1f regen %.1f]\n",k,T,foldSwitch,c_node/segs,lane_node/it_node,c_leaf/segs,lane_leaf/max(1.0,it_leaf),c_sw/segs,c_shade/segs,c_regen/segs);
  return cost/segs;
}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); seq.resize(n); fread(seq.data(),1,n,f); fclose(f);
  pix.push_back(0); for(long i=0;i+1<n;i++) if(seq[i]==254) pix.push_back(i+1);
  // ideal
  double id=0; uint64_t ns=0; for(long i=0;i<n;i++){ uint8_t t=seq[i]; if(t==0) id+=N_NODE; else if(t<=8) id+=LEAF_BASE+LEAF_PER*t; else if(t==255){ id+=SHADE; ns++; } else if(t==253) id+=CAMERA; }
  printf("pixels %zu segments %lu ideal %.1f warp-inst/seg\n",pix.size(),ns,id/ns/32);
  int nw = argc>2?atoi(argv[2]):32; uint64_t s; if(argc>3) N_NODE=atoi(argv[3]);
  double a=modelA(nw,s); printf("A (current): %.1f\n",a); for(int T: {8,12}){ double b=modelA(nw,s,1,T); printf("A vote T=%d: %.1f\n",T,b);} for(int KS: {0,2,4}) for(int kd: {18,20,22,24,26,28,30}){ KSD=kd; double b=modelA(nw,s,1,12,KS); printf("A vote T=12 slice=%d minDone=%d: %.1f\n",KS,kd,b);} return 0; for(int k: {1,2,3,4,8}) for(int T: {8,12}){ double b=modelB(nw,k,T,true,s,true); printf("B k=%d T=%d: %.1f\n",k,T,b);} return 0;
  for(int k: {4}) for(int T: {12}) for(int fold=0; fold<2; fold++){ double b=modelB(nw,k,T,fold,s,true); printf("B k=%d T=%2d fold=%d: %.1f\n",k,T,fold,b); }
}
