// tools/simt_sim_slots.cpp -- SIMT schedule simulator, fully voted scheduling with K path slots per lane (the model behind
// slot_kernels.cu): node / leaf / switch / shade / camera operations chosen by thresholds.  Input as simt_sim_phases.cpp.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
using namespace std;
static int N_NODE=100, LEAF_BASE=20, LEAF_PER=55, SHADE=230, CAMERA=110, FETCH=30, LOOP=4, SWITCH=40, VOTE=4, SCHED=14;
static vector<uint8_t> seq; static vector<uint64_t> pstart; static vector<uint32_t> pix_lo; // path starts; per pixel first path index
static uint64_t nxt_pix;
enum { sR=0, sT=1, sS=2, sDead=3 };
struct Lane { uint32_t path_next, path_end; bool has_pix; int cur; };
struct P { int k, TN, TL, TS, TR; bool foldW; };
static double run(int n_warps, P prm, bool verbose){
  int k=prm.k; nxt_pix=0;
  struct Warp { vector<uint64_t> p; vector<uint8_t> st; Lane L[32]; };
  vector<Warp> W(n_warps);
  for(auto&w:W){ w.p.assign(32*k,0); w.st.assign(32*k,sR); for(int l=0;l<32;l++){ w.L[l].has_pix=false; w.L[l].cur=-1; w.L[l].path_next=w.L[l].path_end=0; } }
  double cost=0,cN=0,cL=0,cW=0,cS=0,cR=0,cV=0; uint64_t segs=0; double lnN=0,itN=0,lnL=0,itL=0,lnS=0,itS=0,lnR=0,itR=0,lnW=0,itW=0;
  uint64_t npix=pix_lo.size()-1;
  bool any=true;
  while(any){ any=false;
    for(auto&w:W){
      // run this warp for a bounded number of scheduler decisions (round robin between warps so the pixel ticket is shared fairly)
      for(int rep=0;rep<64;rep++){
        int nN=0,nL=0,nW=0,nS=0,nR=0,mx=0;
        for(int l=0;l<32;l++){
          Lane&L=w.L[l]; bool hasT=false,hasS=false,hasR=false;
          for(int j=0;j<k;j++){ uint8_t s=w.st[l*k+j]; if(s==sT && j!=L.cur) hasT=true; if(s==sS) hasS=true; if(s==sR) hasR=true; }
          if(L.cur>=0){ uint8_t t=seq[w.p[l*k+L.cur]]; if(t==0) nN++; else if(t<=8){ nL++; mx=max(mx,(int)t);} else { /* finished: needs switch (store result); counts as W */ nW++; } }
          else if(hasT) nW++;
          if(hasS) nS++;
          if(hasR && (L.has_pix && L.path_next<L.path_end || nxt_pix<npix)) nR++;
        }
        if(nN+nL+nW+nS+nR==0) break;
        any=true;
        cost+=VOTE; cV+=VOTE;
        int op=-1; // 0 N 1 L 2 W 3 S 4 R
        if(nN>=prm.TN) op=0;
        else {
          cost+=SCHED; cV+=SCHED;
          // candidates by threshold, else most participants
          if(nL>=prm.TL) op=1; else if(nW>=8) op=2; else if(nS>=prm.TS) op=3; else if(nR>=prm.TR) op=4;
          else { int best=0; op=0; int c[5]={nN,nL,nW,nS,nR}; for(int i=0;i<5;i++) if(c[i]>best){best=c[i];op=i;} }
        }
        auto do_switch=[&](int l){ Lane&L=w.L[l]; if(L.cur>=0){ uint8_t t=seq[w.p[l*k+L.cur]]; if(t==255){ w.st[l*k+L.cur]=sS; L.cur=-1; } else return; }
                                   for(int j=0;j<k;j++) if(w.st[l*k+j]==sT){ L.cur=j; break; } };
        if(op==0){ cost+=N_NODE+LOOP; cN+=N_NODE+LOOP; lnN+=nN; itN++; for(int l=0;l<32;l++){ Lane&L=w.L[l]; if(L.cur>=0 && seq[w.p[l*k+L.cur]]==0) w.p[l*k+L.cur]++; } }
        else if(op==1){ double c=LOOP+LEAF_BASE+LEAF_PER*mx; cost+=c; cL+=c; lnL+=nL; itL++; for(int l=0;l<32;l++){ Lane&L=w.L[l]; if(L.cur>=0){ uint8_t t=seq[w.p[l*k+L.cur]]; if(t>=1&&t<=8) w.p[l*k+L.cur]++; } } }
        else if(op==2){ cost+=SWITCH; cW+=SWITCH; lnW+=nW; itW++; for(int l=0;l<32;l++) do_switch(l); }
        else if(op==3){ cost+=SHADE; cS+=SHADE; lnS+=nS; itS++;
          for(int l=0;l<32;l++){ for(int j=0;j<k;j++) if(w.st[l*k+j]==sS){ uint64_t&p=w.p[l*k+j]; p++; segs++; uint8_t t=seq[p]; w.st[l*k+j]=(t==253||t==254)?sR:sT; break; } }
          if(prm.foldW){ cost+=SWITCH; cW+=SWITCH; for(int l=0;l<32;l++) if(w.L[l].cur<0) do_switch(l); } }
        else if(op==4){ cost+=CAMERA; cR+=CAMERA; lnR+=nR; itR++; bool f=false;
          for(int l=0;l<32;l++){ Lane&L=w.L[l]; for(int j=0;j<k;j++) if(w.st[l*k+j]==sR){
              if(!(L.has_pix && L.path_next<L.path_end)){ if(nxt_pix<npix){ L.path_next=pix_lo[nxt_pix]; L.path_end=pix_lo[nxt_pix+1]; nxt_pix++; L.has_pix=true; f=true; } else break; }
              w.p[l*k+j]=pstart[L.path_next++]+1; w.st[l*k+j]=sT; break; } }
          if(f){ cost+=FETCH; cR+=FETCH; }
          if(prm.foldW){ cost+=SWITCH; cW+=SWITCH; for(int l=0;l<32;l++) if(w.L[l].cur<0) do_switch(l); } }
      }
    }
  }
  if(verbose) printf("   node %.1f (u %.1f) leaf %.1f (u %.1f) switch %.1f (u %.1f) shade %.1f (u %.1f) regen %.1f (u %.1f) vote %.1f\n",cN/segs,lnN/itN,cL/segs,lnL/itL,cW/segs,lnW/max(1.0,itW),cS/segs,lnS/itS,cR/segs,lnR/itR,cV/segs);
  return cost/segs;
}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); seq.resize(n); if(fread(seq.data(),1,n,f)!=(size_t)n) return 1; fclose(f);
  pix_lo.push_back(0);
  for(long i=0;i<n;i++){ if(seq[i]==253) pstart.push_back(i); else if(seq[i]==254) pix_lo.push_back(pstart.size()); }
  int nw=argc>2?atoi(argv[2]):16;
  if(argc>3){ N_NODE=atoi(argv[3]); }
  if(getenv("ONE")){ int k,TN,TL,TS,fold; sscanf(getenv("ONE"),"%d,%d,%d,%d,%d",&k,&TN,&TL,&TS,&fold); P p{k,TN,TL,TS,TS,(bool)fold}; double c=run(nw,p,true); printf("%.1f\n",c); return 0; }
  for(int k: {1,2,3,4,6}) for(int TN: {16,20,24}) for(int TL: {8,12,16}) for(int TS: {12,20}) for(int fold=0; fold<2; fold++){
    P p{k,TN,TL,TS,TS,(bool)fold};
    double c=run(nw,p,false); printf("C k=%d TN=%d TL=%d TS=%d fold=%d: %.1f\n",k,TN,TL,TS,fold,c);
  }
}
