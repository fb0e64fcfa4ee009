// Hot-loop design probe #2 (development, not product): R rays/thread, Q sphere pairs per NaN check,
// NACC NaN-sticky accumulators.  Sphere records: per group of Q pairs, 3*Q*2 floats (gx[],gy[],gz[]).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b){ f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template<int R, int Q, int NACC, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS,1) loopk(int n_groups, int reps, const float* __restrict__ g, float* sink){
  extern __shared__ __align__(16) float s[];
  for(int i=threadIdx.x;i<n_groups*Q*6;i+=THREADS) s[i]=g[i];
  __syncthreads();
  float ex[R],ey[R],ez[R];
  #pragma unroll
  for(int r=0;r<R;++r){ ex[r]=0.01f*(threadIdx.x+r); ey[r]=0.02f*(r+threadIdx.x)+0.3f; ez[r]=0.5f+0.001f*(r+2*threadIdx.x); }
  const f32x2 Z=pack2(0.f,0.f);
  int flagged=0;
  for(int rep=0;rep<reps;++rep){
    f32x2 acc[NACC];
    #pragma unroll
    for(int a=0;a<NACC;++a) acc[a]=Z;
    #pragma unroll UNROLL
    for(int p=0;p<n_groups;++p){
      const float* rec = s + p*Q*6;
      f32x2 u[Q][R];
      #pragma unroll
      for(int q=0;q<Q;++q){
        const f32x2 GX=pack2(rec[2*q],rec[2*q+1]), GY=pack2(rec[2*Q+2*q],rec[2*Q+2*q+1]), GZ=pack2(rec[4*Q+2*q],rec[4*Q+2*q+1]);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=mul2(pack2(ex[r],ex[r]),GX);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=fma2(pack2(ey[r],ey[r]),GY,u[q][r]);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=fma2(pack2(ez[r],ez[r]),GZ,u[q][r]);
        #pragma unroll
        for(int r=0;r<R;++r) acc[r%NACC]=fma2(u[q][r],Z,acc[r%NACC]);
      }
      f32x2 t=acc[0];
      #pragma unroll
      for(int a=1;a<NACC;++a) t=add2(t,acc[a]);
      float lo,hi; unpack2(t,lo,hi);
      if(!(lo==hi)){ flagged++;
        #pragma unroll
        for(int q=0;q<Q;++q){
        #pragma unroll
        for(int r=0;r<R;++r){ float a,b; unpack2(u[q][r],a,b); if(!(fabsf(a)<=3e38f)) flagged+=r; if(!(fabsf(b)<=3e38f)) flagged+=2*r+q; } }
        #pragma unroll
        for(int a=0;a<NACC;++a) acc[a]=Z; }
    }
  }
  if(flagged==12345) sink[0]=flagged;
}
template<int R,int Q,int NACC,int UNROLL,int THREADS> void run(const char* name,int sms,const float* g,float* sink){
  const int n_spheres=1024, n_groups=n_spheres/(2*Q), reps=(R==16?100:200);
  auto k=loopk<R,Q,NACC,UNROLL,THREADS>;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,64*1024);
  cudaFuncAttributes at; cudaFuncGetAttributes(&at,k);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<<<sms,THREADS,n_groups*Q*24>>>(n_groups,2,g,sink);
  cudaEventRecord(a); k<<<sms,THREADS,n_groups*Q*24>>>(n_groups,reps,g,sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double tests=(double)sms*THREADS*reps*(double)n_spheres*R;
  printf("%-34s regs %3d: %5.1f TFLOP/s algorithmic, %.3f cycles/test/lane-slot (ideal 4.000)\n",name,at.numRegs,tests*7/(ms*1e-3)/1e12, ms*1e-3*1.965e9*sms*128/tests);
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,0);
  float* g; cudaMalloc(&g,1<<20); cudaMemset(g,0,1<<20); float* sink; cudaMalloc(&sink,16);
  run<8,1,1,2,512>("R8  Q1 nacc1 u2",sms,g,sink);
  run<8,1,2,2,512>("R8  Q1 nacc2 u2",sms,g,sink);
  run<8,2,1,1,512>("R8  Q2 nacc1 u1",sms,g,sink);
  run<8,2,2,1,512>("R8  Q2 nacc2 u1",sms,g,sink);
  run<8,2,2,2,512>("R8  Q2 nacc2 u2",sms,g,sink);
  run<8,2,4,1,512>("R8  Q2 nacc4 u1",sms,g,sink);
  run<8,4,2,1,512>("R8  Q4 nacc2 u1",sms,g,sink);
  run<16,1,2,1,512>("R16 Q1 nacc2 u1",sms,g,sink);
  run<16,1,2,2,512>("R16 Q1 nacc2 u2",sms,g,sink);
  run<16,1,4,2,512>("R16 Q1 nacc4 u2",sms,g,sink);
  run<16,2,2,1,256>("R16 Q2 nacc2 u1 (256 thr)",sms,g,sink);
  run<16,1,2,2,256>("R16 Q1 nacc2 u2 (256 thr)",sms,g,sink);
  run<8,2,2,1,256>("R8  Q2 nacc2 u1 (256 thr)",sms,g,sink);
  run<8,2,2,1,1024>("R8  Q2 nacc2 u1 (1024 thr)",sms,g,sink);
  return 0;
}
