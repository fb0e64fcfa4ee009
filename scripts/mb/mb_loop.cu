// Hot-loop scheduling probe (development, not product): the v2 ray-kernel iteration with operands from
// shared memory, several accumulator arrangements.  NACC = number of NaN-sticky accumulators.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b){ f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#define T 512
template<int NACC, int UNROLL>
__global__ void __launch_bounds__(T,1) loopk(int n_pairs, int reps, const float4* __restrict__ g, float* sink){
  extern __shared__ float4 s[];
  for(int i=threadIdx.x;i<2*n_pairs;i+=T) s[i]=g[i];
  __syncthreads();
  float ex[8],ey[8],ez[8];
  #pragma unroll
  for(int r=0;r<8;++r){ ex[r]=0.01f*(threadIdx.x+r); ey[r]=0.02f*(r+threadIdx.x)+0.3f; ez[r]=0.5f+0.001f*(r+2*threadIdx.x); }
  const f32x2 Z=pack2(0.f,0.f);
  int flagged=0;
  for(int rep=0;rep<reps;++rep){
    f32x2 acc[NACC];
    #pragma unroll
    for(int a=0;a<NACC;++a) acc[a]=Z;
    #pragma unroll UNROLL
    for(int p=0;p<n_pairs;++p){
      const float4 A=s[2*p], B=s[2*p+1];
      const f32x2 GX=pack2(A.x,A.y), GY=pack2(A.z,A.w), GZ=pack2(B.x,B.y);
      f32x2 u[8];
      #pragma unroll
      for(int r=0;r<8;++r) u[r]=mul2(pack2(ex[r],ex[r]),GX);
      #pragma unroll
      for(int r=0;r<8;++r) u[r]=fma2(pack2(ey[r],ey[r]),GY,u[r]);
      #pragma unroll
      for(int r=0;r<8;++r) u[r]=fma2(pack2(ez[r],ez[r]),GZ,u[r]);
      #pragma unroll
      for(int r=0;r<8;++r) acc[r%NACC]=fma2(u[r],Z,acc[r%NACC]);
      f32x2 t=acc[0];
      #pragma unroll
      for(int a=1;a<NACC;++a) t=add2(t,acc[a]);
      float lo,hi; unpack2(t,lo,hi);
      if(!(lo==hi)){ flagged++; 
        #pragma unroll
        for(int a=0;a<NACC;++a) acc[a]=Z; }
    }
  }
  if(flagged==12345) sink[0]=flagged;
}
template<int NACC,int UNROLL> void run(const char* name,int sms,const float4* g,float* sink){
  const int n_pairs=512, reps=200;
  auto k=loopk<NACC,UNROLL>;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,64*1024);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<<<sms,T,2*n_pairs*16>>>(n_pairs,2,g,sink);
  cudaEventRecord(a); k<<<sms,T,2*n_pairs*16>>>(n_pairs,reps,g,sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double tests=(double)sms*T*reps*n_pairs*16;
  printf("%-28s: %.1f TFLOP/s algorithmic (7 FLOP/test), %.2f cycles/pair-iteration/SMSP at 1.965GHz\n",name,tests*7/(ms*1e-3)/1e12, ms*1e-3*1.965e9/((double)reps*n_pairs*4));
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,0);
  float4* g; cudaMalloc(&g,1024*16); cudaMemset(g,0,1024*16); float* sink; cudaMalloc(&sink,16);
  run<1,2>("nacc1 unroll2",sms,g,sink);
  run<2,2>("nacc2 unroll2",sms,g,sink);
  run<4,2>("nacc4 unroll2",sms,g,sink);
  run<8,2>("nacc8 unroll2",sms,g,sink);
  run<1,1>("nacc1 unroll1",sms,g,sink);
  run<1,4>("nacc1 unroll4",sms,g,sink);
  run<2,4>("nacc2 unroll4",sms,g,sink);
  return 0;
}
