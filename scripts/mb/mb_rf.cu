// Register-file / FMA-pipe microbenchmarks for sm_100a (development probe, not product code).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c){ float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
#define T 512
#define INNER 32

// K1: scalar FFMA, 3 fresh registers per instruction
__global__ void __launch_bounds__(T,1) k1(int iters, float seed, float* sink){
  float x[16], y[16], a[16];
  #pragma unroll
  for(int i=0;i<16;++i){ x[i]=seed+0.999f+i*1e-6f; y[i]=seed+1e-3f*(i+1); a[i]=seed+threadIdx.x+i; }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<INNER;++k){
      #pragma unroll
      for(int i=0;i<16;++i) a[i]=fmaf(x[i],a[i],y[i]);
    }
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<16;++i) s+=a[i];
  if(s==1234.5f) sink[0]=s;
}
// K2: FFMA2, 3 fresh pairs per instruction
__global__ void __launch_bounds__(T,1) k2(int iters, float seed, float* sink){
  f32x2 x[8], y[8], a[8];
  #pragma unroll
  for(int i=0;i<8;++i){ x[i]=pack2(seed+0.999f+i*1e-6f, seed+0.998f); y[i]=pack2(seed+1e-3f*(i+1), seed+2e-3f); a[i]=pack2(seed+threadIdx.x+i, seed-i); }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<INNER;++k){
      #pragma unroll
      for(int i=0;i<8;++i) a[i]=fma2(x[i],a[i],y[i]);
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<8;++i){ float lo,hi; unpack2(a[i],lo,hi); s+=lo+hi; }
  if(s==1234.5f) sink[0]=s;
}
// K3: FFMA2, scalar broadcast (fresh) + fresh pair + fresh accumulator
__global__ void __launch_bounds__(T,1) k3(int iters, float seed, float* sink){
  float sc[8]; f32x2 y[8], a[8];
  #pragma unroll
  for(int i=0;i<8;++i){ sc[i]=seed+0.999f+i*1e-6f; y[i]=pack2(seed+1e-3f*(i+1), seed+2e-3f); a[i]=pack2(seed+threadIdx.x+i, seed-i); }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<INNER;++k){
      #pragma unroll
      for(int i=0;i<8;++i) a[i]=fma2(pack2(sc[i],sc[i]),a[i],y[i]);
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<8;++i){ float lo,hi; unpack2(a[i],lo,hi); s+=lo+hi; }
  if(s==1234.5f) sink[0]=s;
}
// K4: FFMA2, scalar broadcast (fresh) + ONE shared pair + fresh accumulator (operand-major order)
__global__ void __launch_bounds__(T,1) k4(int iters, float seed, float* sink){
  float sc[8]; f32x2 y0, a[8];
  y0=pack2(seed+1e-3f, seed+2e-3f);
  #pragma unroll
  for(int i=0;i<8;++i){ sc[i]=seed+0.999f+i*1e-6f; a[i]=pack2(seed+threadIdx.x+i, seed-i); }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<INNER;++k){
      #pragma unroll
      for(int i=0;i<8;++i) a[i]=fma2(pack2(sc[i],sc[i]),y0,a[i]);
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<8;++i){ float lo,hi; unpack2(a[i],lo,hi); s+=lo+hi; }
  if(s==1234.5f) sink[0]=s;
}
// K5: K4 plus one FMNMX3 per 4 FFMA2 (ray-kernel ratio)
__global__ void __launch_bounds__(T,1) k5(int iters, float seed, float* sink){
  float sc[8]; f32x2 y0, a[8]; float m=-1.f;
  y0=pack2(seed+1e-3f, seed+2e-3f);
  #pragma unroll
  for(int i=0;i<8;++i){ sc[i]=seed+0.999f+i*1e-6f; a[i]=pack2(seed+threadIdx.x+i, seed-i); }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<INNER;++k){
      #pragma unroll
      for(int i=0;i<8;++i) a[i]=fma2(pack2(sc[i],sc[i]),y0,a[i]);
      if((k&3)==3){
        #pragma unroll
        for(int i=0;i<8;++i){ float lo,hi; unpack2(a[i],lo,hi); m=max3(m,lo,hi); }
      }
    }
  }
  float s=m;
  #pragma unroll
  for(int i=0;i<8;++i){ float lo,hi; unpack2(a[i],lo,hi); s+=lo+hi; }
  if(s==1234.5f) sink[0]=s;
}
// K6: scalar FFMA version of the ray mix: 4 rays x 4 spheres in registers, scalar ops + FMNMX3 (per 2 tests)
__global__ void __launch_bounds__(T,1) k6(int iters, float seed, float* sink){
  float dx[8],dy[8],dz[8]; float ox[2],oy[2],oz[2],nc[2]; float m=-1.f;
  #pragma unroll
  for(int r=0;r<8;++r){ dx[r]=seed+0.01f*(threadIdx.x+r); dy[r]=seed+0.02f*r+0.3f; dz[r]=seed+0.5f+0.001f*r; }
  #pragma unroll
  for(int j=0;j<2;++j){ ox[j]=seed+j+1; oy[j]=seed+2*j+0.5f; oz[j]=seed-3.f+j; nc[j]=-1.f-seed-j; }
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int k=0;k<8;++k){
      float q[16];
      #pragma unroll
      for(int j=0;j<2;++j){
        #pragma unroll
        for(int r=0;r<8;++r){ float t=dx[r]*ox[j]; t=fmaf(dy[r],oy[j],t); t=fmaf(dz[r],oz[j],t); q[j*8+r]=fmaf(t,t,nc[j]); }
      }
      #pragma unroll
      for(int i=0;i<8;++i) m=max3(m,q[2*i],q[2*i+1]);
      ox[k&1]=fmaf(ox[k&1],1.0000001f,m*1e-30f);
    }
  }
  if(m==1234.5f) sink[0]=m;
}
template<class K> float timeit(K kern, int ctas, int iters, float* sink){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<ctas,T>>>(8,0.f,sink);
  cudaEventRecord(a); kern<<<ctas,T>>>(iters,0.f,sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); return ms;
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink; cudaMalloc(&sink,16);
  int ctas=sms*4, iters=2000;
  double thr=(double)ctas*T*iters*INNER;
  for(int rep=0;rep<2;++rep){
  float ms;
  ms=timeit(k1,ctas,iters,sink); printf("k1 FFMA  3 fresh regs           : %.1f TFLOP/s\n", thr*16*2/(ms*1e-3)/1e12);
  ms=timeit(k2,ctas,iters,sink); printf("k2 FFMA2 3 fresh pairs          : %.1f TFLOP/s\n", thr*8*4/(ms*1e-3)/1e12);
  ms=timeit(k3,ctas,iters,sink); printf("k3 FFMA2 bcast+pair+acc fresh   : %.1f TFLOP/s\n", thr*8*4/(ms*1e-3)/1e12);
  ms=timeit(k4,ctas,iters,sink); printf("k4 FFMA2 bcast+shared pair+acc  : %.1f TFLOP/s\n", thr*8*4/(ms*1e-3)/1e12);
  ms=timeit(k5,ctas,iters,sink); printf("k5 k4 + FMNMX3 per 4            : %.1f TFLOP/s\n", thr*8*4/(ms*1e-3)/1e12);
  ms=timeit(k6,ctas,iters,sink); printf("k6 scalar ray mix (7 FLOP/test) : %.1f TFLOP/s algorithmic\n", (double)ctas*T*iters*8*16*7/(ms*1e-3)/1e12);
  }
  return 0;
}
