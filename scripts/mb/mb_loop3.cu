// Hot-loop design probe #3 (development, not product): hit detection WITHOUT the sticky FFMA2.
// DET 0: current design (acc = fma2(u, 0, acc): 1 packed op per 2 tests on the FMA pipe)
// DET 1: NaN-propagating 3-input max of |u| on the ALU pipe (FMNMX3), NCH independent chains, checked once per group
// DET 2: integer 3-input max of the sign-stripped bit patterns (VIMNMX3) -- same idea on the integer side
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b){ f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3nan(float a, float b, float c){ float d; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned umax3(unsigned a, unsigned b, unsigned c){ unsigned d; asm("max.u32 %0, %1, %2;" : "=r"(d) : "r"(b), "r"(c)); return max(a, d); }
template<int R, int DET, int NCH, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS,1) loopk(int n_groups, int reps, const float* __restrict__ g, float* sink){
  extern __shared__ __align__(16) float s[];
  for(int i=threadIdx.x;i<n_groups*12;i+=THREADS) s[i]=g[i];
  __syncthreads();
  float ex[R],ey[R],ez[R];
  #pragma unroll
  for(int r=0;r<R;++r){ ex[r]=0.01f*(threadIdx.x+r); ey[r]=0.02f*(r+threadIdx.x)+0.3f; ez[r]=0.5f+0.001f*(r+2*threadIdx.x); }
  const f32x2 Z=pack2(0.f,0.f);
  int flagged=0;
  for(int rep=0;rep<reps;++rep){
    #pragma unroll UNROLL
    for(int p=0;p<n_groups;++p){
      const float4 FX=*(const float4*)(s+p*12), FY=*(const float4*)(s+p*12+4), FZ=*(const float4*)(s+p*12+8);
      f32x2 u[2][R];
      f32x2 acc0=Z, acc1=Z; float m[NCH]; unsigned im[NCH];
      #pragma unroll
      for(int c=0;c<NCH;++c){ m[c]=0.f; im[c]=0u; }
      #pragma unroll
      for(int q=0;q<2;++q){
        const f32x2 GX=q?pack2(FX.z,FX.w):pack2(FX.x,FX.y), GY=q?pack2(FY.z,FY.w):pack2(FY.x,FY.y), GZ=q?pack2(FZ.z,FZ.w):pack2(FZ.x,FZ.y);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=mul2(pack2(ex[r],ex[r]),GX);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=fma2(pack2(ey[r],ey[r]),GY,u[q][r]);
        #pragma unroll
        for(int r=0;r<R;++r) u[q][r]=fma2(pack2(ez[r],ez[r]),GZ,u[q][r]);
        if(DET==0){
          #pragma unroll
          for(int r=0;r<R;++r){ if(r&1) acc1=fma2(u[q][r],Z,acc1); else acc0=fma2(u[q][r],Z,acc0); }
        } else if(DET==1){
          #pragma unroll
          for(int r=0;r<R;++r){ float lo,hi; unpack2(u[q][r],lo,hi); m[(q*R+r)%NCH]=max3nan(m[(q*R+r)%NCH],fabsf(lo),fabsf(hi)); }
        } else {
          #pragma unroll
          for(int r=0;r<R;++r){ float lo,hi; unpack2(u[q][r],lo,hi);
            im[(q*R+r)%NCH]=umax3(im[(q*R+r)%NCH],__float_as_uint(lo)<<1,__float_as_uint(hi)<<1); }
        }
      }
      bool flag;
      if(DET==0){ float lo,hi; unpack2(add2(acc0,acc1),lo,hi); flag=!(lo==hi); }
      else if(DET==1){ float mm=m[0];
        #pragma unroll
        for(int c=1;c<NCH;++c) mm=max3nan(mm,m[c],0.f);
        flag=!(mm<3.0e38f); }
      else { unsigned mm=im[0];
        #pragma unroll
        for(int c=1;c<NCH;++c) mm=max(mm,im[c]);
        flag=mm>=0xff000000u; }
      if(__any_sync(0xffffffffu,flag)){ flagged++;
        #pragma unroll
        for(int q=0;q<2;++q){
        #pragma unroll
        for(int r=0;r<R;++r){ float a,b; unpack2(u[q][r],a,b); if(!(fabsf(a)<=3e38f)) flagged+=r; if(!(fabsf(b)<=3e38f)) flagged+=2*r+q; } } }
    }
  }
  if(flagged==12345) sink[0]=flagged;
}
template<int R,int DET,int NCH,int UNROLL,int THREADS> void run(const char* name,int sms,const float* g,float* sink){
  const int n_spheres=1024, n_groups=n_spheres/4, reps=200;
  auto k=loopk<R,DET,NCH,UNROLL,THREADS>;
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,64*1024);
  cudaFuncAttributes at; cudaFuncGetAttributes(&at,k);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<<<sms,THREADS,n_groups*48>>>(n_groups,2,g,sink);
  cudaEventRecord(a); k<<<sms,THREADS,n_groups*48>>>(n_groups,reps,g,sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double tests=(double)sms*THREADS*reps*(double)n_spheres*R;
  printf("%-40s regs %3d: %5.1f TFLOP/s algorithmic, %.3f cycles/test/lane-slot\n",name,at.numRegs,tests*7/(ms*1e-3)/1e12, ms*1e-3*1.965e9*sms*128/tests);
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,0);
  float* g; cudaMalloc(&g,1<<20); cudaMemset(g,0,1<<20); float* sink; cudaMalloc(&sink,16);
  run<8,0,2,2,512>("sticky FFMA2 (current), 512 thr",sms,g,sink);
  run<8,0,2,2,896>("sticky FFMA2 (current), 896 thr",sms,g,sink);
  run<8,1,1,2,512>("FMNMX3.NaN 1 chain, 512 thr",sms,g,sink);
  run<8,1,2,2,512>("FMNMX3.NaN 2 chains, 512 thr",sms,g,sink);
  run<8,1,4,2,512>("FMNMX3.NaN 4 chains, 512 thr",sms,g,sink);
  run<8,1,4,2,896>("FMNMX3.NaN 4 chains, 896 thr",sms,g,sink);
  run<8,1,8,2,896>("FMNMX3.NaN 8 chains, 896 thr",sms,g,sink);
  run<8,1,4,1,896>("FMNMX3.NaN 4 chains u1, 896 thr",sms,g,sink);
  run<8,2,4,2,512>("integer max 4 chains, 512 thr",sms,g,sink);
  run<8,2,4,2,896>("integer max 4 chains, 896 thr",sms,g,sink);
  return 0;
}
