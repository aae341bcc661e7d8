// Micro-benchmarks that size the Baum-Welch kernel design on B200 (results in DESIGN.md):
// FP64 FMA issue rate, shared fp64 atomicAdd (CAS loop), global fp64 RED, MATCH.ANY.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__global__ void k_dfma(double* out, int iters){
  double a0=threadIdx.x*1e-3,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7; double b=1.0000001,c=1e-9;
  for(int i=0;i<iters;++i){a0=fma(a0,b,c);a1=fma(a1,b,c);a2=fma(a2,b,c);a3=fma(a3,b,c);a4=fma(a4,b,c);a5=fma(a5,b,c);a6=fma(a6,b,c);a7=fma(a7,b,c);}
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
__global__ void k_smem_atomic(double* out, int iters, int nbins){
  extern __shared__ double s[];
  for(int i=threadIdx.x;i<nbins;i+=blockDim.x) s[i]=0; __syncthreads();
  unsigned x = threadIdx.x*2654435761u + blockIdx.x*40503u;
  for(int i=0;i<iters;++i){ x = x*1664525u+1013904223u; atomicAdd(&s[(x>>8)%nbins], 1.0); }
  __syncthreads(); if(threadIdx.x==0) out[blockIdx.x]=s[0];
}
__global__ void k_gmem_red(double* acc, int iters, int nbins){
  unsigned x = threadIdx.x*2654435761u + blockIdx.x*40503u;
  for(int i=0;i<iters;++i){ x = x*1664525u+1013904223u; atomicAdd(&acc[(x>>8)%nbins], 1.0); }
}
__global__ void k_match(unsigned* out, int iters, int nbins){
  unsigned x = threadIdx.x*2654435761u + blockIdx.x*40503u; unsigned accm=0;
  for(int i=0;i<iters;++i){ x = x*1664525u+1013904223u; accm += __match_any_sync(0xffffffffu,(x>>8)%nbins); }
  out[blockIdx.x*blockDim.x+threadIdx.x]=accm;
}
// warp-private non-atomic RMW with match-based conflict rounds (the bwd4 count update)
__global__ void k_match_rmw(double* out, int iters, int nbins){
  extern __shared__ double s[];
  int warp=threadIdx.x>>5, lane=threadIdx.x&31; double* cw = s + (size_t)warp*nbins*4;
  for(int i=threadIdx.x;i<nbins*4*(blockDim.x/32);i+=blockDim.x) s[i]=0; __syncthreads();
  unsigned x = threadIdx.x*2654435761u + blockIdx.x*40503u;
  for(int i=0;i<iters;++i){ x = x*1664525u+1013904223u; unsigned sym=(x>>8)%nbins;
    unsigned peers=__match_any_sync(0xffffffffu,sym); int rank=__popc(peers&((1u<<lane)-1)); int mr=__reduce_max_sync(0xffffffffu,rank);
    double2* row=(double2*)(cw+sym*4);
    for(int r=0;r<=mr;++r){ if(rank==r){double2 a=row[0],b=row[1]; a.x+=1;a.y+=1;b.x+=1;b.y+=1; row[0]=a; row[1]=b;} __syncwarp(); } }
  __syncthreads(); if(threadIdx.x==0) out[blockIdx.x]=s[0];
}
int main(){
  int dev=0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev)); int sm=p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n",p.name,sm,p.clockRate);
  double* d; CK(cudaMalloc(&d, 1<<26)); CK(cudaMemset(d,0,1<<26));
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  // DFMA
  for(int rep=0;rep<2;++rep){ int iters=20000; cudaEventRecord(a); k_dfma<<<sm*8,256>>>(d,iters); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
    double fl=(double)sm*8*256*iters*8*2; if(rep) printf("dfma: %.2f TFLOP/s fp64 (%.1f FMA/clk/SM at %d MHz nominal)\n", fl/ms/1e9, fl/2/(ms*1e-3)/sm/(p.clockRate*1e3), p.clockRate/1000);}
  int binsv[3]={256*4, 1024*16, 64};
  for(int bi=0;bi<3;++bi){ int nb=binsv[bi]; int iters=4000;
    CK(cudaFuncSetAttribute(k_smem_atomic,cudaFuncAttributeMaxDynamicSharedMemorySize,nb*8));
    for(int rep=0;rep<2;++rep){cudaEventRecord(a); k_smem_atomic<<<sm*2,256,nb*8>>>(d,iters,nb); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      if(rep) printf("smem fp64 atomicAdd (CAS) bins=%d: %.2f G atomics/s chip, %.3f per clk per SM\n", nb, (double)sm*2*256*iters/ms/1e6, (double)sm*2*256*iters/(ms*1e-3)/sm/(p.clockRate*1e3));}
    for(int rep=0;rep<2;++rep){cudaEventRecord(a); k_gmem_red<<<sm*8,256>>>(d,iters,nb*10); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      if(rep) printf("gmem fp64 RED bins=%d: %.2f G atomics/s chip, %.3f per clk per SM\n", nb*10, (double)sm*8*256*iters/ms/1e6, (double)sm*8*256*iters/(ms*1e-3)/sm/(p.clockRate*1e3));}
  }
  for(int nb=48; nb<=256; nb+= 208){ int iters=4000;
    for(int rep=0;rep<2;++rep){cudaEventRecord(a); k_match<<<sm*8,256>>>((unsigned*)d,iters,nb); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      if(rep) printf("match.any bins=%d: %.3f warp-instr per clk per SM\n", nb, (double)sm*8*8*iters/(ms*1e-3)/sm/(p.clockRate*1e3));}
    int smem=nb*4*8*4; CK(cudaFuncSetAttribute(k_match_rmw,cudaFuncAttributeMaxDynamicSharedMemorySize,smem));
    for(int occ=1; occ<=4; occ*=2) for(int rep=0;rep<2;++rep){cudaEventRecord(a); k_match_rmw<<<sm*occ,128,smem>>>(d,iters,nb); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      if(rep) printf("match+rmw(32B) bins=%d ctas/SM=%d: %.3f lane-updates per clk per SM\n", nb, occ, (double)sm*occ*128*iters/(ms*1e-3)/sm/(p.clockRate*1e3));}
  }
  return 0;
}
