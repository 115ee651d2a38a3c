// Microbenchmark: FP64 pipe rates on B200 (sm_100a). Measures DMMA.8x8x4 and DFMA issue
// throughput per SM as a function of resident warps, to fix the roofline denominator
// for the DMMA GEMM kernels (see DESIGN.md "FP64 peak").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void dmma_loop(double* out, int iters, double av, double bv){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  double a=av+threadIdx.x, b=bv;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++){
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void dfma_loop(double* out, int iters, double av, double bv){
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  double a=av, b=bv+threadIdx.x*1e-9;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(a,c[i],b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

int main(){
  int dev=0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount;
  int clk_khz=0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  printf("device %s SMs %d clockRate %d kHz\n", p.name, sms, clk_khz);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*1024*4));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters=20000;
  for(int warps : {1,2,4,8,16,32}){
    for(int rep=0; rep<2; ++rep){
      dmma_loop<16><<<sms, warps*32>>>(out, iters, 1.0, 1.0);
    }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    dmma_loop<16><<<sms, warps*32>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*256*16*(double)iters*warps*sms;
    printf("DMMA.8x8x4 NACC=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops/ms*1e-9);
  }
  for(int warps : {4,8}){
    cudaEventRecord(e0);
    dmma_loop<4><<<sms, warps*32>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*256*4*(double)iters*warps*sms;
    printf("DMMA.8x8x4 NACC=4  warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops/ms*1e-9);
  }
  for(int warps : {4,8,16,32}){
    dfma_loop<16><<<sms, warps*32>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    dfma_loop<16><<<sms, warps*32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*32*16*(double)iters*warps*sms;
    printf("DFMA NACC=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops/ms*1e-9);
  }
  // sustained DMMA for ~3 s to see the power-capped rate
  {
    int warps=8; float total=0; int n=0; double flops = 2.0*256*16*(double)(iters*10)*warps*sms;
    float best=1e30f, last=0;
    while(total<3000.f){
      cudaEventRecord(e0);
      dmma_loop<16><<<sms, warps*32>>>(out, iters*10, 1.0, 1.0);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1); total+=ms; n++; if(ms<best)best=ms; last=ms;
    }
    printf("DMMA sustained 3s: best %7.2f TFLOP/s, last %7.2f TFLOP/s (%d launches)\n", flops/best*1e-9, flops/last*1e-9, n);
  }
  return 0;
}
