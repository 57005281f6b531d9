// fp64_peaks.cu — measures the FP64 roofline denominators on the box this runs on:
// DFMA (vector pipe), DMMA.8x8x4 (mma.sync f64; the only FP64 tensor shape sm_100a SASS has),
// and both interleaved (do they share a pipe?).  Prints one JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peaks tools/fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

template<int NCHAIN>
__global__ void k_dfma(double* out, int iters, double s){
  double acc[NCHAIN];
  #pragma unroll
  for(int i=0;i<NCHAIN;i++) acc[i]=threadIdx.x*1e-3+i;
  double a=s, b=1.0-s;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NCHAIN;i++) acc[i]=fma(acc[i],a,b);
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<NCHAIN;i++) r+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template<int NACC>
__global__ void k_dmma(double* out, int iters, double s){
  double c0[NACC], c1[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++){c0[i]=0;c1[i]=0;}
  double a=s*(threadIdx.x&31), b=1e-3*s;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c0[i],c1[i],a,b);
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) r+=c0[i]+c1[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}

// NACC dmma + NF dfma per iteration
template<int NACC,int NF>
__global__ void k_mix(double* out, int iters, double s){
  double c0[NACC], c1[NACC], acc[NF];
  #pragma unroll
  for(int i=0;i<NACC;i++){c0[i]=0;c1[i]=0;}
  #pragma unroll
  for(int i=0;i<NF;i++) acc[i]=threadIdx.x*1e-3+i;
  double a=s*(threadIdx.x&31), b=1e-3*s, fa=s, fb=1.0-s;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++){
      dmma884(c0[i],c1[i],a,b);
      #pragma unroll
      for(int j=0;j<NF/NACC;j++) acc[i*(NF/NACC)+j]=fma(acc[i*(NF/NACC)+j],fa,fb);
    }
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) r+=c0[i]+c1[i];
  #pragma unroll
  for(int i=0;i<NF;i++) r+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}

// Matern52 entry cost: r2 -> sqrt -> exp -> poly  (fp64 special-function budget)
__global__ void k_mat52(double* out, int iters, double s){
  double r2=s*(threadIdx.x+1)*1e-3, acc=0;
  const double s5=2.23606797749979;
  for(int it=0;it<iters;it++){
    double r=sqrt(r2);
    double e=exp(-s5*r);
    acc+= (1.0+s5*r+(5.0/3.0)*r2)*e;
    r2+=1e-6;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

template<class F> float timeit(F f,int reps){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best) best=ms;
  }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, (size_t)sms*8*1024*sizeof(double)));
  int iters=4096;
  printf("{\"gpu\":\"%s\",\"sms\":%d",p.name,sms);
  // DFMA: threads per block sweep, 2 blocks per SM
  for(int tpb: {256,512,1024}){
    int blocks=sms*(2048/tpb);
    float ms=timeit([&]{k_dfma<8><<<blocks,tpb>>>(out,iters,0.5);},5);
    double fl=2.0*8*iters*(double)blocks*tpb;
    printf(",\"dfma_tpb%d_tflops\":%.2f",tpb,fl/ms*1e-9);
  }
  for(int tpb: {128,256,512,1024}){
    int blocks=sms*(tpb>=512?2048/tpb:4);
    float ms=timeit([&]{k_dmma<8><<<blocks,tpb>>>(out,iters,0.5);},5);
    double fl=2.0*256*8*iters*(double)blocks*(tpb/32);
    printf(",\"dmma_acc8_tpb%d_tflops\":%.2f",tpb,fl/ms*1e-9);
  }
  {
    int tpb=256, blocks=sms*4;
    float ms=timeit([&]{k_dmma<2><<<blocks,tpb>>>(out,iters,0.5);},5);
    printf(",\"dmma_acc2_tpb256_tflops\":%.2f",2.0*256*2*iters*(double)blocks*(tpb/32)/ms*1e-9);
    ms=timeit([&]{k_dmma<4><<<blocks,tpb>>>(out,iters,0.5);},5);
    printf(",\"dmma_acc4_tpb256_tflops\":%.2f",2.0*256*4*iters*(double)blocks*(tpb/32)/ms*1e-9);
    ms=timeit([&]{k_dmma<16><<<blocks,tpb>>>(out,iters,0.5);},5);
    printf(",\"dmma_acc16_tpb256_tflops\":%.2f",2.0*256*16*iters*(double)blocks*(tpb/32)/ms*1e-9);
  }
  {
    int tpb=256, blocks=sms*4;
    // 4 dmma (1024 FMA) + 8 dfma warp-instr (256 FMA) per iter
    float ms=timeit([&]{k_mix<4,8><<<blocks,tpb>>>(out,iters,0.5);},5);
    double fl=2.0*(256*4+32*8)*iters*(double)blocks*(tpb/32);
    printf(",\"mix_4dmma_8dfma_tflops\":%.2f",fl/ms*1e-9);
    ms=timeit([&]{k_mix<4,32><<<blocks,tpb>>>(out,iters,0.5);},5);
    fl=2.0*(256*4+32*32)*iters*(double)blocks*(tpb/32);
    printf(",\"mix_4dmma_32dfma_tflops\":%.2f",fl/ms*1e-9);
  }
  {
    int tpb=256, blocks=sms*8;
    float ms=timeit([&]{k_mat52<<<blocks,tpb>>>(out,iters,0.5);},5);
    printf(",\"mat52_gentries_per_s\":%.2f",(double)iters*blocks*tpb/ms*1e-6);
  }
  printf("}\n");
  return 0;
}
