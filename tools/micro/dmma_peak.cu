// Microbenchmark: DMMA.8x8x4 issue rate per SM sub-partition as a function of resident warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, int iters) {
  double acc[NACC][2];
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps_per_sm : {4, 8, 12, 16, 32}) {
    int threads = warps_per_sm * 32;
    int blocks = 148;
    if (threads > 1024) { threads = 512; blocks = 148 * (warps_per_sm * 32 / 512); }
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<16><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 16 * (double)iters * warps_per_sm * 148;
    printf("warps/SM %2d (per SMSP %d): %.3f ms  %.2f TFLOP/s\n", warps_per_sm, warps_per_sm / 4, ms, flops / ms / 1e9);
  }
  return 0;
}
