// Do concurrent instances of the DMMA GEMM (csrc/chol.cu) from different streams change each other's
// results?  S independent problems C_s -= P_s P_s^T (lower triangle, K = 512), first one after the other
// (reference), then all at once on S streams, each stream running a chain of R dependent updates; the
// results are compared bit for bit on the device.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/gemm_concurrency.cu -o tools/micro/gemm_concurrency
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cocons_b200/csrc/chol.cu"
namespace cocons { void note_launch(int) {} void set_error(const char*, ...) {} }

__global__ void init_kernel(double* x, size_t n, unsigned seed, double scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    x[i] = scale * ((double)(h & 0xffff) / 65536.0 - 0.5);
  }
}
__global__ void diff_kernel(const double* a, const double* b, size_t n, unsigned long long* count) {
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) ++c;
  if (c) atomicAdd(count, c);
}

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 4, R = argc > 2 ? atoi(argv[2]) : 6;
  const int64_t n = argc > 3 ? atoll(argv[3]) : 8192, k = 512;
  std::vector<double*> C(S), C0(S), Cref(S), P(S);
  std::vector<cudaStream_t> st(S);
  unsigned long long* dcount;
  cudaMalloc(&dcount, 8);
  for (int s = 0; s < S; ++s) {
    cudaMalloc(&C[s], sizeof(double) * n * n), cudaMalloc(&C0[s], sizeof(double) * n * n);
    cudaMalloc(&Cref[s], sizeof(double) * n * n), cudaMalloc(&P[s], sizeof(double) * n * k);
    cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
    init_kernel<<<1024, 256>>>(C0[s], (size_t)n * n, 17u + s, 1.0);
    init_kernel<<<1024, 256>>>(P[s], (size_t)n * k, 91u + s, 1e-2);
  }
  cudaDeviceSynchronize();
  // reference: one problem at a time
  for (int s = 0; s < S; ++s) {
    cudaMemcpy(Cref[s], C0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice);
    for (int r = 0; r < R; ++r) cocons::launch_gemm_nt(0, n, n, k, P[s], n, P[s], n, Cref[s], n, 1, st[0]);
    cudaStreamSynchronize(st[0]);
  }
  for (int pass = 0; pass < 3; ++pass) {
    for (int s = 0; s < S; ++s) cudaMemcpyAsync(C[s], C0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st[s]);
    for (int r = 0; r < R; ++r)  // round-robin submission: every stream holds a chain of R dependent launches
      for (int s = 0; s < S; ++s) cocons::launch_gemm_nt(0, n, n, k, P[s], n, P[s], n, C[s], n, 1, st[s]);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long total = 0;
    for (int s = 0; s < S; ++s) {
      cudaMemset(dcount, 0, 8);
      diff_kernel<<<1024, 256>>>(C[s], Cref[s], (size_t)n * n, dcount);
      unsigned long long c = 0;
      cudaMemcpy(&c, dcount, 8, cudaMemcpyDeviceToHost);
      total += c;
    }
    printf("GEMM_CONCURRENCY streams=%d chain=%d n=%lld pass %d: cuda=%s  entries differing from the serial result: %llu\n", S,
           R, (long long)n, pass, cudaGetErrorString(e), total);
  }
  return 0;
}
