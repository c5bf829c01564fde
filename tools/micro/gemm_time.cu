// Timing of the trailing-update GEMM and of a whole factorisation (CUDA events, after warm-up):
//   gemm_time [n_syrk k n_chol reps]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "../../cocons_b200/csrc/chol.cu"
namespace cocons { void note_launch(int) {} void set_error(const char*, ...) {} }
__global__ void init_kernel(double* x, size_t n, unsigned seed, double scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    x[i] = scale * ((double)(h & 0xffff) / 65536.0 - 0.5);
  }
}
__global__ void diag_kernel(double* x, int64_t n, double v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i * n + i] = v;
}
int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 32768, k = argc > 2 ? atoll(argv[2]) : 768;
  const int64_t nc = argc > 3 ? atoll(argv[3]) : 50048;
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  const int64_t big = nc > n ? nc : n;
  double *C, *P, *A0;
  if (cudaMalloc(&C, sizeof(double) * big * big) != cudaSuccess || cudaMalloc(&P, sizeof(double) * n * k) != cudaSuccess ||
      cudaMalloc(&A0, sizeof(double) * nc * nc) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  init_kernel<<<1024, 256>>>(C, (size_t)n * n, 3u, 1.0);
  init_kernel<<<1024, 256>>>(P, (size_t)n * k, 5u, 1e-2);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) cocons::launch_gemm_nt(0, n, n, k, P, n, P, n, C, n, 1, st);
  float best = 1e30f, sum = 0;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, st);
    cocons::launch_gemm_nt(0, n, n, k, P, n, P, n, C, n, 1, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best, sum += ms;
  }
  const double flops = (double)n * (n + 128) / 2 * 2 * k;  // tiles computed: lower triangle incl. diagonal tiles
  printf("GEMM_TIME syrk n=%lld k=%lld: best %.3f ms = %.2f TFLOP/s, mean %.3f ms = %.2f TFLOP/s (%s)\n", (long long)n, (long long)k, best,
         flops / best / 1e9, sum / reps, flops / (sum / reps) / 1e9, cudaGetErrorString(cudaGetLastError()));
  // factorisation
  cocons::CholWorkspace ws;
  if (cocons::chol_workspace_create(nc, &ws) != 0) { printf("workspace failed\n"); return 1; }
  init_kernel<<<1024, 256>>>(A0, (size_t)nc * nc, 17u, 0.01);
  diag_kernel<<<(unsigned)((nc + 255) / 256), 256>>>(A0, nc, 40.0 + 0.002 * nc);
  best = 1e30f, sum = 0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaMemcpyAsync(C, A0, sizeof(double) * nc * nc, cudaMemcpyDeviceToDevice, st);
    cudaEventRecord(e0, st);
    cocons::chol_factor(C, nc, nc, ws, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r) best = ms < best ? ms : best, sum += ms;
  }
  int info = -1;
  cudaMemcpy(&info, ws.info, 4, cudaMemcpyDeviceToHost);
  const double cf = (double)nc * nc * nc / 3;
  printf("GEMM_TIME chol n=%lld: best %.2f ms = %.2f TFLOP/s, mean %.2f ms = %.2f TFLOP/s, info=%d (%s)\n", (long long)nc, best, cf / best / 1e9,
         sum / reps, cf / (sum / reps) / 1e9, info, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
