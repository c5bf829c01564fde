// Do concurrent factorisations (csrc/chol.cu: chol_factor with its look-ahead stream, both GEMM modes and the
// diagonal-tile kernel) on different streams change each other's results?  S independent SPD matrices, factored
// one after the other (reference) and then all at once; factors compared bit for bit on the device.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/chol_concurrency.cu -o tools/micro/chol_concurrency
//   chol_concurrency S n [disturb]     disturb = 0 none (round 1: bit-identical)
//                                                1 a stream of COOPERATIVE launches (grid barrier loops, like the forward substitution)
//                                                2 a stream of small-CTA FP64 filler kernels (like the assembly)
//                                                3 both, 4 host threads enqueue the factorisations (one each)
#include <cooperative_groups.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <thread>
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
__global__ void diag_kernel(double* x, int64_t n, double v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i * n + i] = v;
}
__global__ void diff_kernel(const double* a, const double* b, size_t n, unsigned long long* count) {
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) ++c;
  if (c) atomicAdd(count, c);
}

// disturbers: what an evaluation runs besides its factorisation
__global__ void __launch_bounds__(256) coop_disturber_kernel(double* buf, int steps) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  for (int s = 0; s < steps; ++s) {
    buf[blockIdx.x * 256 + threadIdx.x] += 1.0;
    grid.sync();
  }
}
__global__ void __launch_bounds__(128, 8) filler_kernel(double* buf, int iters) {
  double x = buf[(blockIdx.x * 128 + threadIdx.x) & 65535];
  for (int i = 0; i < iters; ++i) x = fma(x, 1.0000001, 1e-9);
  buf[(blockIdx.x * 128 + threadIdx.x) & 65535] = x;
}

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 4;
  const int64_t n = argc > 2 ? atoll(argv[2]) : 6144;
  const int disturb = argc > 3 ? atoi(argv[3]) : 0;
  double* dbuf;
  cudaMalloc(&dbuf, sizeof(double) * 296 * 256);
  cudaMemset(dbuf, 0, sizeof(double) * 296 * 256);
  cudaStream_t sc, sf;
  cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking), cudaStreamCreateWithFlags(&sf, cudaStreamNonBlocking);
  std::vector<double*> A(S), A0(S), Aref(S);
  std::vector<cudaStream_t> st(S);
  std::vector<cocons::CholWorkspace> ws(S);
  unsigned long long* dcount;
  cudaMalloc(&dcount, 8);
  for (int s = 0; s < S; ++s) {
    cudaMalloc(&A[s], sizeof(double) * n * n), cudaMalloc(&A0[s], sizeof(double) * n * n);
    cudaMalloc(&Aref[s], sizeof(double) * n * n);
    cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
    if (cocons::chol_workspace_create(n, &ws[s]) != 0) { printf("workspace failed\n"); return 1; }
    init_kernel<<<1024, 256>>>(A0[s], (size_t)n * n, 17u + s, 0.01);
    diag_kernel<<<(unsigned)((n + 255) / 256), 256>>>(A0[s], n, 40.0 + s);
  }
  cudaDeviceSynchronize();
  for (int s = 0; s < S; ++s) {  // reference: one factorisation at a time
    cudaMemcpy(Aref[s], A0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice);
    cocons::chol_factor(Aref[s], n, n, ws[s], st[s]);
    cudaDeviceSynchronize();
    int info = -1;
    cudaMemcpy(&info, ws[s].info, 4, cudaMemcpyDeviceToHost);
    if (info) printf("reference %d: info=%d\n", s, info);
  }
  for (int pass = 0; pass < 4; ++pass) {
    for (int s = 0; s < S; ++s) cudaMemcpyAsync(A[s], A0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st[s]);
    if (disturb == 1 || disturb == 3)
      for (int r = 0; r < 40; ++r) {
        int steps = 48;
        void* args[] = {(void*)&dbuf, (void*)&steps};
        cudaLaunchCooperativeKernel((void*)coop_disturber_kernel, dim3(96), dim3(256), args, 0, sc);
      }
    if (disturb == 2 || disturb == 3)
      for (int r = 0; r < 40; ++r) filler_kernel<<<4465, 128, 0, sf>>>(dbuf, 20000);
    if (disturb == 3) {
      std::vector<std::thread> th;
      for (int s = 0; s < S; ++s) th.emplace_back([&, s]() { cocons::chol_factor(A[s], n, n, ws[s], st[s]); });
      for (auto& t : th) t.join();
    } else {
      for (int s = 0; s < S; ++s) cocons::chol_factor(A[s], n, n, ws[s], st[s]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long total = 0;
    int bad_info = 0;
    for (int s = 0; s < S; ++s) {
      cudaMemset(dcount, 0, 8);
      diff_kernel<<<1024, 256>>>(A[s], Aref[s], (size_t)n * n, dcount);
      unsigned long long c = 0;
      cudaMemcpy(&c, dcount, 8, cudaMemcpyDeviceToHost);
      total += c;
      int info = -1;
      cudaMemcpy(&info, ws[s].info, 4, cudaMemcpyDeviceToHost);
      bad_info += info != 0;
    }
    printf("CHOL_CONCURRENCY disturb=%d matrices=%d n=%lld pass %d: cuda=%s  entries differing from the serial factor: %llu  not-PD flags: %d\n",
           disturb, S, (long long)n, pass, cudaGetErrorString(e), total, bad_info);
  }
  return 0;
}
