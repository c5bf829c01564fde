// Timeline probe of potrf_tile_kernel: nvcc -DCOCONS_POTRF_PROBE ... ; prints per-step clock deltas.
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../cocons_b200/csrc/chol.cu"
namespace cocons { void note_launch(int) {} void set_error(const char*, ...) {} }
int main() {
  const int n = 128;
  std::vector<double> A(n * n);
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) A[j * n + i] = (i == j ? 4.0 : 0.0) + 1.0 / (1.0 + abs(i - j));
  double *dA, *dW; int* dinfo;
  cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dW, sizeof(double) * n * n); cudaMalloc(&dinfo, 4);
  cudaMemset(dinfo, 0, 4);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    cocons::potrf_tile_kernel<<<1, 256>>>(dA, n, dW, dinfo, 0);
    cudaDeviceSynchronize();
  }
  std::vector<long long> p(128 * 8);
  cudaMemcpyFromSymbol(p.data(), cocons::g_probe, sizeof(long long) * 128 * 8);
  printf("step: wait->prio_done shfl rsqrt publish(mul+sts) arrive | other-thread: update-span  step-period\n");
  for (int k = 30; k < 40; ++k) {
    long long* q = &p[k * 8];
    printf("k=%3d  owner: prio %5lld shfl %5lld rsqrt %5lld pub %5lld arrive %5lld | t255: update %5lld  period %5lld\n", k,
           q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[7] - q[6], p[(k + 1) * 8 + 6] - q[6]);
  }
  printf("total clocks steps 1..126: %lld\n", p[126 * 8 + 6] - p[1 * 8 + 6]);
  return 0;
}
