// TEST SCAFFOLDING - race check of the shipped kernels on the host (see "optional race check" in cuda_runtime.h).
// Built by tests/host_emul/build.py::build_racecheck with -fsanitize=thread -DCOCONS_EMUL_TSAN on top of driver.cpp
// (every source of libcocons_b200.so, host build) and run by tests/test_host_emul.py / tools/emul_racecheck.sh.
// Exit status 0 and no "ThreadSanitizer" report = no two threads of a block touched the same location without a
// barrier, an mbarrier hand-over or an atomic in between, in any of the kernels exercised below.
//   racecheck            run every kernel family once on a small problem
//   racecheck --racy     a deliberately racy kernel (neighbour exchange through shared memory without a barrier):
//                        ThreadSanitizer MUST report it - the check has teeth
#include "driver.cpp"

#include <cstdio>
#include <random>

namespace {
std::mt19937_64 rng(20261018);
double unif(double lo, double hi) { return lo + (hi - lo) * (double)(rng() >> 11) * (1.0 / 9007199254740992.0); }
std::vector<double> random_matrix(int64_t rows, int64_t cols) {
  std::vector<double> m((size_t)rows * cols);
  for (double& v : m) v = unif(-1, 1);
  return m;
}
// lower triangle + whole diagonal tiles of a well-conditioned SPD matrix, column-major n x n
std::vector<double> spd(int64_t n) {
  const int64_t k = 24;
  std::vector<double> M = random_matrix(n, k), S((size_t)n * n);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      double s = (i == j) ? 2.0 : 0.0;
      for (int64_t c = 0; c < k; ++c) s += M[c * n + i] * M[c * n + j] / k;
      S[j * n + i] = s;
    }
  return S;
}

__global__ void racy_kernel(double* out) {
  __shared__ double cell[64];
  cell[threadIdx.x] = (double)threadIdx.x;
  // missing __syncthreads(): the neighbour's store may not have happened
  out[threadIdx.x] = cell[(threadIdx.x + 1) % 64];
  __syncthreads();
}
}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--racy") {
    std::vector<double> out(64);
    double* o = out.data();
    emul::launch(dim3(1), dim3(64), true, 0, [&] { racy_kernel(o); });
    std::printf("racy kernel done\n");
    return 0;
  }
  // 1. DMMA GEMM with the bulk-copy ring: K / 16 = 7 stages' worth of fills through 4 slots (slots are re-used),
  //    lower-only update and the in-place panel product
  {
    const int64_t M = 256, N = 256, K = 112;
    std::vector<double> A = random_matrix(M, K), C = random_matrix(M, N);
    emu_gemm_nt(0, M, N, K, A.data(), M, A.data(), M, C.data(), M, 1);
    std::vector<double> P = random_matrix(256, 128), W = random_matrix(128, 128);
    emu_gemm_nt(1, 256, 128, 128, P.data(), 256, W.data(), 128, P.data(), 256, 0);
    std::printf("gemm ok\n");
  }
  // 2. blocked Cholesky: tile kernel, panel steps, look-ahead driver
  {
    const int64_t n = 384;
    std::vector<double> S = spd(n), W((size_t)3 * 128 * 128);
    if (emu_chol_factor(n, S.data(), W.data()) != 0) return 2;
    // 3. forward substitution on that factor: dataflow kernel, two-kernel path, cooperative kernel; logdet; Gram
    for (int mode = 0; mode < 3; ++mode) {
      std::vector<double> B = random_matrix(n, 4);  // 2 right-hand sides + scratch
      if (emu_forward_solve(n, S.data(), W.data(), B.data(), 2, mode) != 0) return 3;
    }
    std::vector<double> Y = random_matrix(n, 2), G(4);
    emu_gram(Y.data(), n - 5, n, 2, G.data());
    if (!(emu_logdet(S.data(), n - 5, n) == emu_logdet(S.data(), n - 5, n))) return 4;
    std::printf("cholesky / solves ok\n");
  }
  // 4. pairwise assembly (general Bessel branch): square with symmetrisation, cross-covariance
  {
    const int64_t n = 200, m = 70, p = 3;
    std::vector<double> locs = random_matrix(n, 2), lp = random_matrix(m, 2), X = random_matrix(n, p), Xp = random_matrix(m, p);
    for (int64_t i = 0; i < n; ++i) X[i] = 1.0;
    for (int64_t i = 0; i < m; ++i) Xp[i] = 1.0;
    const double theta6[18] = {0.2, 0.15, 0.1, -1.6, 0.2, -0.15, 0.1, 0.2, -0.1, 0.3, -0.2, 0.1, 0.2, 0.3, -0.2, -4, 0.1, 0.1};
    const double lim[2] = {0.5, 2.5};
    std::vector<double> out((size_t)n * n), cross((size_t)m * n);
    emu_cov_square(0, n, p, locs.data(), X.data(), theta6, lim, out.data());
    emu_cov_pred(n, m, p, locs.data(), lp.data(), X.data(), Xp.data(), theta6, lim, cross.data());
    std::printf("assembly ok\n");
  }
  std::printf("racecheck done\n");
  return 0;
}
