// TEST SCAFFOLDING - race check of the shipped kernels on the host (see "optional race check" in cuda_runtime.h).
// Built by tests/host_emul/build.py::build_racecheck with -fsanitize=thread -DCOCONS_EMUL_TSAN on top of driver.cpp
// (every source of libcocons_b200.so, host build) and run by tests/test_host_emul.py / tools/emul_racecheck.sh.
// Exit status 0 and no "ThreadSanitizer" report = no two threads of a block touched the same location without a
// barrier, an mbarrier hand-over or an atomic in between, in any of the kernels exercised below.
//   racecheck            run every kernel family once on a small problem, then the C ABI and the block-cyclic path
//   racecheck --quick    smaller problems, without the block-cyclic path (what the CPU test suite runs)
//   racecheck --gemm     the DMMA GEMM section only
//   racecheck --racy     two deliberately racy kernels (a neighbour exchange through shared memory without a barrier;
//                        two concurrent blocks storing to one global location): ThreadSanitizer MUST report both
#include "driver.cpp"

#include <cstdio>
#include <random>
#include <thread>

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
// two blocks of one launch store to the same global location
__global__ void racy_blocks_kernel(double* out) {
  __shared__ double s;
  if (threadIdx.x == 0) s = (double)blockIdx.x;
  __syncthreads();
  if (threadIdx.x == 1) out[0] = s;
}
}  // namespace

int main(int argc, char** argv) {
  const bool quick = argc > 1 && std::string(argv[1]) == "--quick";
  if (argc > 1 && std::string(argv[1]) == "--racy") {
    std::vector<double> out(64);
    double* o = out.data();
    emul::launch(dim3(1), dim3(64), true, 0, [&] { racy_kernel(o); });
    emu_set_concurrent_blocks(2);
    emul::launch(dim3(2), dim3(32), true, 0, [&] { racy_blocks_kernel(o); });
    std::printf("racy kernels done\n");
    return 0;
  }
  // Sections 1, 2 (factorisation) and 4 run FOUR BLOCKS AT A TIME on OS threads: besides the hazards inside a block,
  // ThreadSanitizer then sees every pair of blocks of a launch that touch the same global location (two CTAs writing
  // one tile, a CTA reading what another one of the same launch writes).  The dataflow forward substitution is the one
  // kernel whose blocks DO talk through global memory - by design, partly through plain stores polled with relaxed
  // loads, which a race detector reports by definition; it keeps one block at a time here and is run with concurrent
  // blocks, for its results, in tests/test_host_emul.py.
  // (the mutation self-test keeps one block at a time: it looks for ONE specific pair of accesses inside a block, and
  // ThreadSanitizer remembers only the last few accesses of a location)
  // Each of those sections runs twice - one block at a time, then four - because the detector keeps only the last
  // few accesses of a location: the serial pass is the sharper one for hazards inside a block.
  const bool mutation = std::getenv("COCONS_EMUL_DROP_HANDBACK") != nullptr;
  // 1. DMMA GEMM with the bulk-copy ring: K / 16 = 7 stages' worth of fills through 4 slots (slots are re-used),
  //    lower-only update and the in-place panel product
  for (int many : {1, 4}) {
    if (mutation && many > 1) break;
    emu_set_concurrent_blocks(many);
    const int64_t M = 256, N = 256, K = 112;
    std::vector<double> A = random_matrix(M, K), C = random_matrix(M, N);
    emu_gemm_nt(0, M, N, K, A.data(), M, A.data(), M, C.data(), M, 1);
    std::vector<double> P = random_matrix(256, 128), W = random_matrix(128, 128);
    emu_gemm_nt(1, 256, 128, 128, P.data(), 256, W.data(), 128, P.data(), 256, 0);
    std::printf("gemm ok (%d block(s) at a time)\n", many);
  }
  if (argc > 1 && std::string(argv[1]) == "--gemm") return 0;
  // 2. blocked Cholesky: tile kernel, panel steps, look-ahead driver (quick: the concurrent pass only)
  for (int many : {1, 4}) {
    if (quick && many == 1) continue;
    emu_set_concurrent_blocks(many);
    const int64_t n = quick ? 256 : 384;
    std::vector<double> S = spd(n), W((size_t)3 * 128 * 128);
    if (emu_chol_factor(n, S.data(), W.data()) != 0) return 2;
    emu_set_concurrent_blocks(1);
    if (many == 1) continue;
    // 3. forward substitution on that factor: dataflow kernel, two-kernel path, cooperative kernel; logdet; Gram
    for (int mode = 0; mode < 3; ++mode) {
      std::vector<double> B = random_matrix(n, 4);  // 2 right-hand sides + scratch
      if (emu_forward_solve(n, S.data(), W.data(), B.data(), 2, mode) != 0) return 3;
    }
    std::vector<double> Y = random_matrix(n, 2), G(4);
    emu_gram(Y.data(), n - 5, n, 2, G.data());
    if (!(emu_logdet(S.data(), n - 5, n) == emu_logdet(S.data(), n - 5, n))) return 4;
    {  // the column-sweep variant of the tile kernel (K3, COCONS_POTRF=1)
      std::vector<double> T = spd(128), Wt((size_t)128 * 128);
      if (emu_potrf_tile_sweep(T.data(), 128, Wt.data(), 0) != 0) return 4;
    }
    std::printf("cholesky / solves ok\n");
  }
  // 4. pairwise assembly (general Bessel branch): square with symmetrisation, cross-covariance
  for (int many : {1, 4}) {
    emu_set_concurrent_blocks(many);
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
  emu_set_concurrent_blocks(1);
  // 5. the C ABI end to end on a resident context: REML objective (design columns + z as right-hand sides, Gram
  //    algebra), kept factor -> cocoPredict reductions, marginal and conditional draws, the tapered objective
  {
    const int64_t n = quick ? 130 : 200, m = 40, p = 3;
    std::vector<double> locs = random_matrix(n, 2), lp = random_matrix(m, 2), X = random_matrix(n, p), Xp = random_matrix(m, p);
    std::vector<double> z = random_matrix(n, 1), eps = random_matrix(n, 2), epsm = random_matrix(m, 2);
    for (int64_t i = 0; i < n; ++i) X[i] = 1.0;
    for (int64_t i = 0; i < m; ++i) Xp[i] = 1.0;
    const double theta6[18] = {0.2, 0.15, 0.1, -1.6, 0.2, -0.15, 0.1, 0.2, -0.1, 0.3, -0.2, 0.1, 0.2, 0.3, -0.2, -4, 0.1, 0.1};
    const double lim[2] = {0.5, 2.5}, mean[3] = {0.1, 0.3, -0.2};
    cocons_ctx* c = nullptr;
    if (cocons_ctx_create(0, n, p, 1, locs.data(), X.data(), z.data(), nullptr, &c) != 0) return 5;
    double logdet = 0, quad[1], ldw = 0;
    int rank = 0;
    if (cocons_n2ll(c, COCONS_REML, theta6, lim, nullptr, &logdet, quad, &ldw, &rank) != 0) return 6;
    if (cocons_n2ll(c, COCONS_ML, theta6, lim, mean, &logdet, quad, &ldw, &rank) != 0) return 7;
    if (cocons_factor(c, COCONS_PAR_DIFF, theta6, lim) != 0) return 8;
    std::vector<double> sto(m), expl(m), draws((size_t)n * 2), cond((size_t)m * 2);
    if (cocons_predict(c, m, lp.data(), Xp.data(), z.data(), sto.data(), expl.data()) != 0) return 9;
    if (cocons_sim(c, 2, eps.data(), draws.data()) != 0) return 10;
    if (cocons_sim_cond(c, m, lp.data(), Xp.data(), 2, epsm.data(), cond.data()) != 0) return 11;
    // tapered objective on a banded pattern (diagonal + two neighbours each side), taper 1 on the diagonal
    std::vector<int32_t> col, row(1, 1);
    std::vector<double> tap;
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t j = std::max<int64_t>(0, i - 2); j <= std::min<int64_t>(n - 1, i + 2); ++j)
        col.push_back((int32_t)j + 1), tap.push_back(i == j ? 1.0 : 0.05);
      row.push_back((int32_t)col.size() + 1);
    }
    if (cocons_ctx_set_taper(c, col.data(), row.data(), tap.data(), (int64_t)col.size()) != 0) return 12;
    if (cocons_n2ll_taper(c, theta6, lim, mean, &logdet, quad) != 0) return 13;
    cocons_ctx_destroy(c);
    std::printf("C ABI ok\n");
  }
  if (quick) {
    std::printf("racecheck done\n");
    return 0;
  }
  // 6. the block-cyclic path on one rank (csrc/dist.cu): cyclic-slab assembly, panel factorisation, row packing,
  //    trailing updates, blocked solve with the accumulator update, local reductions.  n_pad = 640: two panels
  {
    const int64_t n = 600, p = 3;
    std::vector<double> locs = random_matrix(n, 2), X = random_matrix(n, p), z = random_matrix(n, 1);
    for (int64_t i = 0; i < n; ++i) X[i] = 1.0;
    const double theta6[18] = {0.2, 0.15, 0.1, -1.6, 0.2, -0.15, 0.1, 0.2, -0.1, 0.3, -0.2, 0.1, 0.2, 0.3, -0.2, -4, 0.1, 0.1};
    const double lim[2] = {0.5, 2.5}, mean[3] = {0.1, 0.3, -0.2};
    cocons_dist* d = nullptr;
    if (cocons_dist_create(0, 0, 1, n, p, 1, locs.data(), X.data(), z.data(), nullptr, &d) != 0) return 14;
    const int64_t np = cocons_dist_npanels(d), n_pad = cocons_dist_npad(d);
    if (cocons_dist_assemble(d, theta6, lim, mean) != 0) return 15;
    std::vector<double> buf((size_t)n_pad * 512);
    for (int64_t K = 0; K < np; ++K) {
      if (cocons_dist_factor_panel(d, K, 0) != 0) return 16;
      if (cocons_dist_panel_elems(d, K) > 0) {
        if (cocons_dist_pack_panel(d, K, buf.data(), 0) != 0) return 17;
        if (cocons_dist_update(d, K, buf.data(), K + 1, np) != 0) return 18;
      }
    }
    int nr = 0;
    std::vector<double> rhs((size_t)np * 512), acc((size_t)np * 512, 0.0), Y((size_t)n_pad), out2(2), gram(1);
    if (cocons_dist_fill_rhs(d, COCONS_ML, rhs.data(), &nr) != 0 || nr != 1) return 19;
    for (int64_t K = 0; K < np; ++K)
      if (cocons_dist_solve_block(d, K, rhs.data() + K * 512, acc.data() + K * 512, acc.data(), Y.data(), nr) != 0) return 20;
    if (cocons_dist_reduce_local(d, Y.data(), nr, out2.data(), gram.data()) != 0 || out2[1] != 0.0) return 21;
    cocons_dist_destroy(d);
    std::printf("block-cyclic path ok\n");
  }
  // 7. several host threads, each driving its own context (DenseLikelihoodPool; INTEGRATION.md "several host threads
  //    may drive several contexts on one device"): the library's host-side state - error buffer, launch counter,
  //    per-context streams / workspaces / staging buffers - must not be shared between them without synchronisation
  {
    const int64_t n = 130, p = 3;
    std::vector<double> locs = random_matrix(n, 2), X = random_matrix(n, p), z = random_matrix(n, 1);
    for (int64_t i = 0; i < n; ++i) X[i] = 1.0;
    const double theta6[18] = {0.2, 0.15, 0.1, -1.6, 0.2, -0.15, 0.1, 0.2, -0.1, 0.3, -0.2, 0.1, 0.2, 0.3, -0.2, -4, 0.1, 0.1};
    const double lim[2] = {0.5, 2.5}, mean[3] = {0.1, 0.3, -0.2};
    double values[3] = {0, 0, 0};
    int status[3] = {0, 0, 0};
    std::vector<std::thread> pool;
    for (int t = 0; t < 3; ++t)
      pool.emplace_back([&, t] {
        cocons_ctx* c = nullptr;
        double quad[1], ldw = 0;
        int rank = 0;
        status[t] = cocons_ctx_create(0, n, p, 1, locs.data(), X.data(), z.data(), nullptr, &c);
        for (int rep = 0; rep < 2 && status[t] == 0; ++rep)
          status[t] = cocons_n2ll(c, COCONS_ML, theta6, lim, mean, &values[t], quad, &ldw, &rank);
        if (c) cocons_ctx_destroy(c);
      });
    for (auto& th : pool) th.join();
    for (int t = 0; t < 3; ++t)
      if (status[t] != 0 || values[t] != values[0]) return 22;
    std::printf("host threads ok\n");
  }
  std::printf("racecheck done\n");
  return 0;
}
