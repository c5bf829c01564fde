// TEST SCAFFOLDING - the host build of libcocons_b200.so's sources (see cuda_runtime.h in this directory).
// *_INC are cocons_b200/csrc/{assembly,taper,solve,chol,dist,capi}.cu after the mechanical rewrites of
// tests/host_emul/build.py (launch chevrons, dynamic shared memory, inline-PTX helpers); everything else in them is
// compiled as it ships - including the C ABI itself (capi.cu, dist.cu), which this library therefore exports with the
// emulated kernels behind it (fixture `product_on_host`, tests/_mp_worker.py).  The emu_* functions below give the
// tests direct access to single launch helpers, repeating only the few lines of orchestration the library's entry
// points put around them:
//   emu_cov_square    cov_square()            capi.cu  (cocons_cov_rns / cocons_cov_rns_classic)
//   emu_cov_pred      cocons_cov_rns_pred()   capi.cu
//   emu_ctx_lower     assemble_and_factor()   capi.cu  (Morton-ordered, padded, lower triangle only)
//   emu_dist_slabs    cocons_dist_assemble()  dist.cu  (one rank's column panels, both launch modes)
//   emu_taper_*       taper_entries_host() / the TS_LOWER sink of assemble_and_factor(taper = true)
//   emu_forward_solve forward_solve_ws() / forward_solve()  solve.cu (K6b dataflow kernel through the library's own
//                     launcher; K6 cooperative kernel on a grid of one block; the two-kernel-per-step path)
//   emu_logdet / emu_gram                     launch_logdet() / launch_gram()
//   emu_chol_factor / emu_potrf_tile / emu_gemm_nt   chol_factor() / launch_potrf_tile() / launch_gemm_nt()  chol.cu
#include "cuda_runtime.h"  // the emulation shim (found first through -I tests/host_emul)

#include ASSEMBLY_INC
#include TAPER_INC
#include SOLVE_INC
#include CHOL_INC
#include DIST_INC  // defines the cocons_dist_* C ABI itself: this library exports it, backed by the emulated kernels
#include CAPI_INC  // ... and the rest of the C ABI of include/cocons_b200.h (with set_error / note_launch)



using namespace cocons;

namespace {
// the exhaustive check of tools/micro/tile_decode_check.cu: total_tiles / tile_decode<W> must be a bijection onto
// the expected tile set of an (ni x njc) update, full or lower-trapezoid
template <int W>
long tile_decode_errors(int ni, int njc, int lower) {
  const int64_t total = total_tiles<W>(ni, njc, lower);
  std::vector<unsigned char> seen((size_t)ni * njc, 0);
  long bad = 0;
  int64_t expect = 0;
  for (int bi = 0; bi < ni; ++bi)
    for (int c = 0; c < njc; ++c)
      if (!lower || bi >= c / W) ++expect;
  if (expect != total) return 1 + std::llabs(expect - total);
  for (int64_t t = 0; t < total; ++t) {
    int bi = -1, bj = -1;
    tile_decode<W>(t, ni, njc, lower, bi, bj);
    if (bi < 0 || bi >= ni || bj < 0 || bj >= njc || (lower && bi < bj / W) || seen[(size_t)bi * njc + bj]++) ++bad;
  }
  return bad;
}
}  // namespace

namespace {
template <int NR>
void run_variant(int mode, const double* L, int64_t ld, const double* winv, double* B, double* Y, int64_t ldb,
                 int64_t nt, int nr) {
  if (mode == 1)
    launch_fwd_steps<NR>(L, ld, winv, B, Y, ldb, nt, nr, nullptr);
  else
    emul::launch(dim3(1), dim3(256), true, 0, [&] { fwd_solve_coop_kernel<NR>(L, ld, winv, B, Y, ldb, nt, nr); });
}
}  // namespace

extern "C" {

const char* emu_last_error() { return cocons_last_error(); }
long emu_launches() { return emul::launches; }
// how many blocks of a launch run at the same time (see emul::concurrent_blocks); returns the previous value
int emu_set_concurrent_blocks(int g) {
  const int old = emul::concurrent_blocks;
  emul::concurrent_blocks = g < 1 ? 1 : (g > 64 ? 64 : g);
  return old;
}
long emu_barrier_launches() { return emul::barrier_launches; }

int emu_cov_square(int par, int64_t n, int64_t p, const double* locs, const double* X, const double* theta6,
                   const double* limits, double* out) {
  const double lim[2] = {limits ? limits[0] : 0.0, limits ? limits[1] : 0.0};
  double nu_fixed;
  const int mode = smooth_mode_for(par, (int)p, theta6, lim, &nu_fixed);
  const double global_range = 1 / std::exp(-2 * theta6[p]);
  std::vector<double> S((size_t)SF_COUNT * n);
  SiteTable T{S.data(), n, nullptr};
  launch_site_stage(n, n, (int)p, X, n, locs, n, theta6, lim[0], lim[1], mode, T, nullptr);
  launch_assemble_lower(n, n, T, global_range, nu_fixed, mode, out, n, nullptr);
  launch_symmetrize(n, out, n, nullptr);
  return mode;
}

int emu_cov_pred(int64_t n, int64_t m, int64_t p, const double* locs, const double* locs_pred, const double* X,
                 const double* X_pred, const double* theta6, const double* limits, double* out) {
  const double global_range = 1 / std::exp(-2 * theta6[p]);
  std::vector<double> S((size_t)SF_COUNT * n), Sp((size_t)SF_COUNT * m);
  SiteTable T{S.data(), n, nullptr}, P{Sp.data(), m, nullptr};
  launch_site_stage(n, n, (int)p, X, n, locs, n, theta6, limits[0], limits[1], SM_GENERAL, T, nullptr);
  launch_site_stage(m, m, (int)p, X_pred, m, locs_pred, m, theta6, limits[0], limits[1], SM_GENERAL, P, nullptr);
  launch_assemble_cross(m, n, P, T, global_range, out, m, nullptr);
  return 0;
}

void emu_morton(int64_t n, const double* locs, int64_t* perm) { morton_order(n, locs, perm); }

// the context's matrix: sites already in the context's order (locs / X with leading dimension n_pad, orig = the
// caller index of each), n_pad x n_pad output of which only the lower triangle (and whole diagonal tiles) is written
int emu_ctx_lower(int par, int64_t n, int64_t n_pad, int64_t p, const double* locs, const double* X,
                  const double* theta6, const double* limits, const int* orig, double* A) {
  const double lim[2] = {limits ? limits[0] : 0.0, limits ? limits[1] : 0.0};
  double nu_fixed;
  const int mode = smooth_mode_for(par, (int)p, theta6, lim, &nu_fixed);
  const double global_range = 1 / std::exp(-2 * theta6[p]);
  std::vector<double> S((size_t)SF_COUNT * n_pad);
  SiteTable T{S.data(), n_pad, orig};
  launch_site_stage(n, n_pad, (int)p, X, n_pad, locs, n_pad, theta6, lim[0], lim[1], mode, T, nullptr);
  launch_assemble_lower(n, n_pad, T, global_range, nu_fixed, mode, A, n_pad, nullptr);
  return mode;
}

// one rank's slab of the block-cyclic layout (512-wide column panels dealt in a snake; csrc/dist.cu):
// cyclic = 1: every local panel in ONE launch (launch_assemble_cyclic), 0: a launch per panel (launch_assemble_panel)
int emu_dist_slabs(int par, int64_t n, int64_t n_pad, int64_t p, const double* locs, const double* X,
                   const double* theta6, const double* limits, const int* orig, int world, int rank, int64_t nlocal,
                   int cyclic, double* slab) {
  const double lim[2] = {limits ? limits[0] : 0.0, limits ? limits[1] : 0.0};
  double nu_fixed;
  const int mode = smooth_mode_for(par, (int)p, theta6, lim, &nu_fixed);
  const double global_range = 1 / std::exp(-2 * theta6[p]);
  std::vector<double> S((size_t)SF_COUNT * n_pad);
  SiteTable T{S.data(), n_pad, orig};
  launch_site_stage(n, n_pad, (int)p, X, n_pad, locs, n_pad, theta6, lim[0], lim[1], mode, T, nullptr);
  const int64_t kPanelW = 512, npanels = (n_pad + kPanelW - 1) / kPanelW;
  if (cyclic) {
    launch_assemble_cyclic(n, n_pad, T, global_range, nu_fixed, mode, slab, n_pad, world, rank, nlocal, nullptr);
  } else {
    for (int64_t lp = 0; lp < nlocal; ++lp) {
      const int64_t K = lp * world + ((lp & 1) ? world - 1 - rank : rank);
      if (K >= npanels) continue;
      const int64_t width = std::min<int64_t>(kPanelW, n_pad - K * kPanelW);
      launch_assemble_panel(n, n_pad, T, global_range, nu_fixed, mode, slab + lp * kPanelW * n_pad, n_pad,
                            (int)(K * 4), (int)(width / 128), nullptr);
    }
  }
  return mode;
}

// entry vectors of cov_rns_taper (m == 0) / cov_rns_taper_pred (m > 0)
int emu_taper_entries(int64_t n, int64_t m, int64_t p, const double* locs, const double* locs_rows, const double* X,
                      const double* X_rows, const double* theta6, const double* limits, const int* colindices,
                      const int* rowpointers, int64_t nnz, double* out) {
  const bool square = (m == 0);
  const int64_t nrows = square ? n : m;
  double nu_fixed = 0.0;
  const int mode = square ? smooth_mode_for(COCONS_PAR_DIFF, (int)p, theta6, limits, &nu_fixed) : (int)SM_GENERAL;
  std::vector<double> S((size_t)TF_COUNT * n), Sr((size_t)TF_COUNT * std::max<int64_t>(m, 1));
  TaperTable C{S.data(), n}, R{S.data(), n};
  launch_taper_site_stage(n, (int)p, X, n, locs, n, theta6, limits[0], limits[1], mode, 0, C, nullptr);
  if (!square) {
    R = TaperTable{Sr.data(), m};
    launch_taper_site_stage(m, (int)p, X_rows, m, locs_rows, m, theta6, limits[0], limits[1], mode, 1, R, nullptr);
  }
  launch_taper_entries(0, nnz, nrows, colindices, rowpointers, R, C, square ? 1 : 0, mode, nu_fixed,
                       TaperSink{TS_VECTOR, out, nullptr, nullptr, nullptr, 0, 0}, nullptr);
  return mode;
}

// the dense sink of the tapered objective: A (n_pad x n_pad, zeroed by the caller) receives taper[e] * cov[e] on the
// pattern's lower triangle in the context's ordering (inv: caller index -> position), unit diagonal in the padding
int emu_taper_lower(int64_t n, int64_t n_pad, int64_t p, const double* locs_sorted, const double* X_sorted,
                    const double* theta6, const double* limits, const int* colindices, const int* rowpointers,
                    int64_t nnz, const double* taper, const int* inv, double* A) {
  double nu_fixed = 0.0;
  const int mode = smooth_mode_for(COCONS_PAR_DIFF, (int)p, theta6, limits, &nu_fixed);
  std::vector<double> S((size_t)TF_COUNT * n_pad);
  TaperTable T{S.data(), n_pad};
  launch_taper_site_stage(n, (int)p, X_sorted, n_pad, locs_sorted, n_pad, theta6, limits[0], limits[1], mode, 0, T,
                          nullptr);
  launch_taper_pad_diag(n, n_pad, A, n_pad, nullptr);
  launch_taper_entries(0, nnz, n, colindices, rowpointers, T, T, 1, mode, nu_fixed,
                       TaperSink{TS_LOWER, nullptr, taper, inv, A, n_pad, 0}, nullptr);
  return mode;
}

// L Y = B.  L: n_pad x n_pad lower factor, winv: the inverted diagonal tiles (n_pad / 128 of 128 x 128, zeros above the
// diagonal), B: n_pad x (2 nrhs) - right-hand sides in the first nrhs columns, scratch behind them (the library's
// contract); on return the first nrhs columns hold Y.  mode 0: dataflow kernel (K6b) through forward_solve_ws with a
// workspace from solve_workspace_create; 1: the two-kernels-per-tile-step path; 2: the cooperative kernel (K6), one block.

int emu_forward_solve(int64_t n_pad, const double* L, const double* winv, double* B, int nrhs, int mode) {
  const int64_t ld = n_pad, ldb = n_pad, nt = n_pad / kTile;
  if (mode == 0) {
    CholWorkspace ws{};
    int info = 0;
    ws.winv = const_cast<double*>(winv);
    ws.info = &info;
    if (solve_workspace_create(n_pad, &ws) != 0) return -1;
    forward_solve_ws(L, n_pad, ld, ws, B, ldb, nrhs, nullptr);
    const int err = info ? info : (int)ws.solve_ctrl[2];
    solve_workspace_destroy(&ws);
    return err;
  }
  double* Y = B + (int64_t)nrhs * ldb;
  for (int c0 = 0; c0 < nrhs; c0 += 8) {  // the dispatch of forward_solve()
    const int nr = (nrhs - c0 < 8) ? nrhs - c0 : 8;
    double *Bc = B + (int64_t)c0 * ldb, *Yc = Y + (int64_t)c0 * ldb;
    if (nr == 1)
      run_variant<1>(mode, L, ld, winv, Bc, Yc, ldb, nt, nr);
    else if (nr == 2)
      run_variant<2>(mode, L, ld, winv, Bc, Yc, ldb, nt, nr);
    else if (nr <= 4)
      run_variant<4>(mode, L, ld, winv, Bc, Yc, ldb, nt, nr);
    else
      run_variant<8>(mode, L, ld, winv, Bc, Yc, ldb, nt, nr);
  }
  std::memcpy(B, Y, sizeof(double) * (size_t)nrhs * (size_t)ldb);
  return 0;
}

double emu_logdet(const double* L, int64_t n, int64_t ld) {
  double out = 0.0;
  launch_logdet(L, n, ld, &out, nullptr);
  return out;
}

// G (k x k) = Y^T Y over the first n rows of the n_pad-leading-dimension block Y
void emu_gram(const double* Y, int64_t n, int64_t ldy, int k, double* G) {
  std::vector<double> buf(16 * 16 + 296 * 256);
  launch_gram(Y, n, ldy, k, buf.data(), nullptr);
  std::memcpy(G, buf.data(), sizeof(double) * k * k);
}

// ---- csrc/chol.cu --------------------------------------------------------------------------------------------------
// chol_factor() itself (the look-ahead driver with its two streams: launches are synchronous here, so the program
// order of the driver is the execution order).  A: n_pad x n_pad, lower triangle + whole diagonal tiles set;
// winv receives the n_pad / 128 inverted diagonal tiles.  Returns dpotrf's info (0, or the first failing pivot).
int emu_chol_factor(int64_t n_pad, double* A, double* winv) {
  CholWorkspace ws{};
  int info = 0;
  ws.winv = winv, ws.info = &info;
  chol_factor(A, n_pad, n_pad, ws, nullptr);
  return info;
}

// one 128 x 128 diagonal tile: factor in place + explicit inverse (launch_potrf_tile: the blocked DMMA kernel K3b)
int emu_potrf_tile(double* A, int64_t ld, double* winv, int first_index) {
  int info = 0;
  launch_potrf_tile(A, ld, winv, &info, first_index, nullptr);
  return info;
}

// the register-resident column-sweep variant of the tile kernel (K3, COCONS_POTRF=1): half-warp shuffles for the
// pivot, three mbarriers in static shared memory for the one-step-ahead column pipeline
int emu_potrf_tile_sweep(double* A, int64_t ld, double* winv, int first_index) {
  int info = 0;
  int* pinfo = &info;
  emul::launch(dim3(1), dim3(256), true, 0, [&] { potrf_tile_kernel(A, ld, winv, pinfo, first_index); });
  return info;
}

// launch_gemm_nt: mode 0  C -= A B^T (128 x 64 tiles, optionally only the tiles on or below the diagonal),
//                 mode 1  C  = A B^T (128 x 128 tiles)
void emu_gemm_nt(int mode, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* C, int64_t ldc, int lower_only) {
  launch_gemm_nt(mode, M, N, K, A, lda, B, ldb, C, ldc, lower_only, nullptr);
}

// -> number of wrong decodes over *shapes shapes (0 = the numbering is a bijection everywhere)
long emu_tile_decode_check(int ni_max, long* shapes_out) {
  long shapes = 0, bad = 0;
  for (int ni = 1; ni <= ni_max; ++ni) {
    const int step = ni < 80 ? 1 : 7;
    for (int nj = 1; nj <= ni; nj += step) {
      bad += tile_decode_errors<1>(ni, nj, 1), bad += tile_decode_errors<2>(ni, 2 * nj, 1), shapes += 2;
      if (nj <= 12) bad += tile_decode_errors<1>(ni, nj, 0), bad += tile_decode_errors<2>(ni, 2 * nj, 0), shapes += 2;
    }
    bad += tile_decode_errors<2>(ni, 2 * ni, 1), bad += tile_decode_errors<1>(ni, ni, 1), shapes += 2;
  }
  for (int ni : {781, 782, 1563, 1564})
    bad += tile_decode_errors<2>(ni, 2 * ni, 1), bad += tile_decode_errors<2>(ni, 12, 1),
        bad += tile_decode_errors<1>(ni, 1, 0), shapes += 3;
  *shapes_out = shapes;
  return bad;
}

int64_t emu_tile_count(int w, int ni, int njc, int lower_only) {
  return w == 1 ? total_tiles<1>(ni, njc, lower_only) : total_tiles<2>(ni, njc, lower_only);
}

}  // extern "C"
