// Multi-GPU dense likelihood for matrices beyond one GPU's memory (n = 200 000: 320 GB):
// the per-rank half of a block-cyclic right-looking Cholesky.  One process per GPU; this file
// holds every kernel launch of one rank, the exchange step (one broadcast of the factored
// panel per outer step, small reductions in the solve) is issued by the host driver
// (cocons_b200/distributed.py) over NCCL on buffers it owns.
//
// Distribution: column panels of 512 (4 tiles) dealt round-robin to the ranks - the 1 x N case
// of a 2-D block-cyclic grid.  With every peer at full NVSwitch bandwidth the P x Q trade
// (fewer broadcast bytes per GPU against two exchange steps per panel and a distributed panel
// factorisation) does not pay at N <= 8: the whole factorisation moves 4 n^2 bytes per GPU
// (160 GB at n = 200k, ~0.3 s of NVLink time against ~11 s of DMMA time) and the panel stays on
// one GPU, so POTRF/TRSM need no communication at all.
//   * assembly: every rank builds ONLY its own panels, in place, from the replicated per-site
//     table (14 doubles x n) - zero communication;
//   * step K: owner factors panel K (look-ahead: right after it has applied update K-1 to it),
//     packs the rows below it into a contiguous buffer, the driver broadcasts the buffer, every
//     rank applies  C_J -= P_J P_Jcols^T  to its own panels J > K with the DMMA kernel;
//   * solve: the right-hand sides are replicated; each rank accumulates the contributions of
//     its own panels, the driver sums the 512-row block that is due next (one small reduce per
//     panel) and the owner finishes it.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/cocons_b200.h"
#include "common.cuh"

namespace cocons {

constexpr int kPanelTiles = 4;
constexpr int kPanelW = kPanelTiles * kTile;  // 512
constexpr int kDistMaxRhs = 16;

__host__ __device__ inline int snake_owner(int64_t K, int world) {
  const int pos = (int)(K % world);
  return ((K / world) & 1) ? world - 1 - pos : pos;
}

// rows [r0, r0+rows) x w columns of a slab (ld) -> contiguous rows x w
__global__ void pack_rows_kernel(const double* __restrict__ src, int64_t ld, int64_t rows, int w,
                                 double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int c = blockIdx.y; c < w; c += gridDim.y) dst[(int64_t)c * rows + i] = src[(int64_t)c * ld + i];
}

// blocked right-hand sides: panel K occupies nr consecutive columns of kPanelW rows
//   kind ML: columns = z - X mean (r columns); otherwise [design (q columns) | z (r columns)]
__global__ void fill_rhs_kernel(int64_t n, int64_t n_pad, int p, int q, int r, const double* __restrict__ X,
                                const double* __restrict__ Xb, const double* __restrict__ Z,
                                const double* __restrict__ mean, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int nr = q + r;
  const int64_t K = i / kPanelW;
  const int ri = (int)(i % kPanelW);
  double* o = out + K * (int64_t)kPanelW * nr + ri;
  double trend = 0.0;
  if (mean && i < n)
    for (int k = 0; k < p; ++k) trend = fma(X[(int64_t)k * n_pad + i], mean[k], trend);
  for (int c = 0; c < q; ++c) o[(int64_t)c * kPanelW] = (i < n) ? Xb[(int64_t)c * n_pad + i] : 0.0;
  for (int c = 0; c < r; ++c) o[(int64_t)(q + c) * kPanelW] = (i < n) ? Z[(int64_t)c * n_pad + i] - trend : 0.0;
}

// tmp (w x nr, ld w) = b_K - t_K
__global__ void rhs_minus_kernel(const double* __restrict__ b, const double* __restrict__ t, double* __restrict__ out,
                                 int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = b[i] - t[i];
}

// y_K (w x nr, ld w) -> rows K*w.. of the column-major n_pad x nr solution array
__global__ void store_y_kernel(const double* __restrict__ y, int w, int nr, double* __restrict__ Y, int64_t ldy,
                               int64_t row0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w) return;
  for (int c = 0; c < nr; ++c) Y[(int64_t)c * ldy + row0 + i] = y[(int64_t)c * w + i];
}

// acc (blocked like the rhs) += L[rows below panel K, columns of panel K] * y_K
template <int NR>
__global__ void __launch_bounds__(128) acc_update_kernel(const double* __restrict__ L, int64_t ld, int64_t row0,
                                                         int64_t n_pad, int w, const double* __restrict__ y, int nr,
                                                         int nr_total, double* __restrict__ acc) {
  // y: columns of kPanelW doubles; nr columns handled here out of nr_total per blocked panel
  __shared__ double ys[NR][kPanelW];
  for (int idx = threadIdx.x; idx < NR * w; idx += 128) {
    const int c = idx / w, k = idx % w;
    ys[c][k] = (c < nr) ? y[(int64_t)c * kPanelW + k] : 0.0;
  }
  __syncthreads();
  const int64_t i = row0 + (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n_pad) return;
  double a[NR];
#pragma unroll
  for (int c = 0; c < NR; ++c) a[c] = 0.0;
  const double* Lp = L + i;
#pragma unroll 4
  for (int k = 0; k < w; ++k) {
    const double l = Lp[(int64_t)k * ld];
#pragma unroll
    for (int c = 0; c < NR; ++c) a[c] = fma(l, ys[c][k], a[c]);
  }
  const int64_t I = i / kPanelW;
  const int ri = (int)(i % kPanelW);
  double* o = acc + I * (int64_t)kPanelW * nr_total + ri;
#pragma unroll
  for (int c = 0; c < NR; ++c)
    if (c < nr) o[(int64_t)c * kPanelW] += a[c];
}

// sum of log diag over the rank's own panels (rows < n)
__global__ void __launch_bounds__(256) local_logdet_kernel(const double* __restrict__ slab, int64_t ld, int64_t n,
                                                           int rank, int world, int64_t npanels,
                                                           double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t lp = 0; lp * world < npanels; ++lp) {
    const int64_t K = lp * world + ((lp & 1) ? world - 1 - rank : rank);
    if (K >= npanels) continue;
    for (int t = threadIdx.x; t < kPanelW; t += 256) {
      const int64_t g = K * kPanelW + t;
      if (g < n) s += log(slab[(lp * kPanelW + t) * ld + g]);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int wd = 128; wd > 0; wd >>= 1) {
    if (threadIdx.x < wd) red[threadIdx.x] += red[threadIdx.x + wd];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

}  // namespace cocons

using namespace cocons;

struct cocons_dist {
  int device = 0, rank = 0, world = 1;
  int64_t n = 0, n_pad = 0, p = 0, r = 0, q = 0, npanels = 0, nlocal = 0;
  cudaStream_t stream = nullptr;
  std::vector<int64_t> perm;
  double *dX = nullptr, *dLocs = nullptr, *dZ = nullptr, *dXb = nullptr, *dTheta = nullptr, *dSite = nullptr;
  int* dOrig = nullptr;
  double* slab = nullptr;  // n_pad x (local columns), ld = n_pad
  double* tmp = nullptr;   // kPanelW x 2*kDistMaxRhs scratch for the diagonal-block solve
  double* dScal = nullptr;
  CholWorkspace ws{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // the per-panel launches of one update are independent: they CAN be spread over a few streams so that
  // the tail of one launch is filled by the head of the next (opt-in, see cocons_dist_update)
  static constexpr int kUpdStreams = 3;
  cudaStream_t upd[kUpdStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kUpdStreams] = {nullptr, nullptr, nullptr};
  int mode = 0;
  double nu_fixed = 0, global_range = 1, lim[2] = {0, 0};
  int64_t width(int64_t K) const { return std::min<int64_t>(kPanelW, n_pad - K * kPanelW); }
  // panels are dealt in a snake (0..N-1, N-1..0, ...): in every round each rank gets one panel, and the
  // rank that got the longest panel of one round gets the shortest of the next, which evens out the
  // per-step update work (plain round-robin leaves the owner of the earliest panels ~5 % more flops)
  int owner(int64_t K) const { return snake_owner(K, world); }
  int64_t global_panel(int64_t lp) const { return lp * world + ((lp & 1) ? world - 1 - rank : rank); }
  double* panel(int64_t K) const { return slab + (K / world) * (int64_t)kPanelW * n_pad; }
  SiteTable table() const { return SiteTable{dSite, n_pad, dOrig}; }
};

extern "C" {

void cocons_dist_destroy(cocons_dist* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->dX), cudaFree(c->dLocs), cudaFree(c->dZ), cudaFree(c->dXb), cudaFree(c->dTheta), cudaFree(c->dSite);
  cudaFree(c->dOrig), cudaFree(c->slab), cudaFree(c->tmp), cudaFree(c->dScal);
  chol_workspace_destroy(&c->ws);
  for (int i = 0; i < cocons_dist::kUpdStreams; ++i) {
    if (c->upd[i]) cudaStreamDestroy(c->upd[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  delete c;
}

int cocons_dist_create(int device, int rank, int world, int64_t n, int64_t p, int64_t r, const double* locs,
                       const double* X, const double* z, void* stream, cocons_dist** out) {
  if (!out || n <= 0 || p <= 0 || r <= 0 || !locs || !X || !z || world <= 0 || rank < 0 || rank >= world ||
      r >= kDistMaxRhs) {
    set_error("dist_create: bad argument");
    return COCONS_ERR_ARG;
  }
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_error("no CUDA device available; cocons_b200 has no CPU fallback");
    return COCONS_ERR_NO_DEVICE;
  }
  COCONS_CUDA_TRY(cudaSetDevice(device));
  cocons_dist* c = new cocons_dist();
  c->device = device, c->rank = rank, c->world = world;
  c->n = n, c->p = p, c->r = r;
  c->n_pad = (n + kTile - 1) / kTile * kTile;
  c->npanels = (c->n_pad + kPanelW - 1) / kPanelW;
  c->nlocal = 0;
  for (int64_t K = 0; K < c->npanels; ++K)
    if (snake_owner(K, world) == rank) c->nlocal = K / world + 1;  // local slot index = round number
  c->stream = (cudaStream_t)stream;
  const int64_t np = c->n_pad;
  c->perm.resize((size_t)n);
  morton_order(n, locs, c->perm.data());
  bool ok = cudaMalloc(&c->dX, sizeof(double) * np * p) == cudaSuccess &&
            cudaMalloc(&c->dLocs, sizeof(double) * np * 2) == cudaSuccess &&
            cudaMalloc(&c->dZ, sizeof(double) * np * r) == cudaSuccess &&
            cudaMalloc(&c->dTheta, sizeof(double) * 7 * p) == cudaSuccess &&
            cudaMalloc(&c->dSite, sizeof(double) * SF_COUNT * np) == cudaSuccess &&
            cudaMalloc(&c->dOrig, sizeof(int) * np) == cudaSuccess &&
            cudaMalloc(&c->slab, sizeof(double) * np * (size_t)std::max<int64_t>(1, c->nlocal) * kPanelW) ==
                cudaSuccess &&
            cudaMalloc(&c->tmp, sizeof(double) * kPanelW * 2 * kDistMaxRhs) == cudaSuccess &&
            cudaMalloc(&c->dScal, sizeof(double) * (kDistMaxRhs * kDistMaxRhs * 300 + 16)) == cudaSuccess &&
            chol_workspace_create(np, &c->ws) == 0;
  ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < cocons_dist::kUpdStreams; ++i)
    ok = cudaStreamCreateWithFlags(&c->upd[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("dist_create: out of device memory (n_pad=%lld, local panels=%lld)", (long long)np,
              (long long)c->nlocal);
    cocons_dist_destroy(c);
    return COCONS_ERR_ALLOC;
  }
  auto upload = [&](const double* src, int64_t k, double* dst) {
    std::vector<double> tmp((size_t)np * k, 0.0);
    for (int64_t col = 0; col < k; ++col)
      for (int64_t s = 0; s < n; ++s) tmp[(size_t)col * np + s] = src[(size_t)col * n + c->perm[(size_t)s]];
    return cudaMemcpy(dst, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice);
  };
  std::vector<int> orig((size_t)np);
  for (int64_t s = 0; s < np; ++s) orig[(size_t)s] = (s < n) ? (int)c->perm[(size_t)s] : (int)s;
  if (cudaMemcpy(c->dOrig, orig.data(), sizeof(int) * np, cudaMemcpyHostToDevice) != cudaSuccess ||
      upload(X, p, c->dX) != cudaSuccess || upload(locs, 2, c->dLocs) != cudaSuccess ||
      upload(z, r, c->dZ) != cudaSuccess) {
    set_error("dist_create: upload failed");
    cocons_dist_destroy(c);
    return COCONS_ERR_CUDA;
  }
  *out = c;
  return 0;
}

int cocons_dist_set_xbetas(cocons_dist* c, int64_t q, const double* xb) {
  if (!c || q <= 0 || !xb || q + c->r > kDistMaxRhs) {
    set_error("dist_set_xbetas: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  if (c->dXb) cudaFree(c->dXb), c->dXb = nullptr;
  COCONS_CUDA_TRY(cudaMalloc(&c->dXb, sizeof(double) * c->n_pad * q));
  std::vector<double> tmp((size_t)c->n_pad * q, 0.0);
  for (int64_t col = 0; col < q; ++col)
    for (int64_t s = 0; s < c->n; ++s) tmp[(size_t)col * c->n_pad + s] = xb[(size_t)col * c->n + c->perm[(size_t)s]];
  COCONS_CUDA_TRY(cudaMemcpy(c->dXb, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice));
  c->q = q;
  return 0;
}

int64_t cocons_dist_npanels(cocons_dist* c) { return c ? c->npanels : 0; }
int64_t cocons_dist_npad(cocons_dist* c) { return c ? c->n_pad : 0; }

/* doubles in the packed panel K: rows below the panel x its width */
int64_t cocons_dist_panel_elems(cocons_dist* c, int64_t K) {
  if (!c || K < 0 || K >= c->npanels) return 0;
  return (c->n_pad - (K + 1) * kPanelW > 0 ? c->n_pad - (K + 1) * kPanelW : 0) * c->width(K);
}

/* site table + this rank's panels of the lower triangle, in place */
int cocons_dist_assemble(cocons_dist* c, const double* theta6, const double* limits, const double* mean_p) {
  if (!c || !theta6 || !limits) {
    set_error("dist_assemble: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  const int64_t p = c->p, np = c->n_pad;
  c->mode = smooth_mode_for(COCONS_PAR_DIFF, (int)p, theta6, limits, &c->nu_fixed);
  c->global_range = 1 / std::exp(-2 * theta6[p]);
  c->lim[0] = limits[0], c->lim[1] = limits[1];
  std::vector<double> h((size_t)7 * p, 0.0);
  std::memcpy(h.data(), theta6, sizeof(double) * 6 * p);
  if (mean_p) std::memcpy(h.data() + 6 * p, mean_p, sizeof(double) * p);
  COCONS_CUDA_TRY(cudaMemcpyAsync(c->dTheta, h.data(), sizeof(double) * 7 * p, cudaMemcpyHostToDevice, c->stream));
  COCONS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  COCONS_CUDA_TRY(cudaMemsetAsync(c->ws.info, 0, sizeof(int), c->stream));
  launch_site_stage(c->n, np, (int)p, c->dX, np, c->dLocs, np, c->dTheta, limits[0], limits[1], c->mode, c->table(),
                    c->stream);
  // every local panel in ONE launch (a launch per panel was ~1.3 waves of CTAs each: 129 ms at n = 50 000 on one rank
  // against 51 ms for the same pairs in the resident path)
  if (c->world <= 63) {
    launch_assemble_cyclic(c->n, np, c->table(), c->global_range, c->nu_fixed, c->mode, c->slab, np, c->world, c->rank,
                           c->nlocal, c->stream);
  } else {
    for (int64_t lp = 0; lp < c->nlocal; ++lp) {
      const int64_t K = c->global_panel(lp);
      if (K >= c->npanels) continue;
      launch_assemble_panel(c->n, np, c->table(), c->global_range, c->nu_fixed, c->mode, c->panel(K), np,
                            (int)(K * kPanelTiles), (int)(c->width(K) / kTile), c->stream);
    }
  }
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* the context's high-priority side stream (a cudaStream_t): the host driver runs the next panel's
 * factorisation, packing and broadcast on it while the main stream keeps applying the current update */
void* cocons_dist_side_stream(cocons_dist* c) { return c ? (void*)c->ws.panel_stream : nullptr; }

/* owner only: POTRF / panel solve / in-panel update of panel K; side != 0: on the side stream */
int cocons_dist_factor_panel(cocons_dist* c, int64_t K, int side) {
  if (!c || K < 0 || K >= c->npanels || c->owner(K) != c->rank) {
    set_error("dist_factor_panel: panel %lld is not owned by rank %d", (long long)K, c ? c->rank : -1);
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  double* virt = c->panel(K) - K * (int64_t)kPanelW * c->n_pad;  // where column 0 of the full matrix would be
  factor_panel(virt, c->n_pad, c->n_pad, c->ws, K * kPanelTiles, c->width(K) / kTile,
               side ? c->ws.panel_stream : c->stream);
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* owner only: rows below panel K, contiguous, into dst (device); side != 0: on the side stream */
int cocons_dist_pack_panel(cocons_dist* c, int64_t K, void* dst, int side) {
  if (!c || !dst || K < 0 || K >= c->npanels || c->owner(K) != c->rank) {
    set_error("dist_pack_panel: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  const int64_t r0 = (K + 1) * kPanelW, rows = c->n_pad - r0;
  if (rows <= 0) return 0;
  note_launch();
  pack_rows_kernel<<<dim3((unsigned)((rows + 255) / 256), 32), 256, 0, side ? c->ws.panel_stream : c->stream>>>(c->panel(K) + r0, c->n_pad, rows,
                                                                                   (int)c->width(K), (double*)dst);
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* every rank: apply the update of (packed) panel K to the rank's own panels J, J_lo <= J < J_hi */
int cocons_dist_update(cocons_dist* c, int64_t K, const void* src, int64_t J_lo, int64_t J_hi) {
  if (!c || !src || K < 0 || K >= c->npanels) {
    set_error("dist_update: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  const int64_t r0 = (K + 1) * kPanelW, rows = c->n_pad - r0;
  if (rows <= 0) return 0;
  const double* P = (const double*)src;
  if (J_lo <= K) J_lo = K + 1;
  if (J_hi > c->npanels) J_hi = c->npanels;
  // The per-panel launches of a step are independent: they are spread over three streams so that the tail of
  // one launch is filled by the next (+7 % at 2 GPUs, 255 -> 264 TFLOP/s at 8 GPUs, n = 200 000).  Round 1 kept
  // this opt-in because a repeated evaluation then reported a false "not positive definite"; that was the slot
  // release race of the GEMM pipeline (csrc/chol.cu, fixed in round 2).  COCONS_DIST_UPD_STREAMS=1 puts every
  // update back on the main stream.
  static int nstreams = -1;
  if (nstreams < 0) {
    const char* e = getenv("COCONS_DIST_UPD_STREAMS");
    nstreams = e ? std::max(1, std::min(atoi(e), (int)cocons_dist::kUpdStreams)) : (int)cocons_dist::kUpdStreams;
  }
  int launched = 0;
  for (int64_t J = J_lo; J < J_hi; ++J) {
    if (c->owner(J) != c->rank) continue;
    const int64_t off = J * kPanelW - r0;  // first row of panel J inside the packed buffer
    const double* Pj = P + off;
    cudaStream_t s = c->stream;
    if (nstreams > 1) {
      const int sidx = launched % nstreams;
      if (launched == 0) COCONS_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
      if (launched < nstreams) COCONS_CUDA_TRY(cudaStreamWaitEvent(c->upd[sidx], c->ev_fork, 0));
      s = c->upd[sidx];
    }
    launch_gemm_nt(0, c->n_pad - J * kPanelW, c->width(J), c->width(K), Pj, rows, Pj, rows,
                   c->panel(J) + J * kPanelW, c->n_pad, 1, s);
    ++launched;
  }
  for (int i = 0; nstreams > 1 && i < launched && i < nstreams; ++i) {  // join
    COCONS_CUDA_TRY(cudaEventRecord(c->ev_join[i], c->upd[i]));
    COCONS_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
  }
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* blocked right-hand sides into rhs (device, npanels x nr x 512): kind ML -> z - X mean; else [design | z] */
int cocons_dist_fill_rhs(cocons_dist* c, int kind, void* rhs, int* nr_out) {
  if (!c || !rhs || !nr_out) {
    set_error("dist_fill_rhs: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  int q = 0;
  const double* xb = nullptr;
  if (kind == COCONS_PROFILE) {
    if (!c->dXb) {
      set_error("dist_fill_rhs: COCONS_PROFILE needs cocons_dist_set_xbetas first");
      return COCONS_ERR_STATE;
    }
    q = (int)c->q, xb = c->dXb;
  } else if (kind == COCONS_REML) {
    q = (int)c->p, xb = c->dX;
  }
  if (q + c->r > kDistMaxRhs) {
    set_error("dist_fill_rhs: too many right-hand sides");
    return COCONS_ERR_ARG;
  }
  note_launch();
  fill_rhs_kernel<<<(unsigned)((c->n_pad + 255) / 256), 256, 0, c->stream>>>(
      c->n, c->n_pad, (int)c->p, q, (int)c->r, c->dX, xb, c->dZ, kind == COCONS_ML ? c->dTheta + 6 * c->p : nullptr,
      (double*)rhs);
  *nr_out = q + (int)c->r;
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* owner only: y_K = inv(L_KK) (b_K - t_K);  Y[rows of K] = y_K;  acc[rows below] += L[rows, K] y_K.
 * b, t: blocks of panel K (512 x nr, ld 512); acc: the whole blocked accumulator; Y: n_pad x nr. */
int cocons_dist_solve_block(cocons_dist* c, int64_t K, const void* bK, const void* tK, void* acc, void* Y, int nr) {
  if (!c || !bK || !tK || !acc || !Y || nr <= 0 || nr > kDistMaxRhs || c->owner(K) != c->rank) {
    set_error("dist_solve_block: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  const int w = (int)c->width(K);
  cudaStream_t st = c->stream;
  note_launch(2);
  rhs_minus_kernel<<<(kPanelW * nr + 255) / 256, 256, 0, st>>>((const double*)bK, (const double*)tK, c->tmp,
                                                              kPanelW * nr);
  // the diagonal block of the panel is a w x w lower factor with its own inverted tiles (a narrower
  // last panel solves on the leading w rows of each 512-long column)
  forward_solve(c->panel(K) + K * kPanelW, w, c->n_pad, c->ws.winv + K * kPanelTiles * (int64_t)kTile * kTile, c->tmp,
                kPanelW, nr, st);
  store_y_kernel<<<(w + 127) / 128, 128, 0, st>>>(c->tmp, kPanelW, nr, (double*)Y, c->n_pad, K * kPanelW);
  const int64_t r0 = (K + 1) * kPanelW, rows = c->n_pad - r0;
  if (rows > 0) {
    for (int c0 = 0; c0 < nr; c0 += 8) {  // 8 right-hand sides per pass (32 KB of shared y)
      note_launch();
      acc_update_kernel<8><<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(
          c->panel(K), c->n_pad, r0, c->n_pad, w, c->tmp + (int64_t)c0 * kPanelW, std::min(8, nr - c0), nr,
          (double*)acc + (int64_t)c0 * kPanelW);
    }
  }
  COCONS_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* local pieces of the final reductions, written to device memory the driver all-reduces:
 *   out[0] = sum log diag over own panels; out[1] = info flag; gram (nr x nr) = Y^T Y over the rows this
 *   rank solved (Y is zero elsewhere) */
int cocons_dist_reduce_local(cocons_dist* c, const void* Y, int nr, void* out2, void* gram) {
  if (!c || !Y || !out2 || !gram || nr <= 0 || nr > kDistMaxRhs) {
    set_error("dist_reduce_local: bad argument");
    return COCONS_ERR_ARG;
  }
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  note_launch();
  local_logdet_kernel<<<1, 256, 0, st>>>(c->slab, c->n_pad, c->n, c->rank, c->world, c->npanels, (double*)out2);
  launch_gram((const double*)Y, c->n, c->n_pad, nr, c->dScal, st);
  COCONS_CUDA_TRY(cudaMemcpyAsync(gram, c->dScal, sizeof(double) * nr * nr, cudaMemcpyDeviceToDevice, st));
  int info = 0;
  COCONS_CUDA_TRY(cudaMemcpyAsync(&info, c->ws.info, sizeof(int), cudaMemcpyDeviceToHost, st));
  COCONS_CUDA_TRY(cudaStreamSynchronize(st));
  const double dinfo = (double)info;
  COCONS_CUDA_TRY(cudaMemcpyAsync((double*)out2 + 1, &dinfo, sizeof(double), cudaMemcpyHostToDevice, st));
  COCONS_CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

/* caller-order permutation of the context (perm[s] = caller index of internal site s) */
int cocons_dist_perm(cocons_dist* c, int64_t* perm) {
  if (!c || !perm) return COCONS_ERR_ARG;
  for (int64_t s = 0; s < c->n; ++s) perm[s] = c->perm[(size_t)s];
  return 0;
}

}  // extern "C"
