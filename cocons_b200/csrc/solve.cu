// Post-factorisation kernels (K6/K7): log-determinant, blocked forward
// substitution with many right-hand sides, and the small Gram reductions from
// which the ML / profile / REML quadratic forms are taken.  They replace
//   sum(log(diag(cholS)))                          R/neg2loglikelihood.R:153,208,267
//   forwardsolve(cholS, ., transpose=TRUE, ...)    :145-147, :214-217, :273-275
//   crossprod(.)                                   :148, :157, :214, :276, :285
// All of them are HBM-bound: the factor is streamed once per block of
// right-hand sides (4 n^2 bytes).
#include <cooperative_groups.h>

#include "../../include/cocons_b200.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace cocons {

// sum_{i<n} log L_ii, one CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(1024) logdet_kernel(const double* __restrict__ L, int64_t n, int64_t ld,
                                                      double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += log(L[i * ld + i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

void launch_logdet(const double* L, int64_t n, int64_t ld, double* out, cudaStream_t st) {
  note_launch();
  logdet_kernel<<<1, 1024, 0, st>>>(L, n, ld, out);
}

// ---------------------------------------------------------------------------
// Blocked forward substitution  L Y = B  for NR right-hand sides in ONE cooperative launch:
//   for every tile J:   y_J = inv(L_JJ) b_J ;   b_I -= L_IJ y_J  for every tile row I > J ;  grid sync.
// Every CTA recomputes y_J in shared memory (a 128 x 128 product out of L2) instead of waiting for a
// broadcast, then updates the tile rows it owns (I = J+1+blockIdx.x, stride gridDim.x); CTA 0 stores y_J.
// The n/128 dependent steps cost one grid barrier each instead of one kernel launch each.
// ---------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(256) fwd_solve_coop_kernel(const double* __restrict__ L, int64_t ld,
                                                             const double* __restrict__ Winv, double* B,
                                                             double* __restrict__ Y, int64_t ldb, int64_t nt, int nr) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double bj[NR][kTile];
  __shared__ double yj[NR][kTile];
  __shared__ double part[NR][256];
  const int tid = threadIdx.x;
  const int row = tid & (kTile - 1), h = tid >> 7;
  for (int64_t J = 0; J < nt; ++J) {
    const int64_t j0 = J * kTile;
    const double* W = Winv + J * (int64_t)kTile * kTile;
    for (int idx = tid; idx < NR * kTile; idx += 256) {
      const int c = idx / kTile, k = idx % kTile;
      bj[c][k] = (c < nr) ? __ldcg(B + (int64_t)c * ldb + j0 + k) : 0.0;  // written by other CTAs: bypass L1
    }
    __syncthreads();
    {  // y = W b: thread (row, half) sums half of the k range of its row
      double acc[NR];
#pragma unroll
      for (int c = 0; c < NR; ++c) acc[c] = 0.0;
      // W is stored with explicit zeros above the diagonal: fixed trip count, 16 loads in flight
      const double* Wp = W + (int64_t)h * (kTile / 2) * kTile + row;
#pragma unroll 16
      for (int k = 0; k < kTile / 2; ++k) {
        const double w = __ldg(Wp + (int64_t)k * kTile);
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[c] = fma(w, bj[c][h * (kTile / 2) + k], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NR; ++c) part[c][tid] = acc[c];
    }
    __syncthreads();
    if (tid < kTile) {
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        const double y = part[c][tid] + part[c][tid + kTile];
        yj[c][tid] = y;
        if (blockIdx.x == 0 && c < nr) Y[(int64_t)c * ldb + j0 + tid] = y;
      }
    }
    __syncthreads();
    for (int64_t I = J + 1 + blockIdx.x; I < nt; I += gridDim.x) {  // b_I -= L_IJ y_J
      const int64_t i0 = I * kTile;
      double acc[NR];
#pragma unroll
      for (int c = 0; c < NR; ++c) acc[c] = 0.0;
      const double* Lp = L + (j0 + (int64_t)h * (kTile / 2)) * ld + i0 + row;
#pragma unroll 16
      for (int k = 0; k < kTile / 2; ++k) {
        const double l = __ldg(Lp + (int64_t)k * ld);
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[c] = fma(l, yj[c][h * (kTile / 2) + k], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NR; ++c) part[c][tid] = acc[c];
      __syncthreads();
      if (tid < kTile) {
#pragma unroll
        for (int c = 0; c < NR; ++c)
          if (c < nr) {  // this row block was last updated by ANOTHER CTA (previous step): read it from L2
            double* bp = B + (int64_t)c * ldb + i0 + tid;
            *bp = __ldcg(bp) - (part[c][tid] + part[c][tid + kTile]);
          }
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------------------
// The same substitution WITHOUT a cooperative launch: two ordinary kernels per tile step (y_J = W_J b_J by one
// CTA; b_I -= L_IJ y_J by one CTA per remaining tile row), same arithmetic and summation order as the
// cooperative kernel.  Opt-in (COCONS_SOLVE_COOP=0): written at the end of round 1 to test whether the
// cooperative launch is what makes overlapping evaluations irreproducible; it costs 2 n/128
// launches per solve instead of one and has not been run on a GPU yet.
// ---------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(256) fwd_tile_solve_kernel(const double* __restrict__ W, double* __restrict__ B,
                                                             double* __restrict__ Y, int64_t ldb, int64_t j0, int nr) {
  __shared__ double bj[NR][kTile];
  __shared__ double part[NR][256];
  const int tid = threadIdx.x;
  const int row = tid & (kTile - 1), h = tid >> 7;
  for (int idx = tid; idx < NR * kTile; idx += 256) {
    const int c = idx / kTile, k = idx % kTile;
    bj[c][k] = (c < nr) ? __ldcg(B + (int64_t)c * ldb + j0 + k) : 0.0;
  }
  __syncthreads();
  double acc[NR];
#pragma unroll
  for (int c = 0; c < NR; ++c) acc[c] = 0.0;
  const double* Wp = W + (int64_t)h * (kTile / 2) * kTile + row;
#pragma unroll 16
  for (int k = 0; k < kTile / 2; ++k) {
    const double w = __ldg(Wp + (int64_t)k * kTile);
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = fma(w, bj[c][h * (kTile / 2) + k], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NR; ++c) part[c][tid] = acc[c];
  __syncthreads();
  if (tid < kTile) {
#pragma unroll
    for (int c = 0; c < NR; ++c)
      if (c < nr) Y[(int64_t)c * ldb + j0 + tid] = part[c][tid] + part[c][tid + kTile];
  }
}

template <int NR>
__global__ void __launch_bounds__(256) fwd_tile_update_kernel(const double* __restrict__ L, int64_t ld,
                                                              const double* __restrict__ Y, double* __restrict__ B,
                                                              int64_t ldb, int64_t J, int nr) {
  __shared__ double yj[NR][kTile];
  __shared__ double part[NR][256];
  const int tid = threadIdx.x;
  const int row = tid & (kTile - 1), h = tid >> 7;
  const int64_t j0 = J * kTile, i0 = (J + 1 + blockIdx.x) * kTile;
  for (int idx = tid; idx < NR * kTile; idx += 256) {
    const int c = idx / kTile, k = idx % kTile;
    yj[c][k] = (c < nr) ? __ldcg(Y + (int64_t)c * ldb + j0 + k) : 0.0;
  }
  __syncthreads();
  double acc[NR];
#pragma unroll
  for (int c = 0; c < NR; ++c) acc[c] = 0.0;
  const double* Lp = L + (j0 + (int64_t)h * (kTile / 2)) * ld + i0 + row;
#pragma unroll 16
  for (int k = 0; k < kTile / 2; ++k) {
    const double l = __ldg(Lp + (int64_t)k * ld);
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = fma(l, yj[c][h * (kTile / 2) + k], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NR; ++c) part[c][tid] = acc[c];
  __syncthreads();
  if (tid < kTile) {
#pragma unroll
    for (int c = 0; c < NR; ++c)
      if (c < nr) {
        double* bp = B + (int64_t)c * ldb + i0 + tid;
        *bp = __ldcg(bp) - (part[c][tid] + part[c][tid + kTile]);
      }
  }
}

template <int NR>
static void launch_fwd_steps(const double* L, int64_t ld, const double* winv, double* B, double* Y, int64_t ldb,
                             int64_t nt, int nr, cudaStream_t st) {
  for (int64_t J = 0; J < nt; ++J) {
    note_launch();
    fwd_tile_solve_kernel<NR><<<1, 256, 0, st>>>(winv + J * (int64_t)kTile * kTile, B, Y, ldb, J * kTile, nr);
    if (J + 1 < nt) {
      note_launch();
      fwd_tile_update_kernel<NR><<<(unsigned)(nt - J - 1), 256, 0, st>>>(L, ld, Y, B, ldb, J, nr);
    }
  }
}

static bool solve_uses_cooperative_launch() {
  static int coop = -1;
  if (coop < 0) {
    const char* e = getenv("COCONS_SOLVE_COOP");
    coop = (e && atoi(e) == 0) ? 0 : 1;
  }
  return coop != 0;
}

template <int NR>
static void launch_fwd_coop(const double* L, int64_t ld, const double* winv, double* B, double* Y, int64_t ldb,
                            int64_t nt, int nr, cudaStream_t st) {
  static int max_blocks[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && max_blocks[dev] == 0) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd_solve_coop_kernel<NR>, 256, 0);
    max_blocks[dev] = sms * (per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm));
  }
  if (!solve_uses_cooperative_launch()) {
    launch_fwd_steps<NR>(L, ld, winv, B, Y, ldb, nt, nr, st);
    return;
  }
  int grid = (dev < 16) ? max_blocks[dev] : 148;
  if (grid > nt) grid = (int)(nt > 0 ? nt : 1);
  void* args[] = {(void*)&L, (void*)&ld, (void*)&winv, (void*)&B, (void*)&Y, (void*)&ldb, (void*)&nt, (void*)&nr};
  note_launch();
  cudaLaunchCooperativeKernel((void*)fwd_solve_coop_kernel<NR>, dim3(grid), dim3(256), args, 0, st);
}

// L (n_pad x n_pad, lower) Y = B for nrhs columns; on return B holds Y.
// Y scratch is provided by the caller through the upper half of B's allocation:
// B must have room for 2*nrhs columns of ldb doubles.
void forward_solve(const double* L, int64_t n_pad, int64_t ld, const double* winv, double* B, int64_t ldb, int nrhs,
                   cudaStream_t st) {
  const int64_t nt = n_pad / kTile;
  double* Y = B + (int64_t)nrhs * ldb;
  for (int c0 = 0; c0 < nrhs; c0 += 8) {
    const int nr = (nrhs - c0 < 8) ? nrhs - c0 : 8;
    double* Bc = B + (int64_t)c0 * ldb;
    double* Yc = Y + (int64_t)c0 * ldb;
    if (nr == 1)
      launch_fwd_coop<1>(L, ld, winv, Bc, Yc, ldb, nt, nr, st);
    else if (nr == 2)
      launch_fwd_coop<2>(L, ld, winv, Bc, Yc, ldb, nt, nr, st);
    else if (nr <= 4)
      launch_fwd_coop<4>(L, ld, winv, Bc, Yc, ldb, nt, nr, st);
    else
      launch_fwd_coop<8>(L, ld, winv, Bc, Yc, ldb, nt, nr, st);
  }
  cudaMemcpyAsync(B, Y, sizeof(double) * (size_t)nrhs * (size_t)ldb, cudaMemcpyDeviceToDevice, st);
}

// ---------------------------------------------------------------------------
// G = Y^T Y for a tall n x k block (k <= 16), two deterministic passes.
// ---------------------------------------------------------------------------
constexpr int kGramMaxK = 16;
constexpr int kGramBlocks = 296;

__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ Y, int64_t n, int64_t ldy, int k,
                                                           double* __restrict__ partial) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b <= a; ++b) {
      double s = 0.0;
      for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < n; i += (int64_t)gridDim.x * 256)
        s = fma(Y[(int64_t)a * ldy + i], Y[(int64_t)b * ldy + i], s);
      red[tid] = s;
      __syncthreads();
      for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) red[tid] += red[tid + w];
        __syncthreads();
      }
      if (tid == 0) partial[(int64_t)blockIdx.x * kGramMaxK * kGramMaxK + a * kGramMaxK + b] = red[0];
      __syncthreads();
    }
}

__global__ void gram_final_kernel(const double* __restrict__ partial, int nblocks, int k, double* __restrict__ G) {
  const int a = threadIdx.x / kGramMaxK, b = threadIdx.x % kGramMaxK;
  if (a >= k || b > a) return;
  double s = 0.0;
  for (int blk = 0; blk < nblocks; ++blk) s += partial[(int64_t)blk * kGramMaxK * kGramMaxK + a * kGramMaxK + b];
  G[a * k + b] = s;
  G[b * k + a] = s;
}

// G (k x k, host layout row-major == column-major by symmetry); scratch: kGramBlocks*256 doubles after G
void launch_gram(const double* Y, int64_t n, int64_t ldy, int k, double* G, cudaStream_t st) {
  double* partial = G + kGramMaxK * kGramMaxK;
  note_launch(2);
  gram_partial_kernel<<<kGramBlocks, 256, 0, st>>>(Y, n, ldy, k, partial);
  gram_final_kernel<<<1, kGramMaxK * kGramMaxK, 0, st>>>(partial, kGramBlocks, k, G);
}

}  // namespace cocons
