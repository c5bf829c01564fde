// Round-2 reproducer / bisection harness for the factorisation's reproducibility.
// S independent SPD matrices are factored one at a time (reference) and then all at once on S streams,
// `passes` times; every concurrent factor is compared bit for bit with its serial factor and the FIRST
// wrong tile (smallest tile column, then tile row) is located together with the shape of the damage
// inside it - the granularity (a 1 KB bulk row, a 16-column stage, a 128 x 64 CTA tile, a whole
// column block) says which mechanism lost or re-ordered an update.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DCOCONS_GEMM_...] tools/micro/chol_race.cu -o chol_race
//   chol_race S n passes [self]     self = 1: S = 1 and the reference is the first pass of the same stream
//                                   (single-chain run-to-run reproducibility, the bench.py situation)
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
__global__ void diag_kernel(double* x, int64_t n, double v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i * n + i] = v;
}
// one CTA per lower 128 x 128 tile: number of differing entries (lower triangle of the matrix only)
__global__ void tile_diff_kernel(const double* a, const double* b, int64_t n, int nt, unsigned* bad) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (ti < tj) return;
  unsigned c = 0;
  for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
    const int64_t i = ti * 128 + (e & 127), j = tj * 128 + (e >> 7);
    if (i >= j && __double_as_longlong(a[j * n + i]) != __double_as_longlong(b[j * n + i])) ++c;
  }
  if (c) atomicAdd(&bad[tj * nt + ti], c);
}
// damage map of one tile: per column and per row counts
__global__ void tile_map_kernel(const double* a, const double* b, int64_t n, int ti, int tj, unsigned* colcnt,
                                unsigned* rowcnt, double* maxrel) {
  for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
    const int r = e & 127, c = e >> 7;
    const int64_t i = ti * 128 + r, j = tj * 128 + c;
    if (i < j) continue;
    const double x = a[j * n + i], y = b[j * n + i];
    if (__double_as_longlong(x) != __double_as_longlong(y)) {
      atomicAdd(&colcnt[c], 1u), atomicAdd(&rowcnt[r], 1u);
      const double rel = fabs(x - y) / fmax(fabs(y), 1e-300);
      atomicMax((unsigned long long*)maxrel, (unsigned long long)__double_as_longlong(rel));
    }
  }
}

static void summarize(const char* what, const unsigned* v) {
  printf("      %s:", what);
  int run0 = -1;
  for (int i = 0; i <= 128; ++i) {
    const bool on = i < 128 && v[i];
    if (on && run0 < 0) run0 = i;
    if (!on && run0 >= 0) { printf(" [%d..%d]", run0, i - 1); run0 = -1; }
  }
  printf("\n");
}

int main(int argc, char** argv) {
  int S = argc > 1 ? atoi(argv[1]) : 4;
  const int64_t n = argc > 2 ? atoll(argv[2]) : 12032;
  const int passes = argc > 3 ? atoi(argv[3]) : 8;
  const int self = argc > 4 ? atoi(argv[4]) : 0;
  if (self) S = 1;
  const int nt = (int)(n / 128);
  std::vector<double*> A(S), A0(S), Aref(S);
  std::vector<cudaStream_t> st(S);
  std::vector<cocons::CholWorkspace> ws(S);
  unsigned *dbad, *dmap;
  double* dmax;
  cudaMalloc(&dbad, sizeof(unsigned) * nt * nt), cudaMalloc(&dmap, sizeof(unsigned) * 256), cudaMalloc(&dmax, 8);
  for (int s = 0; s < S; ++s) {
    if (cudaMalloc(&A[s], sizeof(double) * n * n) != cudaSuccess || cudaMalloc(&A0[s], sizeof(double) * n * n) != cudaSuccess ||
        cudaMalloc(&Aref[s], sizeof(double) * n * n) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
    if (cocons::chol_workspace_create(n, &ws[s]) != 0) { printf("workspace failed\n"); return 1; }
    init_kernel<<<1024, 256>>>(A0[s], (size_t)n * n, 17u + s, 0.01);
    diag_kernel<<<(unsigned)((n + 255) / 256), 256>>>(A0[s], n, 40.0 + s + 0.002 * n);
  }
  cudaDeviceSynchronize();
  for (int s = 0; s < S; ++s) {  // reference: one factorisation at a time, nothing else on the device
    cudaMemcpy(Aref[s], A0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice);
    cocons::chol_factor(Aref[s], n, n, ws[s], st[s]);
    cudaDeviceSynchronize();
    int info = -1;
    cudaMemcpy(&info, ws[s].info, 4, cudaMemcpyDeviceToHost);
    if (info) printf("reference %d: info=%d\n", s, info);
  }
  int bad_passes = 0;
  std::vector<unsigned> hbad((size_t)nt * nt);
  for (int pass = 0; pass < passes; ++pass) {
    for (int s = 0; s < S; ++s) cudaMemcpyAsync(A[s], A0[s], sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st[s]);
    for (int s = 0; s < S; ++s) cocons::chol_factor(A[s], n, n, ws[s], st[s]);
    cudaError_t e = cudaDeviceSynchronize();
    bool any = false;
    for (int s = 0; s < S; ++s) {
      cudaMemset(dbad, 0, sizeof(unsigned) * nt * nt);
      tile_diff_kernel<<<dim3(nt, nt), 256>>>(A[s], Aref[s], n, nt, dbad);
      cudaMemcpy(hbad.data(), dbad, sizeof(unsigned) * nt * nt, cudaMemcpyDeviceToHost);
      int info = -1;
      cudaMemcpy(&info, ws[s].info, 4, cudaMemcpyDeviceToHost);
      unsigned long long total = 0;
      int ftj = -1, fti = -1, tiles_in_col = 0;
      for (int tj = 0; tj < nt; ++tj)
        for (int ti = tj; ti < nt; ++ti)
          if (hbad[(size_t)tj * nt + ti]) {
            total += hbad[(size_t)tj * nt + ti];
            if (ftj < 0) ftj = tj, fti = ti;
            if (tj == ftj) ++tiles_in_col;
          }
      if (!total && !info) continue;
      any = true;
      printf("  pass %d matrix %d: %llu entries differ, info=%d; first wrong tile column %d (of %d; outer panel %d, step %d in it): "
             "%d wrong tiles in that column, first tile row %d (%u entries)\n",
             pass, s, total, info, ftj, nt, ftj >= 0 ? ftj / cocons::chol_outer(n) : -1, ftj >= 0 ? ftj % cocons::chol_outer(n) : -1,
             tiles_in_col, fti, ftj >= 0 ? hbad[(size_t)ftj * nt + fti] : 0);
      if (ftj >= 0) {
        printf("    wrong tile rows in column %d:", ftj);
        for (int ti = ftj; ti < nt; ++ti)
          if (hbad[(size_t)ftj * nt + ti]) printf(" %d(%u)", ti, hbad[(size_t)ftj * nt + ti]);
        printf("\n");
        unsigned hmap[256];
        double hmax = 0;
        cudaMemset(dmap, 0, sizeof(unsigned) * 256), cudaMemset(dmax, 0, 8);
        tile_map_kernel<<<1, 256>>>(A[s], Aref[s], n, fti, ftj, dmap, dmap + 128, dmax);
        cudaMemcpy(hmap, dmap, sizeof(hmap), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hmax, dmax, 8, cudaMemcpyDeviceToHost);
        printf("    first wrong tile (%d,%d): max rel diff %.3e\n", fti, ftj, hmax);
        summarize("columns hit", hmap), summarize("rows hit   ", hmap + 128);
      }
    }
    bad_passes += any;
    if (self && pass == 0 && !any) {}
    printf("CHOL_RACE matrices=%d n=%lld pass %d: cuda=%s %s\n", S, (long long)n, pass, cudaGetErrorString(e),
           any ? "MISMATCH" : "identical");
    fflush(stdout);
  }
  printf("CHOL_RACE_SUMMARY matrices=%d n=%lld passes=%d mismatching_passes=%d\n", S, (long long)n, passes, bad_passes);
  return 0;
}
