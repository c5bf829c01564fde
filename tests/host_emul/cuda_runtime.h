// TEST SCAFFOLDING - a stand-in for <cuda_runtime.h> that lets g++ compile the pairwise-assembly kernels
// (cocons_b200/csrc/assembly.cu, taper.cu, bessel.cuh) and RUN them on the host, thread for thread, so the
// `-m "not gpu"` suite can check the very kernel source that ships - its arithmetic order, the staging through
// shared memory, tile / slab / slice index math - against the reference-made goldens without a device
// (tests/test_host_emul.py).  Nothing under cocons_b200/ includes or links this; the product library has no
// CPU path (every computing entry returns COCONS_ERR_NO_DEVICE without an sm_100 device).
//
// Model: a kernel launch `k<<<grid, block, 0, st>>>(args)` is rewritten by the test (tests/host_emul/build.py)
// into emul::launch(grid, block, has_barrier, [&] { k(args); }).  Blocks run one after another.  A kernel
// without __syncthreads() runs its threads one after another too; one with barriers gets `block` OS threads
// that live for the whole launch and meet at a barrier that, like the hardware's, counts exited threads as
// arrived.  __shared__ becomes `static` (one block at a time, so one copy is what a block sees).
// Round-to-nearest intrinsics map to the plain IEEE operation (compile with -ffp-contract=off), __fma_rn to
// fma(); exp / sin / cos / log come from glibc, which legitimately differs from CUDA's libm by an ulp.
#ifndef COCONS_TEST_CUDA_EMUL_H
#define COCONS_TEST_CUDA_EMUL_H

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }

namespace emul {

// __syncthreads() of one block: releases when every thread of the block has either arrived or left the kernel
class BlockBarrier {
 public:
  explicit BlockBarrier(unsigned n) : n_(n) {}
  void sync() {
    std::unique_lock<std::mutex> lk(m_);
    ++waiting_;
    release_or_wait(lk);
  }
  // the calling thread has returned from the kernel: counts as arrived for every later barrier of this block;
  // returns once the whole block has finished (the next block reuses the static "shared" arrays)
  void leave_block() {
    std::unique_lock<std::mutex> lk(m_);
    ++exited_;
    if (exited_ == n_) {
      exited_ = 0;
      ++block_gen_;
      cv_.notify_all();
      return;
    }
    if (waiting_ && waiting_ + exited_ == n_) {
      waiting_ = 0;
      ++gen_;
      cv_.notify_all();
    }
    const unsigned long g = block_gen_;
    cv_.wait(lk, [&] { return block_gen_ != g; });
  }

 private:
  void release_or_wait(std::unique_lock<std::mutex>& lk) {
    if (waiting_ + exited_ == n_) {
      waiting_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      const unsigned long g = gen_;
      cv_.wait(lk, [&] { return gen_ != g; });
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  unsigned n_, waiting_ = 0, exited_ = 0;
  unsigned long gen_ = 0, block_gen_ = 0;
};

struct ThreadCtx {
  dim3 tid, bid, bdim, gdim;
  BlockBarrier* bar = nullptr;
};
inline thread_local ThreadCtx ctx;
inline long launches = 0, barrier_launches = 0;

template <class Body>
void launch(dim3 grid, dim3 block, bool has_barrier, Body&& body) {
  ++launches;
  const unsigned nt = block.x * block.y * block.z;
  auto thread_id = [&](unsigned t) { return dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y)); };
  if (!has_barrier) {
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
          for (unsigned t = 0; t < nt; ++t) {
            ctx.tid = thread_id(t), ctx.bid = dim3(bx, by, bz), ctx.bdim = block, ctx.gdim = grid, ctx.bar = nullptr;
            body();
          }
    return;
  }
  ++barrier_launches;
  BlockBarrier bar(nt);
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (unsigned t = 0; t < nt; ++t)
    pool.emplace_back([&, t] {
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            ctx.tid = thread_id(t), ctx.bid = dim3(bx, by, bz), ctx.bdim = block, ctx.gdim = grid, ctx.bar = &bar;
            body();
            bar.leave_block();
          }
    });
  for (auto& th : pool) th.join();
}

}  // namespace emul

#define threadIdx (emul::ctx.tid)
#define blockIdx (emul::ctx.bid)
#define blockDim (emul::ctx.bdim)
#define gridDim (emul::ctx.gdim)

inline void __syncthreads() {
  if (!emul::ctx.bar) {
    std::fprintf(stderr, "host emulation: __syncthreads() in a kernel launched without barrier support\n");
    std::abort();
  }
  emul::ctx.bar->sync();
}

// IEEE round-to-nearest intrinsics (the build uses -ffp-contract=off, so a * b + c is never fused behind our back)
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
using std::max;
using std::min;

#endif
