// TEST SCAFFOLDING - a stand-in for <cuda_runtime.h> that lets g++ compile EVERY source of libcocons_b200.so
// (cocons_b200/csrc/*.cu, bessel.cuh) and RUN the kernels on the host, thread for thread, so the `-m "not gpu"` suite
// can check the very source that ships - arithmetic order, staging through shared memory, tile / slab / slice index
// math, the mbarrier / bulk-copy pipeline of the DMMA GEMM, the dataflow solve, the host orchestration of the C ABI -
// against the reference-made goldens without a device, and so that the ordinary sanitizers can look at the kernels
// (tools/emul_memcheck.sh, tools/emul_racecheck.sh).  Nothing under cocons_b200/ includes or links this; the product
// library has no CPU path (every computing entry returns COCONS_ERR_NO_DEVICE without an sm_100 device).
//
// Model: a kernel launch `k<<<grid, block, smem, st>>>(args)` is rewritten by the test (tests/host_emul/build.py)
// into emul::launch(grid, block, has_barrier, smem, [&] { k(args); }); `extern __shared__ T name[];` becomes a
// pointer to the launch's dynamic shared memory; a one-statement inline-PTX helper becomes a call to its stand-in
// emul::ptx_<name> below.  Launches are synchronous (streams and events are no-ops: program order is execution
// order).  Blocks run one after another (or, on request, a few at a time on OS threads).  A kernel without barriers
// runs its threads one after another too; in one with barriers every CUDA thread is a fiber, and the scheduler knows
// three ways of waiting - the block barrier, the warp-level rendezvous (__syncwarp, mma.sync), a polling loop - with
// exited threads counting as arrived, as on the hardware.  __shared__ becomes a thread_local static (one block per OS
// thread at a time, so one copy is what a block sees).  Round-to-nearest intrinsics map to the plain IEEE operation
// (compile with -ffp-contract=off), __fma_rn to fma(); exp / sin / cos / log come from glibc, which legitimately
// differs from CUDA's libm by an ulp.  Memory is sequentially consistent: fences are no-ops, and nothing here can say
// anything about the device's memory model or its asynchronous proxy.
#ifndef COCONS_TEST_CUDA_EMUL_H
#define COCONS_TEST_CUDA_EMUL_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#ifdef COCONS_EMUL_TSAN
#define EMUL_NO_TSAN __attribute__((no_sanitize("thread")))
#else
#define EMUL_NO_TSAN
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct double2 {
  double x, y;
};

// host memory is "device" memory here: the handful of runtime calls the launch helpers of solve.cu make
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <class T>
cudaError_t cudaMalloc(T** p, size_t bytes) {
  *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
  return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = 148;
  return cudaSuccess;
}
template <class F>
cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return cudaSuccess;
}
template <class F>
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) {
  *n = 1;
  return cudaSuccess;
}
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) {  // the emulated "device"
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
constexpr cudaError_t cudaErrorMemoryAllocation = 2;
struct cudaDeviceProp {
  int major = 10, minor = 0, multiProcessorCount = 148;  // what check_device() of capi.cu asks for
};
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  *p = cudaDeviceProp();
  return cudaSuccess;
}
template <class T>
cudaError_t cudaMallocHost(T** p, size_t bytes) {
  *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void* p) {
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2D(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                cudaMemcpyKind) {
  for (size_t r = 0; r < height; ++r)
    std::memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.0f;
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
  *lo = 0, *hi = -1;
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) {
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }  // synchronous launches:
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }  // program order
inline cudaError_t cudaLaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) {
  std::fprintf(stderr, "host emulation: cooperative launches are rewritten by build.py (grid of one block)\n");
  std::abort();
}

// ---- context switch between the fibers of a block ---------------------------------------------------------------
// x86-64: a dozen instructions (callee-saved registers + stack pointer; swapcontext() would make two signal-mask
// system calls per switch, and an emulated factorisation makes tens of millions of switches).  Elsewhere: ucontext.
#if defined(__x86_64__)
extern "C" void emul_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.hidden emul_switch
.globl emul_switch
.type emul_switch,@function
emul_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emul_switch,.-emul_switch
)");
namespace emul {
struct Context {
  void* sp = nullptr;
  // a fresh context that starts in entry() on the given stack (entry must not return)
  EMUL_NO_TSAN void prepare(char* stack, size_t size, void (*entry)()) {
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + size) & ~uintptr_t(15);
    void** p = reinterpret_cast<void**>(top - 64);  // six saved registers, the entry address, one pad slot
    for (int k = 0; k < 6; ++k) p[k] = nullptr;
    p[6] = reinterpret_cast<void*>(entry);
    p[7] = nullptr;
    sp = p;
  }
  EMUL_NO_TSAN static void swap(Context& from, Context& to) { emul_switch(&from.sp, to.sp); }
};
}  // namespace emul
#else
namespace emul {
struct Context {
  ucontext_t uc;
  void prepare(char* stack, size_t size, void (*entry)()) {
    getcontext(&uc);
    uc.uc_stack.ss_sp = stack, uc.uc_stack.ss_size = size, uc.uc_link = nullptr;
    makecontext(&uc, entry, 0);
  }
  static void swap(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
};
}  // namespace emul
#endif

namespace emul {

struct ThreadCtx {
  dim3 tid, bid, bdim, gdim;
  void* dyn_smem = nullptr;  // the launch's dynamic shared memory (one block at a time)
  size_t dyn_bytes = 0;
  bool in_block = false;     // inside a block that runs on fibers (has barriers)
};
inline thread_local ThreadCtx ctx;
inline long launches = 0, barrier_launches = 0;

// ---- optional race check (-DCOCONS_EMUL_TSAN, built with -fsanitize=thread: tools/emul_racecheck.sh) ---------------
// Every CUDA thread becomes a ThreadSanitizer fiber, switched WITHOUT synchronisation, so the only happens-before
// edges TSan sees are the ones the kernel itself establishes: __syncthreads() / __syncwarp() / mma.sync (release by
// every arriving thread, acquire by every leaving one), mbarrier arrive / expect_tx / complete_tx (release) ->
// successful try_wait (acquire), the __atomic builtins behind atomicAdd & co., and block / launch boundaries.  Two
// threads of a block touching the same shared or global location without such an edge are reported as a data race -
// what compute-sanitizer's racecheck reports on the device.  The scheduler's own bookkeeping is not instrumented.
#ifdef COCONS_EMUL_TSAN
extern "C" {
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void* __tsan_get_current_fiber(void);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
inline void hb_release(void* a) { __tsan_release(a); }
inline void hb_acquire(void* a) { __tsan_acquire(a); }
#else
inline void hb_release(void*) {}
inline void hb_acquire(void*) {}
#endif

// One block of a kernel with barriers: every CUDA thread is a fiber of the calling OS thread.  The
// scheduler resumes the runnable fibers one after another; a fiber runs until it has to wait:
//   WAIT_BLOCK  __syncthreads(): released when every live thread of the block waits there
//   WAIT_WARP   __syncwarp(), the warp-collective mma.sync and __shfl_sync(mask, ...): released when every live lane
//               named in the mask waits there with the same mask
//   (RUNNABLE)  a polling loop (mbarrier.try_wait) gives the turn away and is resumed in the next pass
// A thread that has left the kernel counts as arrived, as on the hardware.  Nothing runnable while threads are
// alive is a deadlock and ends the test with a message.  Deterministic; a barrier costs `block` context switches.
class FiberBlock {
 public:
  enum State : unsigned char { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };
  static constexpr size_t kStack = 256 * 1024;
  EMUL_NO_TSAN explicit FiberBlock(unsigned nt)
      : nt_(nt), fibers_(nt), tctx_(nt), state_(nt, DONE), ops_(nt, 0), xa_(((nt + 31) / 32) * 64), xb_(xa_.size()),
        warp_sync_((nt + 31) / 32, 0), wmask_(nt, 0), shc_(nt, 0), shv_(2 * (size_t)nt, 0.0) {
    stacks_ = static_cast<char*>(std::malloc(kStack * nt));
    if (!stacks_) std::abort();
    // raw views for the scheduler's (uninstrumented) code: std::vector's accessors are functions of their own
    fib_ = fibers_.data(), tc_ = tctx_.data(), st_ = state_.data(), op_ = ops_.data(), ws_ = warp_sync_.data();
    xa_p_ = xa_.data(), xb_p_ = xb_.data();
    wm_ = wmask_.data(), sc_ = shc_.data(), sv_ = shv_.data();
#ifdef COCONS_EMUL_TSAN
    tsan_main_ = __tsan_get_current_fiber();
    tsan_.resize(nt);
    ts_ = tsan_.data();
    for (unsigned t = 0; t < nt; ++t) ts_[t] = __tsan_create_fiber(0);
#endif
  }
  EMUL_NO_TSAN ~FiberBlock() {
#ifdef COCONS_EMUL_TSAN
    for (void* f : tsan_) __tsan_destroy_fiber(f);
#endif
    std::free(stacks_);
  }
  FiberBlock(const FiberBlock&) = delete;

  template <class Body>
  EMUL_NO_TSAN void run(const Body& body, dim3 bid, dim3 block, dim3 grid, void* dyn, size_t dyn_bytes) {
    hb_release(&launch_sync_);  // what the host (and the previous block) wrote is visible to every thread of this block
    body_ = [](void* b) { (*static_cast<const Body*>(b))(); };
    body_arg_ = const_cast<Body*>(&body);
    current_block_ = this;
    for (unsigned t = 0; t < nt_; ++t) {
      tc_[t].tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      tc_[t].bid = bid, tc_[t].bdim = block, tc_[t].gdim = grid, tc_[t].dyn_smem = dyn;
      tc_[t].dyn_bytes = dyn_bytes;
      tc_[t].in_block = true;
      st_[t] = RUNNABLE;
      op_[t] = 0, sc_[t] = 0;
      fib_[t].prepare(stacks_ + kStack * t, kStack, &FiberBlock::trampoline);
    }
    unsigned live = nt_;
    or_acc_ = 0;
    while (live) {
      // releases
      bool all_at_block = true;
      for (unsigned t = 0; t < nt_; ++t)
        if (st_[t] == RUNNABLE || st_[t] == WAIT_WARP) all_at_block = false;
      if (all_at_block) {
        for (unsigned t = 0; t < nt_; ++t)
          if (st_[t] == WAIT_BLOCK) st_[t] = RUNNABLE;
        or_result_ = or_acc_;  // what __syncthreads_or() of the phase just completed returns
        or_acc_ = 0;
      }
      for (unsigned w0 = 0; w0 < nt_; w0 += 32) {
        const unsigned w1 = w0 + 32 < nt_ ? w0 + 32 : nt_;
        for (unsigned t = w0; t < w1; ++t) {
          if (st_[t] != WAIT_WARP) continue;
          const unsigned mask = wm_[t];
          bool all = true;  // every live lane named in the mask waits with this very mask
          for (unsigned l = w0; l < w1 && all; ++l)
            if (((mask >> (l - w0)) & 1u) && st_[l] != DONE && !(st_[l] == WAIT_WARP && wm_[l] == mask)) all = false;
          if (all)
            for (unsigned l = w0; l < w1; ++l)
              if (((mask >> (l - w0)) & 1u) && st_[l] == WAIT_WARP) st_[l] = RUNNABLE;
        }
      }
      bool progressed = false;
      for (unsigned t = 0; t < nt_; ++t) {
        if (st_[t] != RUNNABLE) continue;
        progressed = true;
        cur_ = t;
        ctx = tc_[t];
        switch_to(t);
        if (st_[t] == DONE) --live;
      }
      if (!progressed && live) {
        std::fprintf(stderr, "host emulation: deadlock - %u live thread(s) of block (%u,%u,%u), none runnable "
                     "(a barrier that not every thread of the block / warp reaches)\n", live, bid.x, bid.y, bid.z);
        std::abort();
      }
    }
    ctx.in_block = false;
    current_block_ = nullptr;
    hb_acquire(&launch_sync_);  // ... and what the block wrote is visible to the host and to the next block
  }
  // from a fiber: wait in `state` (WAIT_BLOCK / WAIT_WARP), or give the turn away (RUNNABLE: a polling loop)
  EMUL_NO_TSAN int wait(State state, int pred = 0, unsigned mask = 0xFFFFFFFFu) {
    if (pred) or_acc_ = 1;
    const unsigned me = cur_;
    wm_[me] = mask;
    void* sync = state == WAIT_BLOCK ? static_cast<void*>(&block_sync_)
                                     : (state == WAIT_WARP ? static_cast<void*>(&ws_[me >> 5]) : nullptr);
    if (sync) hb_release(sync);
    st_[me] = state;
    switch_to_main(me);
    if (sync) hb_acquire(sync);
    return or_result_;
  }
  EMUL_NO_TSAN void poll() {
    if (++polls_ > (1ull << 33)) {
      std::fprintf(stderr, "host emulation: a polling loop does not end\n");
      std::abort();
    }
    wait(RUNNABLE);
  }
  // warp-collective exchange (mma.sync operands): every lane deposits, the warp meets, every lane reads.  Two
  // slots by the parity of the lane's collective-operation count: a lane can be at most one operation ahead of
  // the slowest lane of its warp (it has to meet the warp again), so the slot being read is never overwritten
  EMUL_NO_TSAN void exchange(double a, double b, const double*& all_a, const double*& all_b) {
    const unsigned t = cur_, w = t >> 5, lane = t & 31, par = op_[t]++ & 1;
    double* sa = xa_p_ + (w * 2 + par) * 32;
    double* sb = xb_p_ + (w * 2 + par) * 32;
    sa[lane] = a, sb[lane] = b;
    wait(WAIT_WARP);
    all_a = sa, all_b = sb;
  }
  // __shfl_sync(mask, v, src): the lanes of the mask meet; each reads lane src's deposit.  The lanes of a mask group
  // are assumed to have done the same number of shuffles (they shuffle together), which makes the reader's parity the
  // writer's; a lane can run at most one shuffle ahead of its group, into the other slot
  EMUL_NO_TSAN double shuffle(unsigned mask, double v, int src) {
    const unsigned t = cur_, w0 = t & ~31u, par = sc_[t]++ & 1;
    sv_[2 * t + par] = v;
    wait(WAIT_WARP, 0, mask);
    return sv_[2 * (w0 + (unsigned)(src & 31)) + par];
  }
  unsigned lane() const { return cur_ & 31; }
  static FiberBlock* current() { return current_block_; }

 private:
  EMUL_NO_TSAN static void trampoline() {
    FiberBlock* b = current_block_;
    hb_acquire(&b->launch_sync_);
    b->body_(b->body_arg_);
    hb_release(&b->launch_sync_);
    b->st_[b->cur_] = DONE;
    b->switch_to_main(b->cur_);  // never resumed
    std::abort();
  }
  EMUL_NO_TSAN void switch_to(unsigned t) {
#ifdef COCONS_EMUL_TSAN
    __tsan_switch_to_fiber(ts_[t], 1);  // 1 = no synchronisation implied by the switch
#endif
    Context::swap(main_, fib_[t]);
  }
  EMUL_NO_TSAN void switch_to_main(unsigned me) {
#ifdef COCONS_EMUL_TSAN
    __tsan_switch_to_fiber(tsan_main_, 1);
#endif
    Context::swap(fib_[me], main_);
  }
  static inline thread_local FiberBlock* current_block_ = nullptr;
  unsigned nt_, cur_ = 0;
  std::vector<Context> fibers_;
  std::vector<ThreadCtx> tctx_;
  std::vector<unsigned char> state_;
  std::vector<unsigned> ops_;
  std::vector<double> xa_, xb_;
  Context main_;
  char* stacks_ = nullptr;
  void (*body_)(void*) = nullptr;
  void* body_arg_ = nullptr;
  int or_acc_ = 0, or_result_ = 0;
  unsigned long long polls_ = 0;
  char block_sync_ = 0, launch_sync_ = 0;  // addresses for the happens-before annotations of the race check
  std::vector<char> warp_sync_;
  std::vector<unsigned> wmask_, shc_;
  std::vector<double> shv_;
  unsigned *wm_ = nullptr, *sc_ = nullptr;
  double* sv_ = nullptr;
  Context* fib_ = nullptr;
  ThreadCtx* tc_ = nullptr;
  unsigned char* st_ = nullptr;
  unsigned* op_ = nullptr;
  char* ws_ = nullptr;
  double *xa_p_ = nullptr, *xb_p_ = nullptr;
#ifdef COCONS_EMUL_TSAN
  void* tsan_main_ = nullptr;
  std::vector<void*> tsan_;
  void** ts_ = nullptr;
#endif
};

// One launch at a time in the whole process: the "shared memory" statics, the thread context and the fiber scheduler
// are global, and the product's host code may launch from several threads (DenseLikelihoodPool)
inline std::mutex launch_lock;

// Blocks of a kernel with barriers that run AT THE SAME TIME (default 1: one after another).  With g > 1 a launch
// spreads its blocks over g OS threads (block b on thread b mod g), each with its own fibers, "shared memory"
// (thread_local statics) and dynamic shared memory - real, preemptive concurrency between blocks, for the kernels whose
// blocks talk to each other through global memory (the dataflow forward substitution: tickets, front counter,
// partial sums).  Set through emu_set_concurrent_blocks() of driver.cpp.
inline int concurrent_blocks = 1;

template <class Body>
void launch(dim3 grid, dim3 block, bool has_barrier, size_t smem_bytes, Body&& body) {
  std::lock_guard<std::mutex> one_at_a_time(launch_lock);
  ++launches;
  const size_t smem_doubles = (smem_bytes + sizeof(double) - 1) / sizeof(double);  // exactly what was asked for
  const unsigned nt = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  auto block_id = [&](unsigned b) { return dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)); };
  if (!has_barrier) {
    std::vector<double> smem(smem_doubles);
    for (unsigned b = 0; b < nblocks; ++b)
      for (unsigned t = 0; t < nt; ++t) {
        ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        ctx.bid = block_id(b), ctx.bdim = block, ctx.gdim = grid, ctx.dyn_smem = smem.data(), ctx.in_block = false;
        ctx.dyn_bytes = smem_doubles * sizeof(double);
        body();
      }
    return;
  }
  ++barrier_launches;
  const unsigned workers = (unsigned)concurrent_blocks < nblocks ? (unsigned)concurrent_blocks : nblocks;
  auto work = [&](unsigned w, unsigned stride) {
    std::vector<double> smem(smem_doubles);
    FiberBlock fb(nt);
    for (unsigned b = w; b < nblocks; b += stride)
      fb.run(body, block_id(b), block, grid, smem.data(), smem_doubles * sizeof(double));
  };
  if (workers <= 1) {
    work(0, 1);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < workers; ++w) pool.emplace_back(work, w, workers);
  for (auto& th : pool) th.join();
}

}  // namespace emul

#define threadIdx (emul::ctx.tid)
#define blockIdx (emul::ctx.bid)
#define blockDim (emul::ctx.bdim)
#define gridDim (emul::ctx.gdim)

inline emul::FiberBlock* emul_block() {
  if (!emul::ctx.in_block || !emul::FiberBlock::current()) {
    std::fprintf(stderr, "host emulation: a barrier in a kernel launched without barrier support\n");
    std::abort();
  }
  return emul::FiberBlock::current();
}
inline int emul_barrier(int pred) { return emul_block()->wait(emul::FiberBlock::WAIT_BLOCK, pred); }
inline void __syncthreads() { emul_barrier(0); }
inline void __syncwarp() { emul_block()->wait(emul::FiberBlock::WAIT_WARP); }

inline int __syncthreads_or(int pred) { return emul_barrier(pred); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) {
  if (emul::concurrent_blocks > 1) std::this_thread::yield();
}
template <class T>
T __ldg(const T* p) {
  return *p;
}
template <class T>
T __ldcg(const T* p) {
  return *p;
}
template <class T>
T __ldcv(const T* p) {
  return *p;
}
template <class T>
T __ldcs(const T* p) {
  return *p;
}
template <class T>
void __stcs(T* p, T v) {
  *p = v;
}
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double __shfl_sync(unsigned mask, double v, int src) { return emul_block()->shuffle(mask, v, src); }
// offset of a shared-memory object inside the launch's dynamic shared memory (what the 32-bit shared-window
// address is used for in the kernels: mbarrier and bulk-copy operands)
// ... or, for an object in a static __shared__ array (the mbarriers of potrf_tile_kernel), a handle: bit 31 + an index
// into a per-thread table of such objects.  No arithmetic is done on those handles by the kernels.
namespace emul {
inline thread_local std::vector<const void*> static_shared_objects;
}
inline size_t __cvta_generic_to_shared(const void* p) {
  const char* base = static_cast<const char*>(emul::ctx.dyn_smem);
  const char* q = static_cast<const char*>(p);
  if (base && q >= base && q < base + emul::ctx.dyn_bytes) return (size_t)(q - base);
  auto& tab = emul::static_shared_objects;
  for (size_t i = 0; i < tab.size(); ++i)
    if (tab[i] == p) return 0x80000000u | i;
  tab.push_back(p);
  return 0x80000000u | (tab.size() - 1);
}

// ---- stand-ins for the inline-PTX helper functions of the kernels (tests/host_emul/build.py replaces the body of a
//      __device__ helper made of one asm statement by a call to emul::ptx_<name> with the same arguments) ---------
namespace emul {
// mma.sync.aligned.m8n8k4.row.col.f64: a = A[g][c4], b = B[c4][g], {d0, d1} = D[g][2 c4 + {0, 1}], g = lane / 4,
// c4 = lane % 4; the four products of an entry are accumulated in k order (the hardware's order is not documented;
// the tests compare at a tolerance)
inline void ptx_dmma884(double& d0, double& d1, double a, double b) {
  FiberBlock* fb = emul_block();
  const unsigned lane = fb->lane(), g = lane >> 2, c4 = lane & 3;
  const double *A, *B;
  fb->exchange(a, b, A, B);
  for (unsigned k = 0; k < 4; ++k) {
    d0 = std::fma(A[g * 4 + k], B[(2 * c4) * 4 + k], d0);
    d1 = std::fma(A[g * 4 + k], B[(2 * c4 + 1) * 4 + k], d1);
  }
}
// mbarrier in 8 bytes of shared memory: arrivals still expected in this phase, transaction bytes outstanding, phase
struct MBar {
  uint16_t expected, pending;
  int32_t tx : 31;
  uint32_t phase : 1;
};
static_assert(sizeof(MBar) == 8, "an mbarrier is one 64-bit word");
EMUL_NO_TSAN inline MBar* mbar_at(uint32_t off) {
  if (off & 0x80000000u)
    return reinterpret_cast<MBar*>(const_cast<void*>(static_shared_objects[off & 0x7FFFFFFFu]));
  return reinterpret_cast<MBar*>(static_cast<char*>(ctx.dyn_smem) + off);
}
EMUL_NO_TSAN inline void mbar_check(MBar* m) {
  if (m->pending == 0 && m->tx == 0) m->phase ^= 1, m->pending = m->expected;
}
EMUL_NO_TSAN inline void ptx_mbar_init(uint32_t bar, uint32_t count) {
  MBar* m = mbar_at(bar);
  m->expected = m->pending = (uint16_t)count, m->tx = 0, m->phase = 0;
}
EMUL_NO_TSAN inline void ptx_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  MBar* m = mbar_at(bar);
  hb_release(m);
  m->tx += (int32_t)bytes, m->pending -= 1;
  mbar_check(m);
}
EMUL_NO_TSAN inline void ptx_mbar_arrive(uint32_t bar) {
  MBar* m = mbar_at(bar);
  // COCONS_EMUL_DROP_HANDBACK=1 (self-test of the race check): pretend the consumers' arrive on the `empty` barrier
  // orders nothing - the producer's refill of a slot then races with the reads of that slot and must be reported
  static const bool drop = std::getenv("COCONS_EMUL_DROP_HANDBACK") != nullptr;
  if (!drop) hb_release(m);
  m->pending -= 1;
  mbar_check(m);
}
inline void ptx_mbar_arrive_after(uint32_t bar, double, double, uint32_t) { ptx_mbar_arrive(bar); }
EMUL_NO_TSAN inline void ptx_mbar_wait(uint32_t bar, uint32_t parity) {  // until the phase with this parity has completed
  while (mbar_at(bar)->phase == parity) emul_block()->poll();
  hb_acquire(mbar_at(bar));
}
// cp.async.bulk global -> shared with complete_tx on the barrier; synchronous here
EMUL_NO_TSAN inline void ptx_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  std::memcpy(static_cast<char*>(ctx.dyn_smem) + dst, src, bytes);
  MBar* m = mbar_at(bar);
  hb_release(m);
  m->tx -= (int32_t)bytes;
  mbar_check(m);
}
inline unsigned ptx_ld_acquire_u32(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void ptx_st_release_u32(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long ptx_ld_relaxed_u64(const double* p) {
  return __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_RELAXED);
}
}  // namespace emul
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int expected, int desired) {
  __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;
}
inline double __longlong_as_double(long long v) {
  double d;
  std::memcpy(&d, &v, sizeof d);
  return d;
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
