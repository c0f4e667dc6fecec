// Minimal CUDA-on-CPU shim.  TEST INFRASTRUCTURE ONLY.
//
// Compiling specter_b200/csrc/*.cu with g++ -DSX_EMU -include tests/emu/cuda_emu.h
// produces tests/emu/_build/libspecter_emu.so: the *same kernel source* executed by
// user-level fibers (one per CUDA thread; __syncthreads is a yield).  The non-GPU
// test-suite uses it to check kernel index math / barrier structure in a
// container that has no GPU.  The product package never loads it (specter_b200
// only ever dlopens libspecter_b200.so and fails loudly without a CUDA device).
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <setjmp.h>
#include <ucontext.h>
#include <unordered_map>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__ static

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef int cudaError_t;
typedef void* cudaStream_t;
struct emu_event {
  std::chrono::steady_clock::time_point t;
  unsigned long long recorded = 0, done = 0;   // ticket of the last cudaEventRecord enqueued / executed (deferred streams)
};
typedef emu_event* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };

namespace emu {
// One OS worker per block group; the CUDA threads of a block are user-level fibers (ucontext)
// scheduled round-robin, so __syncthreads() is a yield: every fiber runs up to its next barrier
// before any fiber passes it.
//
// SX_EMU_ADVERSARIAL (bit mask, default 0) turns the emulation into a race check of the kernel sources:
//   1  the fibers of a block run in REVERSE order between barriers        } a result that depends on the order in which
//   2  ... in a pseudo-random order, re-drawn for every barrier phase     } the threads run has a missing __syncthreads
//   4  asynchronous copies complete as LATE as the program allows: cp.async.bulk / tensor loads land when the first
//      thread gets through mbarrier.try_wait on their barrier, cp.async 16-byte copies at the issuing thread's
//      wait_group, tensor-map STORES read their shared-memory source at the issuing thread's wait_group(.read) -- data
//      consumed before its wait, or a store source overwritten before wait_group.read, changes the result.
//   8  streams are lazy queues (see "runtime subset" below): work that no event orders before its consumer has not run
//      when the consumer does;
//  16  (with 8) the streams the library marks as communication / copy streams run as EARLY as their event waits allow
//      while the compute stream stays lazy: a buffer overwritten before its readers are done changes the result.
// With any bit set the mbarrier phase is tracked (expect_tx / complete_tx bytes) and mbarrier waits really wait, so the
// producer thread need not run first.
// Context switches: a fiber is ENTERED through makecontext / swapcontext (which sets up its stack) and from then on
// switched with _setjmp / _longjmp, which do not save and restore the signal mask (two system calls per swapcontext: the
// bulk of the emulation's run time).  SX_EMU_UCONTEXT_ONLY keeps swapcontext throughout (sanitizer builds).
struct Fiber {
  ucontext_t ctx;
  jmp_buf jb;
  bool started = false;
  bool done;
  bool spinning = false;
};
struct BarState {
  unsigned completed = 0, expected = 0, issued = 0;
};
struct Worker {
  ucontext_t main;
  jmp_buf main_jb;
  std::vector<Fiber> fibers;
  std::unique_ptr<char[]> stacks;   // not value-initialised: only the pages a fiber touches are ever mapped
  size_t stack_bytes = 0;
  int current = 0;
  const std::function<void()>* body = nullptr;
  dim3 block;
  char* smem = nullptr;
  std::vector<double> shfl;
  int adv = 0;
  unsigned long long rng = 0x9E3779B97F4A7C15ull;
  std::unordered_map<const void*, BarState> bars;
  unsigned long long events = 0;                                      // mbarrier issues + completions (deadlock detection)
  std::vector<std::pair<const void*, std::function<void()>>> loads;   // deferred bulk / tensor loads, keyed by barrier
  std::vector<std::vector<std::function<void()>>> stores, cps;        // deferred tensor stores / cp.async per thread
};
inline int adv_mode() {
  static const int m = [] { const char* e = std::getenv("SX_EMU_ADVERSARIAL"); return e ? std::atoi(e) : 0; }();
  return m;
}
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local Worker* t_worker = nullptr;

inline void set_tid(Worker* w, unsigned t) {
  t_threadIdx.x = t % w->block.x;
  t_threadIdx.y = (t / w->block.x) % w->block.y;
  t_threadIdx.z = t / (w->block.x * w->block.y);
}
inline void to_main(Worker* w, Fiber& f) {       // from a fiber back to the scheduler
#ifdef SX_EMU_UCONTEXT_ONLY
  swapcontext(&f.ctx, &w->main);
#else
  if (_setjmp(f.jb) == 0) _longjmp(w->main_jb, 1);
#endif
}
inline void to_fiber(Worker* w, Fiber& f) {      // from the scheduler into a fiber, until it comes back
#ifdef SX_EMU_UCONTEXT_ONLY
  swapcontext(&w->main, &f.ctx);
#else
  if (_setjmp(w->main_jb) == 0) {
    if (!f.started) {
      f.started = true;
      swapcontext(&w->main, &f.ctx);             // first entry: onto the fiber's own stack
    } else {
      _longjmp(f.jb, 1);
    }
  }
#endif
}
inline void fiber_entry() {
  Worker* w = t_worker;
  (*w->body)();
  Fiber& f = w->fibers[w->current];
  f.done = true;
  to_main(w, f);
}
inline void yield() {
  Worker* w = t_worker;
  to_main(w, w->fibers[w->current]);
}
// a thread that waits for another thread of its block (mbarrier wait): it is resumed again WITHIN the same barrier phase
inline void spin_yield() {
  Worker* w = t_worker;
  w->fibers[w->current].spinning = true;
  to_main(w, w->fibers[w->current]);
}
inline void flush(std::vector<std::function<void()>>& q) {
  for (auto& fn : q) fn();
  q.clear();
}
inline void run_block(Worker* w, unsigned nthr) {
  for (unsigned t = 0; t < nthr; ++t) {
    Fiber& f = w->fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = w->stacks.get() + (size_t)t * w->stack_bytes;
    f.ctx.uc_stack.ss_size = w->stack_bytes;
    f.ctx.uc_link = nullptr;
    f.done = false;
    f.started = false;
    f.spinning = false;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  unsigned remaining = nthr;
  if (!w->adv) {
    while (remaining) {
      for (unsigned t = 0; t < nthr; ++t) {
        Fiber& f = w->fibers[t];
        if (f.done) continue;
        w->current = (int)t;
        set_tid(w, t);
        to_fiber(w, f);
        if (f.done) --remaining;
      }
    }
    return;
  }
  w->bars.clear();
  w->loads.clear();
  w->stores.assign(nthr, {});
  w->cps.assign(nthr, {});
  std::vector<unsigned> order(nthr), retry, again;
  while (remaining) {
    for (unsigned t = 0; t < nthr; ++t) order[t] = (w->adv & 1) ? nthr - 1 - t : t;
    if (w->adv & 2)
      for (unsigned t = nthr; t > 1; --t) {          // Fisher-Yates with xorshift64
        w->rng ^= w->rng << 13; w->rng ^= w->rng >> 7; w->rng ^= w->rng << 17;
        std::swap(order[t - 1], order[(unsigned)(w->rng % t)]);
      }
    retry.clear();
    for (unsigned t : order)
      if (!w->fibers[t].done) retry.push_back(t);
    while (!retry.empty()) {                          // one barrier phase: every live fiber up to its next barrier
      again.clear();
      const unsigned long long events0 = w->events;
      for (unsigned t : retry) {
        Fiber& f = w->fibers[t];
        f.spinning = false;
        w->current = (int)t;
        set_tid(w, t);
        to_fiber(w, f);
        if (f.done) {
          --remaining;
          flush(w->stores[t]);                        // the stores of a finished thread complete before the grid ends
        } else if (f.spinning) {
          again.push_back(t);
        }
      }
      // every fiber that is left in this phase waits on an mbarrier, and nothing was issued or completed during the pass:
      // the copies they wait for are never issued (on the GPU: bar.sync and mbarrier.try_wait waiting for each other)
      if (!again.empty() && again.size() == retry.size() && w->events == events0) {
        std::fprintf(stderr, "cuda_emu: deadlock -- %zu threads wait on an mbarrier whose copies are never issued\n", again.size());
        std::abort();
      }
      retry.swap(again);
    }
  }
}

// ---- mbarrier / asynchronous-copy bookkeeping of the adversarial modes ------------------------------------------------
inline void bar_init(const void* bar) { t_worker->bars[bar] = BarState(); }
inline void bar_expect(const void* bar, unsigned bytes) { t_worker->bars[bar].expected += bytes; }
inline void bar_complete_if_full(Worker* w, const void* bar, BarState& st) {
  if (st.expected > 0 && st.issued == st.expected) {
    for (size_t i = 0; i < w->loads.size();) {
      if (w->loads[i].first == bar) {
        w->loads[i].second();
        w->loads.erase(w->loads.begin() + (long)i);
      } else {
        ++i;
      }
    }
    st.completed++;
    st.issued = st.expected = 0;
    w->events++;
  }
}
inline void bar_issue(const void* bar, unsigned bytes, std::function<void()> copy) {
  Worker* w = t_worker;
  BarState& st = w->bars[bar];
  st.issued += bytes;
  w->events++;
  if (w->adv & 4) {
    w->loads.emplace_back(bar, std::move(copy));      // lands when a waiter gets through
  } else {
    copy();
    bar_complete_if_full(w, bar, st);
  }
}
inline void bar_wait(const void* bar, unsigned parity) {
  Worker* w = t_worker;
  for (;;) {
    BarState& st = w->bars[bar];
    if ((st.completed & 1u) != (parity & 1u)) return;
    if (w->adv & 4) {
      bar_complete_if_full(w, bar, st);
      if ((st.completed & 1u) != (parity & 1u)) return;
    }
    spin_yield();
  }
}
inline unsigned flat_tid() { return (unsigned)t_worker->current; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body_in) {
  const unsigned nthr = block.x * block.y * block.z;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  if (nthr == 0 || nblocks == 0) return;
  unsigned ngroups = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 8u));
  if ((unsigned long long)ngroups > nblocks) ngroups = (unsigned)nblocks;
  const std::function<void()> body = body_in;
  std::vector<std::thread> pool;
  pool.reserve(ngroups);
  for (unsigned g = 0; g < ngroups; ++g) {
    pool.emplace_back([&, g]() {
      Worker w;
      w.block = block;
      w.body = &body;
      w.stack_bytes = 64 * 1024;
      w.stacks.reset(new char[(size_t)nthr * w.stack_bytes]);
      w.fibers.resize(nthr);
      std::vector<char> smem(smem_bytes + 64, 0);
      w.smem = smem.data();
      w.shfl.assign((size_t)nthr * 2, 0.0);
      w.adv = adv_mode();
      w.rng ^= 0x100000001B3ull * (g + 1);
      t_worker = &w;
      t_blockDim = block;
      t_gridDim = grid;
      for (unsigned long long b = g; b < nblocks; b += ngroups) {
        t_blockIdx.x = (unsigned)(b % grid.x);
        t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
        t_blockIdx.z = (unsigned)(b / ((unsigned long long)grid.x * grid.y));
        run_block(&w, nthr);
      }
      t_worker = nullptr;
    });
  }
  for (auto& th : pool) th.join();
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
static inline void __syncthreads() { emu::yield(); }
#define SX_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::t_worker->smem)

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospi(double x, double* s, double* c) {
  *s = std::sin(3.14159265358979323846 * x);
  *c = std::cos(3.14159265358979323846 * x);
}
static inline double atomicAdd(double* addr, double v) {
  std::atomic_ref<double> r(*addr);
  double old = r.load();
  while (!r.compare_exchange_weak(old, old + v)) {}
  return old;
}
// Warp shuffles emulated through block-wide scratch: valid only when every thread
// of the block executes the call (which is how the kernels use them).
static inline double __shfl_xor_sync(unsigned, double v, int lanemask) {
  unsigned tid = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  emu::t_worker->shfl[tid] = v;
  __syncthreads();
  unsigned nthr = blockDim.x * blockDim.y * blockDim.z;
  unsigned src = tid ^ (unsigned)lanemask;
  double r = (src < nthr) ? emu::t_worker->shfl[src] : v;
  __syncthreads();
  return r;
}
static inline double __shfl_down_sync(unsigned, double v, int delta) {
  unsigned tid = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  emu::t_worker->shfl[tid] = v;
  __syncthreads();
  unsigned nthr = blockDim.x * blockDim.y * blockDim.z;
  unsigned src = tid + (unsigned)delta;
  double r = (src < nthr && (src / 32) == (tid / 32)) ? emu::t_worker->shfl[src] : v;
  __syncthreads();
  return r;
}

// ---- runtime subset -------------------------------------------------------
// Streams.  Default: every operation runs at the call, in program order.  With SX_EMU_ADVERSARIAL & 8 the streams are
// queues that run as LATE and as LITTLE as the program allows: an operation executes only when the host waits for it
// (cudaStreamSynchronize, cudaEventSynchronize, a copy to pageable host memory, cudaFree) or when an operation the host
// waits for depends on it through cudaStreamWaitEvent.  Work of another stream that no event orders before its consumer
// has then NOT run when the consumer does: a missing event wait changes the result.  Pageable host memory follows the
// CUDA rules (sources are staged at the call, copies into it are synchronous); page-locked memory (cudaMallocHost) is
// read and written when the copy executes, so a host buffer reused too early shows as well.
namespace emu {
struct StreamOp {
  std::function<void()> fn;
  emu_event* ev = nullptr;
  unsigned long long ticket = 0;
  int kind = 0;                      // 0 work, 1 record ev/ticket, 2 wait for ev/ticket
};
struct Stream {
  std::deque<StreamOp> q;
  bool eager = false;                // SX_EMU_ADVERSARIAL & 16: runs as EARLY as its event waits allow (mark_eager)
};
struct Runtime {
  std::vector<Stream*> streams;
  std::map<const char*, size_t> pinned;   // page-locked host allocations: start -> bytes
  unsigned long long tickets = 0;
  bool deferred() const { return (adv_mode() & 8) != 0; }
  bool is_pinned(const void* p) const {
    auto it = pinned.upper_bound((const char*)p);
    if (it == pinned.begin()) return false;
    --it;
    return (const char*)p < it->first + it->second;
  }
  void run_until(emu_event* e, unsigned long long ticket) {   // until record `ticket` of e has executed
    if (e->done >= ticket) return;
    for (Stream* s : streams)
      for (const StreamOp& op : s->q)
        if (op.kind == 1 && op.ev == e && op.ticket == ticket) {
          while (e->done < ticket) step(s);
          return;
        }
    e->done = ticket;   // the record is gone with its stream: nothing left to wait for
  }
  void step(Stream* s) {                                      // executes the head of s (and what it waits for)
    StreamOp op = std::move(s->q.front());
    s->q.pop_front();
    if (op.kind == 2) {
      run_until(op.ev, op.ticket);
    } else if (op.kind == 1) {
      op.ev->t = std::chrono::steady_clock::now();
      if (op.ev->done < op.ticket) op.ev->done = op.ticket;
    } else {
      op.fn();
    }
    pump();
  }
  // eager streams: everything at their heads whose event waits are already satisfied runs NOW -- the earliest moment the
  // program allows, e.g. a copy that overwrites a buffer as soon as the events it waits for have fired
  bool pumping = false;
  void pump() {
    if (pumping || !(adv_mode() & 16)) return;
    pumping = true;
    for (bool progress = true; progress;) {
      progress = false;
      for (size_t i = 0; i < streams.size(); ++i) {
        Stream* s = streams[i];
        while (s->eager && !s->q.empty()) {
          StreamOp& head = s->q.front();
          if (head.kind == 2 && head.ev->done < head.ticket) break;   // blocked: the awaited record has not executed
          StreamOp op = std::move(head);
          s->q.pop_front();
          if (op.kind == 1) {
            op.ev->t = std::chrono::steady_clock::now();
            if (op.ev->done < op.ticket) op.ev->done = op.ticket;
          } else if (op.kind == 0) {
            op.fn();
          }
          progress = true;
        }
      }
    }
    pumping = false;
  }
  void drain(Stream* s) {
    while (!s->q.empty()) step(s);
  }
  void drain_all() {
    for (Stream* s : streams) drain(s);
  }
};
inline Runtime& rt() {
  static Runtime r;
  return r;
}
// run fn on the stream: now (default, or the null stream), or when the stream gets there
inline void enqueue(cudaStream_t st, std::function<void()> fn) {
  if (!st || !rt().deferred()) return fn();
  StreamOp op;
  op.fn = std::move(fn);
  static_cast<Stream*>(st)->q.push_back(std::move(op));
  rt().pump();
}
// the library marks its communication / copy streams: with SX_EMU_ADVERSARIAL & 16 they run as early as their event waits
// allow while the compute stream stays lazy -- the adversary of a buffer that is overwritten before its readers are done
inline void mark_eager(cudaStream_t st) {
  if (st) static_cast<Stream*>(st)->eager = true;
}
}  // namespace emu

static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) {
  emu::rt().drain_all();             // cudaFree waits for the device
  std::free(p);
  return 0;
}
static inline cudaError_t cudaMallocHost(void** p, size_t n) {
  *p = std::calloc(1, n ? n : 1);
  if (!*p) return 2;
  emu::rt().pinned[(const char*)*p] = n ? n : 1;
  return 0;
}
static inline cudaError_t cudaFreeHost(void* p) {
  emu::rt().drain_all();
  emu::rt().pinned.erase((const char*)p);
  std::free(p);
  return 0;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind kind, cudaStream_t st = 0) {
  emu::Runtime& r = emu::rt();
  if (!st || !r.deferred()) { std::memmove(d, s, n); return 0; }
  if (kind == cudaMemcpyHostToDevice && !r.is_pinned(s)) {          // pageable source: staged at the call
    auto stage = std::make_shared<std::vector<char>>((const char*)s, (const char*)s + n);
    emu::enqueue(st, [=]() { std::memcpy(d, stage->data(), n); });
  } else if (kind == cudaMemcpyDeviceToHost && !r.is_pinned(d)) {   // pageable destination: the call returns with the data
    r.drain(static_cast<emu::Stream*>(st));
    std::memmove(d, s, n);
  } else {
    emu::enqueue(st, [=]() { std::memmove(d, s, n); });
  }
  return 0;
}
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind kind,
                                            cudaStream_t st = 0) {
  emu::Runtime& r = emu::rt();
  auto copy = [=]() { for (size_t row = 0; row < h; ++row) std::memmove((char*)d + row * dp, (const char*)s + row * sp, w); };
  if (!st || !r.deferred()) { copy(); return 0; }
  if (kind == cudaMemcpyHostToDevice && !r.is_pinned(s)) {
    auto stage = std::make_shared<std::vector<char>>((const char*)s, (const char*)s + (h ? (h - 1) * sp + w : 0));
    emu::enqueue(st, [=]() { for (size_t row = 0; row < h; ++row) std::memcpy((char*)d + row * dp, stage->data() + row * sp, w); });
  } else if (kind == cudaMemcpyDeviceToHost && !r.is_pinned(d)) {
    r.drain(static_cast<emu::Stream*>(st));
    copy();
  } else {
    emu::enqueue(st, copy);
  }
  return 0;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = 0) {
  emu::enqueue(st, [=]() { std::memset(d, v, n); });
  return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  emu::Stream* st = new emu::Stream();
  emu::rt().streams.push_back(st);
  *s = st;
  return 0;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  if (s) emu::rt().drain(static_cast<emu::Stream*>(s));
  return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  if (!s) return 0;
  emu::Runtime& r = emu::rt();
  r.drain(static_cast<emu::Stream*>(s));
  for (size_t i = 0; i < r.streams.size(); ++i)
    if (r.streams[i] == s) { r.streams.erase(r.streams.begin() + (long)i); break; }
  delete static_cast<emu::Stream*>(s);
  return 0;
}
static inline cudaError_t cudaDeviceSynchronize() { emu::rt().drain_all(); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event(); return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event(); return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = 0) {
  emu::Runtime& r = emu::rt();
  const unsigned long long ticket = ++r.tickets;
  e->recorded = ticket;
  if (!st || !r.deferred()) {
    e->t = std::chrono::steady_clock::now();
    e->done = ticket;
    return 0;
  }
  emu::StreamOp op;
  op.kind = 1; op.ev = e; op.ticket = ticket;
  static_cast<emu::Stream*>(st)->q.push_back(std::move(op));
  r.pump();
  return 0;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  emu::rt().run_until(e, e->recorded);
  return 0;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  emu::rt().run_until(e, e->recorded);   // queued operations may still name it
  delete e;
  return 0;
}
// the stream waits for the record of e that is the latest one AT THIS CALL (a never recorded event: no wait)
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned = 0) {
  emu::Runtime& r = emu::rt();
  if (!r.deferred() || e->recorded == 0 || e->done >= e->recorded) return 0;
  if (!st) { r.run_until(e, e->recorded); return 0; }
  emu::StreamOp op;
  op.kind = 2; op.ev = e; op.ticket = e->recorded;
  static_cast<emu::Stream*>(st)->q.push_back(std::move(op));
  r.pump();
  return 0;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  emu::rt().run_until(a, a->recorded);
  emu::rt().run_until(b, b->recorded);
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return 0;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

// the arguments are EVALUATED at the call, like the parameter buffer of a real launch (a deferred launch must not read
// host variables later)
#define SX_LAUNCH(kernel, grid, block, smem, stream, ...)                                                    \
  do {                                                                                                       \
    const dim3 sx_g_ = (grid), sx_b_ = (block);                                                              \
    const size_t sx_s_ = (smem);                                                                             \
    auto sx_k_ = (kernel);                                                                                   \
    auto sx_a_ = std::make_tuple(__VA_ARGS__);                                                               \
    emu::enqueue((stream), [=]() { emu::launch(sx_g_, sx_b_, sx_s_, [&]() { std::apply(sx_k_, sx_a_); }); }); \
  } while (0)
