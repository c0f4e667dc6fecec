// specter_b200 -- common device/host helpers.
//
// All kernels are hand-written for sm_100a (B200).  The SX_EMU switch exists only
// so that tests/emu can run the *same kernel source* on CPU threads in a
// container without a GPU; the shipped library is always built by nvcc.
#pragma once
#ifndef SX_EMU
#include <cuda_runtime.h>
#define SX_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char sx_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(sx_dyn_smem_raw)
#define SX_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>

namespace sx {

typedef double2 cplx;

// ---- cp.async (LDGSTS) 16-byte copies global -> shared, used for software prefetch --------------
#ifndef SX_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#else
// (tests/emu/cuda_emu.h: with SX_EMU_ADVERSARIAL & 4 the copy lands at the issuing thread's wait_group)
inline void cp_async16(void* smem_dst, const void* gsrc) {
  if (emu::t_worker->adv & 4) emu::t_worker->cps[emu::flat_tid()].push_back([=]() { memcpy(smem_dst, gsrc, 16); });
  else memcpy(smem_dst, gsrc, 16);
}
inline void cp_async_commit() {}
inline void cp_async_wait_all() {
  if (emu::t_worker->adv & 4) emu::flush(emu::t_worker->cps[emu::flat_tid()]);
}
#endif

// ---- bulk asynchronous copies (TMA, cp.async.bulk) global -> shared, completion on an mbarrier ------------
// One elected thread issues a copy of a whole contiguous row block; the data lands in shared memory without
// passing through the SM's load/store pipe or the register file, so a ring of stages keeps tens of KB per SM
// in flight at two warps per CTA.  bytes: multiple of 16; both addresses 16-byte aligned.
#ifndef SX_EMU
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// several copies completing on one barrier: expect the total first, then issue the pieces
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_piece(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// pull a contiguous range into L2 ahead of the loads that will use it (no shared memory, no registers)
__device__ __forceinline__ void l2_prefetch(const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SX_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SX_DONE_%=;\n"
      "bra SX_WAIT_%=;\n"
      "SX_DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
#else
// default emulation: the elected thread runs first and copies synchronously, waits are no-ops; the adversarial modes of
// tests/emu/cuda_emu.h track the barrier phase and complete the copies as late as the program allows
inline void mbar_init(unsigned long long* bar, unsigned) {
  if (emu::t_worker->adv) emu::bar_init(bar);
}
inline void mbar_init_fence() {}
inline void mbar_expect(unsigned long long* bar, unsigned bytes) {
  if (emu::t_worker->adv) emu::bar_expect(bar, bytes);
}
inline void bulk_load_piece(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  if (emu::t_worker->adv) emu::bar_issue(bar, bytes, [=]() { memcpy(smem_dst, gsrc, bytes); });
  else memcpy(smem_dst, gsrc, bytes);
}
inline void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  mbar_expect(bar, bytes);
  bulk_load_piece(smem_dst, gsrc, bytes, bar);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  if (emu::t_worker->adv) emu::bar_wait(bar, parity);
}
inline void l2_prefetch(const void*, unsigned) {}
#endif

__host__ __device__ __forceinline__ cplx cmake(double x, double y) { return make_double2(x, y); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
// i*a and -i*a
__host__ __device__ __forceinline__ cplx cmuli(cplx a) { return cmake(-a.y, a.x); }
__host__ __device__ __forceinline__ cplx cmulmi(cplx a) { return cmake(a.y, -a.x); }
// a + s*b (real s)
__host__ __device__ __forceinline__ cplx caxpy(double s, cplx b, cplx a) {
  return cmake(fma(s, b.x, a.x), fma(s, b.y, a.y));
}

__host__ __device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- error plumbing (host) -------------------------------------------------
void set_error(const std::string& msg);
#define SX_CUDA_CHECK(expr)                                                          \
  do {                                                                               \
    cudaError_t sx_e_ = (expr);                                                      \
    if (sx_e_ != cudaSuccess) {                                                      \
      sx::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(sx_e_) +   \
                    " at " + __FILE__ + ":" + std::to_string(__LINE__));             \
      return 1;                                                                      \
    }                                                                                \
  } while (0)
#define SX_KERNEL_CHECK() SX_CUDA_CHECK(cudaGetLastError())
#define SX_REQUIRE(cond, msg)                        \
  do {                                               \
    if (!(cond)) {                                   \
      sx::set_error(std::string("[ERROR] ") + msg);  \
      return 1;                                      \
    }                                                \
  } while (0)

}  // namespace sx
