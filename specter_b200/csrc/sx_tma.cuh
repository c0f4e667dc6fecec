// Tensor-map (TMA) copies between the [row][NP] shared-memory tiles of the fused tile kernels and the
// strided side of their global layouts.  A tile of NP adjacent lines is a 3-D box {NP elements, <=256 rows, 1}
// of a tensor whose fastest dimension is the line index: one elected thread moves the whole tile with one or
// two instructions, the 64-byte pieces never pass through the SM's load/store pipe (where a warp-wide access
// to 8 different 128-byte lines costs 8 wavefronts), and boxes that stick out of the tensor (the continuation
// rows above the physical region) are clipped by the hardware.
//
// Elements are complex128; the maps are encoded in doubles (there is no 16-byte TMA element type), so the
// fastest coordinate is 2 * element index.
#pragma once
#include "sx_common.cuh"

#ifndef SX_EMU
#include <cuda.h>
#endif

namespace sx {

#ifndef SX_EMU
struct TmaMap {
  CUtensorMap m;
};
#define SX_GRID_CONSTANT __grid_constant__

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// smem tile <- box at (c0 doubles, c1 rows, c2 planes); completes on `bar` with the full box byte count
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const TmaMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
      ::"r"(d), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const TmaMap* map, const void* smem_src, int c0, int c1, int c2) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the source tiles of all committed stores have been read (the shared memory may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
#else
// CPU emulation (tests only): the same box semantics with synchronous copies and clipping
struct TmaMap {
  char* base;
  unsigned long long dim[3], stride[3];  // dim[0] in doubles; stride in bytes (stride[0] = 8)
  unsigned box[3];
};
#define SX_GRID_CONSTANT
inline void fence_proxy_async() {}
inline void tma_load_3d_now(void* smem_dst, const TmaMap* m, int c0, int c1, int c2) {
  char* d = (char*)smem_dst;
  for (unsigned k = 0; k < m->box[2]; ++k)
    for (unsigned r = 0; r < m->box[1]; ++r)
      for (unsigned e = 0; e < m->box[0]; ++e) {
        const unsigned long long x = c0 + e, y = c1 + r, z = c2 + k;
        double val = 0.0;
        if (x < m->dim[0] && y < m->dim[1] && z < m->dim[2]) val = *(const double*)(m->base + x * 8 + y * m->stride[1] + z * m->stride[2]);
        *(double*)(d + (((size_t)k * m->box[1] + r) * m->box[0] + e) * 8) = val;
      }
}
inline void tma_load_3d(void* smem_dst, const TmaMap* m, int c0, int c1, int c2, unsigned long long* bar) {
  if (!emu::t_worker->adv) return tma_load_3d_now(smem_dst, m, c0, c1, c2);
  const unsigned bytes = m->box[0] * m->box[1] * m->box[2] * 8u;     // the full box counts, clipped or not
  const TmaMap mm = *m;                                               // the map may live in the issuing thread's frame
  emu::bar_issue(bar, bytes, [=]() { tma_load_3d_now(smem_dst, &mm, c0, c1, c2); });
}
inline void tma_store_3d_now(const TmaMap* m, const void* smem_src, int c0, int c1, int c2) {
  const char* s = (const char*)smem_src;
  for (unsigned k = 0; k < m->box[2]; ++k)
    for (unsigned r = 0; r < m->box[1]; ++r)
      for (unsigned e = 0; e < m->box[0]; ++e) {
        const unsigned long long x = c0 + e, y = c1 + r, z = c2 + k;
        if (x < m->dim[0] && y < m->dim[1] && z < m->dim[2])
          *(double*)(m->base + x * 8 + y * m->stride[1] + z * m->stride[2]) = *(const double*)(s + (((size_t)k * m->box[1] + r) * m->box[0] + e) * 8);
      }
}
// adversarial mode 4: the store reads its shared-memory source when the issuing thread waits for its bulk group
inline void tma_store_3d(const TmaMap* m, const void* smem_src, int c0, int c1, int c2) {
  if (emu::t_worker->adv & 4) {
    const TmaMap mm = *m;
    emu::t_worker->stores[emu::flat_tid()].push_back([=]() { tma_store_3d_now(&mm, smem_src, c0, c1, c2); });
  } else {
    tma_store_3d_now(m, smem_src, c0, c1, c2);
  }
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() {
  if (emu::t_worker->adv & 4) emu::flush(emu::t_worker->stores[emu::flat_tid()]);
}
inline void tma_store_wait_all() { tma_store_wait_read(); }
#endif

// host: tensor of complex128 with extents (n0 elements [fastest, contiguous], n1 rows, n2 planes), row / plane pitch in
// elements, box = {box0 elements, box1 rows, 1}
int tma_encode(TmaMap* out, const void* base, size_t n0, size_t n1, size_t n2, size_t pitch1, size_t pitch2, int box0, int box1);

}  // namespace sx
