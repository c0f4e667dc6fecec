// Fused substep: the y-axis tile kernels and the z-inverse tile kernel (see sx_fused.cu for the pass structure).
#include "sx_fused.h"
#include "sx_tma.cuh"

namespace sx {

// ------------------------------------------------------------------------------------------
// zinv_tile: one tile = NP adjacent ky pencils of one kx.
// ------------------------------------------------------------------------------------------
struct ZinvArgs {
  const cplx* in;   // spectral (nz, ny, nxl)
  cplx* out0;       // IFFT_z(in), exchange layout
  cplx* out1;       // IFFT_z(i kz in) or nullptr
  const double* kz;
  const ZMap* zmap;
  int ny, nxl, nph;
};

template <int N, int NP, int MINB, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_zinv_tile(ZinvArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  TwRegs<N> twr;
  twr.load(tw, j);
  cplx* slot = smem + (size_t)NP * N + threadIdx.x;
  ZMap* zm = reinterpret_cast<ZMap*>(smem + (size_t)2 * NP * N);  // row table, read with one LDS.128
  for (int z = threadIdx.x; z < a.nph; z += NT) zm[z] = a.zmap[z];
  __syncthreads();
  const SIdxPencil si{p, NP};
  const int tiles_y = cdiv(a.ny, NP), ntiles = tiles_y * a.nxl;
  auto issue = [&](int t) {
    if (!PF) return;
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
    if (ky < a.ny) {
      const cplx* src = a.in + ((size_t)kxl * a.ny + ky) * N + j;
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async16(slot + k * NT, src + k * T);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  int t = blockIdx.x;
  if (t < ntiles) issue(t);
  for (; t < ntiles; t += gridDim.x) {
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
    const bool active = ky < a.ny;
    const cplx* src = a.in + ((size_t)kxl * a.ny + (active ? ky : 0)) * N + j;
    cplx v[8];
    if (PF) cp_async_wait_all();
    if (a.out1 != nullptr) {
      // derivative first: the slots (or L1) still hold this tile
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const cplx q = PF ? slot[k * NT] : (active ? src[k * T] : cmake(0.0, 0.0));
        const double kk = __ldg(&a.kz[j + k * T]);
        v[k] = cmake(-kk * q.y, kk * q.x);
      }
      fft_regs<N, 1>(v, j, smem, si, twr);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int z = j + k * T;
        if (active && z < a.nph) {
          const ZMap m = zm[z];
          a.out1[m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky] = v[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = PF ? slot[k * NT] : (active ? src[k * T] : cmake(0.0, 0.0));
    if (t + (int)gridDim.x < ntiles) issue(t + gridDim.x);
    fft_regs<N, 1>(v, j, smem, si, twr);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = j + k * T;
      if (active && z < a.nph) {
        const ZMap m = zm[z];
        a.out0[m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky] = v[k];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// yinv_tile: one tile = NP adjacent kx lines of one local z row; [kx][zl][ky] -> [zl][y][kx].
// ------------------------------------------------------------------------------------------
struct YinvArgs {
  const cplx* in;   // [kx][zl][ky]
  cplx* out0;       // IFFT_y(in)          [zl][y][kx]
  cplx* out1;       // IFFT_y(i ky in) or nullptr
  const double* ky;
  int nxh, nxp, nzf;   // nzf: z rows of the slab (line stride of `in`)
  int nzc;             // z rows to process (in / out already point at the first one)
};

template <int N, int NP, int MINB, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_yinv_tile(YinvArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  TwRegs<N> twr;
  twr.load(tw, j);
  cplx* slot = smem + (size_t)NP * N + threadIdx.x;
  const SIdxPencil si{p, NP};
  const int tiles_x = cdiv(a.nxp, NP), ntiles = tiles_x * a.nzc;
  auto issue = [&](int t) {
    if (!PF) return;
    const int kx = (t % tiles_x) * NP + p, zl = t / tiles_x;
    if (kx < a.nxh) {
      const cplx* src = a.in + ((size_t)kx * a.nzf + zl) * N + j;
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async16(slot + k * NT, src + k * T);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  int t = blockIdx.x;
  if (t < ntiles) issue(t);
  for (; t < ntiles; t += gridDim.x) {
    const int kx = (t % tiles_x) * NP + p, zl = t / tiles_x;
    const bool store = kx < a.nxp, load = kx < a.nxh;
    const size_t dst = (size_t)zl * N * a.nxp + kx;
    const cplx* src = a.in + ((size_t)(load ? kx : 0) * a.nzf + zl) * N + j;
    cplx v[8];
    if (PF) cp_async_wait_all();
    if (a.out1 != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const cplx q = PF ? slot[k * NT] : (load ? src[k * T] : cmake(0.0, 0.0));
        const double kk = __ldg(&a.ky[j + k * T]);
        v[k] = cmake(-kk * q.y, kk * q.x);
      }
      fft_regs<N, 1>(v, j, smem, si, twr);
      if (store) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a.out1[dst + (size_t)(j + k * T) * a.nxp] = v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = PF ? slot[k * NT] : (load ? src[k * T] : cmake(0.0, 0.0));
    if (t + (int)gridDim.x < ntiles) issue(t + gridDim.x);
    fft_regs<N, 1>(v, j, smem, si, twr);
    if (store) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a.out0[dst + (size_t)(j + k * T) * a.nxp] = v[k];
    }
  }
}

// ------------------------------------------------------------------------------------------
// yfwd_tile: [zl][y][kx] -> y-FFT -> [kx][zl][ky]
// ------------------------------------------------------------------------------------------
struct YfwdArgs {
  const cplx* in;
  cplx* out;
  int nxh, nxp, nzf;   // nzf: z rows of the slab (line stride of `out` and of the peers)
  int nzc;             // z rows to process (in / out / peers already point at the first one)
  // peer-to-peer: lines of the kx slab [xs[d], xs[d+1]) go straight into rank d's buffer when peer[d] != nullptr
  cplx* peer[8];
  int xs[9];
  int npeer;
};
__device__ __forceinline__ cplx* yfwd_line(const YfwdArgs& a, int kx, int zl, int n) {
  for (int d = 0; d < a.npeer; ++d)
    if (kx >= a.xs[d] && kx < a.xs[d + 1] && a.peer[d] != nullptr) return a.peer[d] + ((size_t)(kx - a.xs[d]) * a.nzf + zl) * n;
  return a.out + ((size_t)kx * a.nzf + zl) * n;
}

template <int N, int NP, int MINB, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_yfwd_tile(YfwdArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  TwRegs<N> twr;
  twr.load(tw, j);
  cplx* slot = smem + (size_t)NP * N + threadIdx.x;
  const int tiles_x = cdiv(a.nxh, NP), ntiles = tiles_x * a.nzc;
  auto issue = [&](int t) {
    if (!PF) return;
    const int kx = (t % tiles_x) * NP + p, zl = t / tiles_x;
    if (kx < a.nxh) {
      const cplx* src = a.in + ((size_t)zl * N + j) * a.nxp + kx;
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async16(slot + k * NT, src + (size_t)k * T * a.nxp);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  int t = blockIdx.x;
  if (t < ntiles) issue(t);
  for (; t < ntiles; t += gridDim.x) {
    const int kx = (t % tiles_x) * NP + p, zl = t / tiles_x;
    cplx v[8];
    if (PF) {
      cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = slot[k * NT];
    } else {
      const cplx* src = a.in + ((size_t)zl * N + j) * a.nxp + (kx < a.nxh ? kx : 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = kx < a.nxh ? src[(size_t)k * T * a.nxp] : cmake(0.0, 0.0);
    }
    if (t + (int)gridDim.x < ntiles) issue(t + gridDim.x);
    fft_regs<N, -1>(v, j, smem, SIdxPencil{p, NP}, twr);
    if (kx < a.nxh) {
      cplx* dst = yfwd_line(a, kx, zl, N);
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[j + k * T] = v[k];
    }
  }
}

// ==========================================================================================
// Bulk-copy (TMA) versions of the tile kernels.  Same tiles, same transforms, same arithmetic; what changes
// is how a tile travels between HBM and the CTA:
//   * the contiguous side (whole lines) arrives as NP cp.async.bulk copies into a line-major staging tile
//     (pitch N+2 elements: the NP lanes of a quarter-warp hit distinct banks), one tile ahead of the transforms;
//   * the strided side (64-byte pieces of NP adjacent lines) is one or two tensor-map boxes {NP, <=256 rows}
//     between HBM and a dense [row][NP] tile -- which is exactly the layout of the exchange buffer, so the
//     results are written into it after the last pass and stored from there;
//   * nothing passes through the load/store pipe except the conflict-free 16-byte shared-memory accesses.
// ==========================================================================================
template <int N, int NP> struct TileGeo {
  static constexpr int T = N / 8, NT = NP * T;
  static constexpr int PITCH = N + 2;               // line-major staging pitch (elements)
  static constexpr int ROWS = N < 256 ? N : 256;    // rows per tensor-map box
  static constexpr int NBOX = N / ROWS;
  static constexpr size_t ALIGN = 128;
};
template <class B, class A> struct Hook2 {
  B b;
  A a;
  __device__ __forceinline__ void before_sync() const { b(); }
  __device__ __forceinline__ void after_sync() const { a(); }
};
template <class B, class A> __device__ __forceinline__ Hook2<B, A> make_hook(B b, A a) { return Hook2<B, A>{b, a}; }
// 128-byte aligned start inside the dynamic shared memory, computed as an element offset so that the compiler
// keeps the shared address space (a pointer round trip through uintptr_t turns every access into a generic LD/ST)
__device__ __forceinline__ int smem_align128_offset(const cplx* s) {
#ifndef SX_EMU
  const unsigned a = (unsigned)__cvta_generic_to_shared(s);
#else
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
#endif
  return (int)(((128u - (unsigned)(a & 127u)) & 127u) / sizeof(cplx));
}

// inverse transform of NP adjacent lines of one plane, optional derivative, results to the strided side.
//   yinv: lines = kx, plane = zl:  src = in + (line*nzf + zl)*N          out box at (kx0, y, zl)  of [zl][y][kx]
//   zinv: lines = ky, plane = kxl: src = in + (kxl*ny + line)*N          out box at (ky0, z, kxl) of [kxl][z][ky]
struct InvTmaArgs {
  const cplx* in;
  double dk;             // wavenumber spacing of the transformed direction: k(e) = (e < N/2 ? e : e - N) * dk, the
                         // product the host tables hold (specter.fpp:772-789); no table loads in the tile loop
  int nlines;            // lines per plane that hold data (nxh / ny)
  int nplanes;
  long line_stride, plane_stride;   // in lines (units of N elements)
  int has1;
};

// destination blocks of the strided side: one tensor map per block (one per rank in the exchange layout),
// block r holds rows [row0[r], row0[r] + rows) of every tile, stored as nbox[r] boxes of brows[r] rows
constexpr int kMaxBlocks = 8;
struct InvTmaMaps {
  TmaMap m[kMaxBlocks];
  int row0[kMaxBlocks], nbox[kMaxBlocks], brows[kMaxBlocks];
  int nblocks;
};

// S = input stages: the staging tile is a ring of S slots filled S tiles ahead (S = 1: one tile ahead, three CTAs
// per SM; S = 2: two CTAs per SM, twice the bytes in flight per CTA and no gap between a slot's consumption and
// the arrival of the next tile -- the load latency was the largest stall of the one-slot kernel, profiles/r1j)
template <int N, int NP, int MINB, int S>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_inv_tma(InvTmaArgs a, const SX_GRID_CONSTANT InvTmaMaps m0,
                                                              const SX_GRID_CONSTANT InvTmaMaps m1, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem_raw);
  typedef TileGeo<N, NP> G;
  constexpr int T = G::T;
  constexpr int STAGE = NP * G::PITCH;
  cplx* exch = smem_raw + smem_align128_offset(smem_raw);          // [row][NP]: exchange buffer and store tile
  cplx* in0 = exch + (size_t)N * NP;             // S x [NP][PITCH]
  unsigned long long* bar0 = reinterpret_cast<unsigned long long*>(in0 + (size_t)S * STAGE);
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  const bool lead = threadIdx.x == 0;
  TwRegs<N> twr;
  twr.load(tw, j);
  const SIdxPencil si{p, NP};
  const int tiles_l = cdiv(a.nlines, NP), ntiles = tiles_l * a.nplanes;
  auto issue = [&](int t, int st) {   // lead thread only
    const int l0 = (t % tiles_l) * NP, pl = t / tiles_l;
    const int valid = a.nlines - l0 < NP ? a.nlines - l0 : NP;
    cplx* in = in0 + (size_t)st * STAGE;
    unsigned long long* bar = bar0 + st;
    mbar_expect(bar, (unsigned)(valid * N * sizeof(cplx)));
    for (int q = 0; q < valid; ++q)
      bulk_load_piece(in + (size_t)q * G::PITCH, a.in + ((size_t)(l0 + q) * a.line_stride + (size_t)pl * a.plane_stride) * N,
                      (unsigned)(N * sizeof(cplx)), bar);
  };
  if (lead) {
    for (int st = 0; st < S; ++st) mbar_init(bar0 + st, 1);
    mbar_init_fence();
  }
  __syncthreads();
  int t = blockIdx.x;
  unsigned phase = 0;   // bit st = parity of stage st
  int cs = 0;           // stage of the current tile
  if (lead)
    for (int st = 0; st < S; ++st)
      if (t + st * (int)gridDim.x < ntiles) issue(t + st * gridDim.x, st);
  for (; t < ntiles; t += gridDim.x) {
    const int l0 = (t % tiles_l) * NP, pl = t / tiles_l;
    const bool load = l0 + p < a.nlines;
    const int tn = t + S * (int)gridDim.x;
    const int st_now = cs;
    mbar_wait(bar0 + cs, (phase >> cs) & 1u);
    phase ^= 1u << cs;
    const cplx* mine = in0 + (size_t)cs * STAGE + (size_t)p * G::PITCH + j;
    if (++cs == S) cs = 0;
    auto store_tile = [&](const InvTmaMaps* m, cplx (&v)[8]) {
      __syncthreads();   // the last gather of the transform is done: the buffer becomes the store tile
#pragma unroll
      for (int k = 0; k < 8; ++k) exch[(size_t)(j + k * T) * NP + p] = v[k];
      fence_proxy_async();
      __syncthreads();
      if (lead) {
        for (int r = 0; r < m->nblocks; ++r)
          for (int b = 0; b < m->nbox[r]; ++b)
            tma_store_3d(&m->m[r], exch + (size_t)(m->row0[r] + b * m->brows[r]) * NP, 2 * l0, b * m->brows[r], pl);
        tma_store_commit();
      }
    };
    cplx v[8];
    if (a.has1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const cplx q = load ? mine[k * T] : cmake(0.0, 0.0);
        const int e = j + k * T;
        const double kk = (double)(e < N / 2 ? e : e - N) * a.dk;
        v[k] = cmake(-kk * q.y, kk * q.x);
      }
      fft_regs<N, 1>(v, j, exch, si, twr, make_hook([&] { if (lead) tma_store_wait_read(); }, [] {}));
      store_tile(&m1, v);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = load ? mine[k * T] : cmake(0.0, 0.0);
    // at the first barrier of this transform every thread has consumed the staging tile: refill it
    fft_regs<N, 1>(v, j, exch, si, twr,
                   make_hook([&] { if (lead) tma_store_wait_read(); }, [&] { if (lead && tn < ntiles) issue(tn, st_now); }));
    store_tile(&m0, v);
  }
  if (lead) tma_store_wait_all();
}

// yfwd: box loads of the strided side [zl][y][kx] -> y-FFT -> whole lines [kx][zl][ky] stored from registers
template <int N, int NP, int MINB>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_yfwd_tma(YfwdArgs a, const SX_GRID_CONSTANT TmaMap min,
                                                               const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem_raw);
  typedef TileGeo<N, NP> G;
  constexpr int T = G::T;
  cplx* exch = smem_raw + smem_align128_offset(smem_raw);
  cplx* in = exch + (size_t)N * NP;              // [row][NP]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(in + (size_t)N * NP);
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  const bool lead = threadIdx.x == 0;
  TwRegs<N> twr;
  twr.load(tw, j);
  const int tiles_x = cdiv(a.nxh, NP), ntiles = tiles_x * a.nzc;
  auto issue = [&](int t) {
    const int kx0 = (t % tiles_x) * NP, zl = t / tiles_x;
    mbar_expect(bar, (unsigned)((size_t)N * NP * sizeof(cplx)));
#pragma unroll
    for (int b = 0; b < G::NBOX; ++b) tma_load_3d(in + (size_t)b * G::ROWS * NP, &min, 2 * kx0, b * G::ROWS, zl, bar);
  };
  if (lead) {
    mbar_init(bar, 1);
    mbar_init_fence();
  }
  __syncthreads();
  int t = blockIdx.x;
  unsigned phase = 0;
  if (lead && t < ntiles) issue(t);
  for (; t < ntiles; t += gridDim.x) {
    const int kx = (t % tiles_x) * NP + p, zl = t / tiles_x;
    const int tn = t + gridDim.x;
    mbar_wait(bar, phase);
    phase ^= 1;
    cplx v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = kx < a.nxh ? in[(size_t)(j + k * T) * NP + p] : cmake(0.0, 0.0);
    fft_regs<N, -1>(v, j, exch, SIdxPencil{p, NP}, twr, make_hook([] {}, [&] { if (lead && tn < ntiles) issue(tn); }));
    if (kx < a.nxh) {
      cplx* dst = yfwd_line(a, kx, zl, N);
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[j + k * T] = v[k];
    }
  }
}

template <int N, int NP> static size_t inv_tma_smem(int stages) {
  return ((size_t)N * NP + (size_t)stages * NP * TileGeo<N, NP>::PITCH) * sizeof(cplx) + 8 * stages + TileGeo<N, NP>::ALIGN;
}
// returns -1 when the bulk-copy path does not apply (caller falls back to the register-path kernels)
template <int N> static int run_inv_tma(Plan& p, Fused& f, int stage, const InvTmaArgs& a, const InvTmaMaps& m0, const InvTmaMaps& m1) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  const cplx* tw = stage == ST_ZINV ? p.tw_z : p.tw_y;
  const int ntiles = cdiv(a.nlines, NP) * a.nplanes;
  int grid;
  // two input stages at two CTAs per SM (env SX_INV_STAGES: bit 0 zinv, bit 1 yinv), where two such CTAs fit an SM
  if constexpr (N >= 256 && MINB >= 2) {
    if (inv_tma_smem<N, NP>(2) * 2 <= 220 * 1024 && (p.knob_inv_stages & (stage == ST_ZINV ? 1 : 2))) {
      auto kfn2 = k_inv_tma<N, NP, 2, 2>;
      const size_t smem2 = inv_tma_smem<N, NP>(2);
      if (persistent_grid(p, kfn2, NP * (N / 8), smem2, ntiles, &grid)) return 1;
      SX_FUSED_LAUNCH(p, stage, kfn2, dim3(grid), NP * (N / 8), smem2, a, m0, m1, tw);
      return 0;
    }
  }
  auto kfn = k_inv_tma<N, NP, MINB, 1>;
  const size_t smem = inv_tma_smem<N, NP>(1);
  if (persistent_grid(p, kfn, NP * (N / 8), smem, ntiles, &grid)) return 1;
  SX_FUSED_LAUNCH(p, stage, kfn, dim3(grid), NP * (N / 8), smem, a, m0, m1, tw);
  return 0;
}
// one destination block: rows [row0, row0 + rows) of the tiles -> tensor (n0 lines, rows, n2 planes) at `base`
static int add_block(InvTmaMaps& m, const void* base, size_t n0, int row0, int rows, size_t n2, size_t pitch1, size_t pitch2, int np) {
  const int r = m.nblocks++;
  m.row0[r] = row0;
  m.nbox[r] = (rows + 255) / 256;
  m.brows[r] = m.nbox[r] > 1 ? 256 : rows;
  return tma_encode(&m.m[r], base, n0, rows, n2, pitch1, pitch2, np, m.brows[r]);
}
template <int N> static int run_zinv_tma(Plan& p, Fused& f, const cplx* in, cplx* out0, cplx* out1) {
  constexpr int NP = TileNP<N>::value;
  if (N < p.knob_tma_min || p.nprocs > kMaxBlocks || p.ny % NP != 0 || !(p.knob_tma & 1)) return -1;
  // exchange layout [rank][kxl][zl_r][ky], physical rows only; a block's tile rows must start 128-byte aligned
  std::vector<int> z0(p.nprocs), zc(p.nprocs);
  for (int r = 0; r < p.nprocs; ++r) {
    zc[r] = (int)(f.z_count[r] / ((size_t)p.nxl * p.ny));
    z0[r] = r == 0 ? 0 : z0[r - 1] + zc[r - 1];
    if (zc[r] > 0 && ((size_t)z0[r] * NP * sizeof(cplx)) % 128 != 0) return -1;
  }
  InvTmaMaps m0, m1;
  m0.nblocks = m1.nblocks = 0;
  // peer-to-peer: a block may be stored straight into the destination rank's receive buffer (same geometry there)
  const int s0 = slot_of(f.W, out0), s1 = out1 ? slot_of(f.W, out1) : s0;
  f.zinv_direct = f.p2p && s0 >= 0 && s1 >= 0;
  for (int r = 0; r < p.nprocs; ++r) {
    if (zc[r] == 0) continue;
    const bool direct = f.zinv_direct && direct_to(p, f, r);
    cplx* b0 = direct ? peer_r_dst(p, f, s0, r) : out0 + f.z_displ[r];
    cplx* b1 = direct ? peer_r_dst(p, f, s1, r) : (out1 ? out1 : out0) + f.z_displ[r];
    if (add_block(m0, b0, p.ny, z0[r], zc[r], p.nxl, p.ny, (size_t)zc[r] * p.ny, NP)) return 1;
    if (add_block(m1, b1, p.ny, z0[r], zc[r], p.nxl, p.ny, (size_t)zc[r] * p.ny, NP)) return 1;
  }
  InvTmaArgs a{in, p.Dkz, p.ny, p.nxl, 1, p.ny, out1 != nullptr};
  return run_inv_tma<N>(p, f, ST_ZINV, a, m0, m1);
}
template <int N> static int run_yinv_tma(Plan& p, Fused& f, const cplx* in, cplx* out0, cplx* out1) {
  constexpr int NP = TileNP<N>::value;
  if (N < p.knob_tma_min || f.nxp % NP != 0 || !(p.knob_tma & 2)) return -1;
  if (f.nzf == 0) return 0;
  InvTmaMaps m0, m1;
  m0.nblocks = m1.nblocks = 0;
  const size_t zo = f.vz0() * p.ny * f.nxp;   // z window of the [zl][y][kx] outputs
  if (f.zc() == 0) return 0;
  if (add_block(m0, out0 + zo, f.nxp, 0, N, f.zc(), f.nxp, (size_t)p.ny * f.nxp, NP)) return 1;
  if (add_block(m1, (out1 ? out1 : out0) + zo, f.nxp, 0, N, f.zc(), f.nxp, (size_t)p.ny * f.nxp, NP)) return 1;
  InvTmaArgs a{in + (size_t)f.z0() * N, p.Dky, p.nxh, f.zc(), f.nzf, 1, out1 != nullptr};
  return run_inv_tma<N>(p, f, ST_YINV, a, m0, m1);
}
static int yfwd_peers(Plan& p, Fused& f, const cplx* out, YfwdArgs& a) {
  const int slot = slot_of(f.U, out);
  f.yfwd_direct = f.p2p && slot >= 0 && p.nprocs <= 8 && f.direct >= 1;
  if (!f.yfwd_direct) return 0;
  a.npeer = p.nprocs;
  for (int d = 0; d < p.nprocs; ++d) {
    int xs, xc;
    range0(p.nxh, p.nprocs, d, &xs, &xc);
    a.xs[d] = xs;
    a.xs[d + 1] = xs + xc;
    a.peer[d] = direct_to(p, f, d) ? peer_uz_dst(p, f, slot, d) + (size_t)f.z0() * p.ny : nullptr;
  }
  return 0;
}
template <int N> static int run_yfwd_tma(Plan& p, Fused& f, const cplx* in, cplx* out) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  typedef TileGeo<N, NP> G;
  // tensor-map box loads of the strided side: slower than the cp.async slots at N = 512 (r1h), faster from N = 1024
  // (0.773 -> 0.613 ms per launch at ny = 2048, profiles/r2j)
  if (N < p.knob_tma_min || f.nxp % NP != 0 || !((p.knob_tma & 4) || N >= 1024)) return -1;
  if (f.nzf == 0) return 0;
  TmaMap min;
  if (f.zc() == 0) return 0;
  if (tma_encode(&min, in + f.vz0() * p.ny * f.nxp, f.nxp, p.ny, f.zc(), f.nxp, (size_t)p.ny * f.nxp, NP, G::ROWS)) return 1;
  YfwdArgs a{in + f.vz0() * p.ny * f.nxp, out + (size_t)f.z0() * N, p.nxh, f.nxp, f.nzf, f.zc(),
             {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, {0}, 0};
  if (yfwd_peers(p, f, out, a)) return 1;
  auto kfn = k_yfwd_tma<N, NP, MINB>;
  const size_t smem = (size_t)2 * N * NP * sizeof(cplx) + 8 + G::ALIGN;
  int grid;
  if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.nxh, NP) * f.zc(), &grid)) return 1;
  const cplx* tw = p.tw_y;
  SX_FUSED_LAUNCH(p, ST_YFWD, kfn, dim3(grid), NP * (N / 8), smem, a, min, tw);
  return 0;
}

template <int N> static int run_zinv(Plan& p, Fused& f, const cplx* in, cplx* out0, cplx* out1) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  if (const int rc = run_zinv_tma<N>(p, f, in, out0, out1); rc >= 0) return rc;
  f.zinv_direct = false;
  ZinvArgs a{in, out0, out1, p.d_kz, f.d_zmap, p.ny, p.nxl, f.nph};
  const cplx* tw = p.tw_z;
  const size_t smem = (size_t)2 * NP * N * sizeof(cplx) + (size_t)N * sizeof(ZMap);
  int grid;
  if (p.knob_pf & 1) {
    auto kfn = k_zinv_tile<N, NP, MINB, true>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_zinv_tile<N, NP, MINB, false>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
template <int N> static int run_yinv(Plan& p, Fused& f, const cplx* in, cplx* out0, cplx* out1) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  if (const int rc = run_yinv_tma<N>(p, f, in, out0, out1); rc >= 0) return rc;
  if (f.zc() == 0) return 0;
  const size_t zo = f.vz0() * N * f.nxp;
  YinvArgs a{in + (size_t)f.z0() * N, out0 + zo, out1 ? out1 + zo : nullptr, p.d_ky, p.nxh, f.nxp, f.nzf, f.zc()};
  const cplx* tw = p.tw_y;
  int grid;
  if (p.knob_pf & 2) {
    auto kfn = k_yinv_tile<N, NP, MINB, true>;
    const size_t smem = (size_t)2 * NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(f.nxp, NP) * f.zc(), &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_yinv_tile<N, NP, MINB, false>;
    const size_t smem = (size_t)NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(f.nxp, NP) * f.zc(), &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
template <int N> static int run_yfwd(Plan& p, Fused& f, const cplx* in, cplx* out) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  if (const int rc = run_yfwd_tma<N>(p, f, in, out); rc >= 0) return rc;
  if (f.zc() == 0) return 0;
  YfwdArgs a{in + f.vz0() * p.ny * f.nxp, out + (size_t)f.z0() * N, p.nxh, f.nxp, f.nzf, f.zc(),
             {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, {0}, 0};
  if (yfwd_peers(p, f, out, a)) return 1;
  const cplx* tw = p.tw_y;
  int grid;
  if (p.knob_pf & 4) {
    auto kfn = k_yfwd_tile<N, NP, MINB, true>;
    const size_t smem = (size_t)2 * NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.nxh, NP) * f.zc(), &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YFWD, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_yfwd_tile<N, NP, MINB, false>;
    const size_t smem = (size_t)NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.nxh, NP) * f.zc(), &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YFWD, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
int fused_zinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1) {
#define C_(N) run_zinv<N>(p, f, in, o0, o1)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}
int fused_yinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1) {
#define C_(N) run_yinv<N>(p, f, in, o0, o1)
  SX_SIZE_SWITCH(p.ny, C_);
#undef C_
}
int fused_yfwd(Plan& p, Fused& f, const cplx* in, cplx* out) {
#define C_(N) run_yfwd<N>(p, f, in, out)
  SX_SIZE_SWITCH(p.ny, C_);
#undef C_
}

}  // namespace sx
