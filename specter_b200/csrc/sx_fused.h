// Shared declarations of the fused-substep translation units (sx_fused*.cu).
#pragma once
#include "sx_fft.cuh"
#include "sx_plan.h"

namespace sx {


// where physical row z lives in the exchange layout [rank][kxl][zl][ky]
struct alignas(16) ZMap {
  long long base;  // complex elements before this rank's block
  int nzl;         // rows held by the owning rank
  int zl;          // row index inside the owning rank
};

struct Fused {
  int nph = 0;     // physical rows nz - Cz
  int nzf = 0;     // physical rows owned by this rank in real space (balanced partition)
  int zf0 = 0;     // first owned physical row (0-based)
  int nxp = 0;     // padded kx extent of the [z][y][kx] arrays (multiple of 8)
  size_t wsize = 0;  // complex elements of one exchange-layout slot: nxl * nph * ny
  size_t vsize = 0;  // complex elements of one [z][y][kx] slot: nzf * ny * nxp
  ZMap* d_zmap = nullptr;
  // work fields, grown on demand by fused_reserve (HD 6/9/3, BOUSS 8/12/4, MHD 12/12/6)
  std::vector<cplx*> W;   // z-stage side, exchange layout (fields and their z derivatives)
  std::vector<cplx*> R;   // y-stage side [kx][zl][ky] (aliases W on one GPU)
  std::vector<cplx*> V;   // [zl][y][kx]: fields, dy, dz
  std::vector<cplx*> X;   // [zl][y][kx]: nonlinear terms after the x pass
  std::vector<cplx*> U;   // y-stage side of the way back [kx][zl][ky]
  std::vector<cplx*> Uz;  // z-stage side of the way back (aliases U on one GPU)
  // all-to-all-v block tables in complex elements: z side [rank][kxl][zl_r][ky], xy side [kx][zl][ky]
  std::vector<size_t> z_displ, z_count, x_displ, x_count;
  // peer-to-peer exchange (sx_plan_p2p_export / _import): R and Uz are carved from one IPC-exported arena per rank
  cplx* arena = nullptr;
  int arena_nw = 0, arena_nx = 0;
  std::vector<cplx*> peer_arena;   // IPC-mapped arenas of the other ranks (own entry = arena)
  bool p2p = false;
  // which blocks the producing kernels store straight into the destination rank's buffer instead of the send
  // buffer: 0 none (all blocks travel by copy engine), 1 the local block, 2 every block (stores over NVLink)
  int direct = 1;
  // with direct == 1: the blocks of the next `direct_peers` ranks (ring order) are ALSO stored straight over NVLink by the
  // producing kernels while the copy engines carry the rest -- the two paths add up on the links (env SX_P2P_DIRECT_PEERS)
  int direct_peers = 0;
  bool zinv_direct = false, yfwd_direct = false;   // set by the launchers when the kernel variant in use does so
  // z window of the xy stage (y-inverse, x pass, y-forward launchers): rows [zw0, zw0 + zwc) of the local slab;
  // zwc < 0 = the whole slab.  The chunked multi-rank pipeline (sx_fused.cu) moves it from chunk to chunk.
  int zw0 = 0, zwc = -1;
  bool chunked_now = false;   // the current substep ran the chunked pipeline (its exchanges are already waited for)
  int z0() const { return zwc < 0 ? 0 : zw0; }
  int zc() const { return zwc < 0 ? nzf : zwc; }
  // The [zl][y][kx] arrays V / X only live between the y-inverse and the y-forward kernels of one z window: the
  // chunked pipeline allocates them for the rows of ONE chunk (vwin) and every chunk reuses them from row 0, which is
  // what lets 2048 x 2048 x 1024 on 8 GPUs fit 180 GB (V[9] + X[3] are 12 of the 40 work fields otherwise).
  int vrows = 0;        // rows the V / X pools are allocated for
  bool vwin = false;    // this substep addresses V / X relative to the current z window
  size_t vz0() const { return vwin ? 0 : (size_t)z0(); }
};

// does the producing kernel store the block of rank r itself (instead of leaving it to the copy engines)?
inline bool direct_to(const Plan& p, const Fused& f, int r) {
  if (f.direct >= 2) return true;
  if (f.direct < 1) return false;
  return (r - p.myrank + p.nprocs) % p.nprocs <= f.direct_peers;
}

inline void range0(int n, int nprocs, int r, int* sta, int* cnt) {  // `range` on [0,n)
  const int w = n / nprocs, m = n % nprocs;
  *sta = r * w + (r < m ? r : m);
  *cnt = w + (m > r ? 1 : 0);
}
// The real-space z partition of the fused substep is internal (the ABI only exposes the reference's ksta:kend for
// the per-operator entries): rows are dealt in pairs so that every rank's first row is even and its rows start
// 128-byte aligned in the [row][4 lines] tiles the tensor-map stores read.
inline void zrange(int nph, int nprocs, int r, int* sta, int* cnt) {
  int us, uc;
  range0((nph + 1) / 2, nprocs, r, &us, &uc);
  *sta = 2 * us < nph ? 2 * us : nph;
  const int end = 2 * (us + uc) < nph ? 2 * (us + uc) : nph;
  *cnt = end - *sta;
}

// element counts (rounded to 256 B) of one receive buffer of rank r: xy side [kx][zl_r][ky], z side [rank][kxl_r][zl][ky]
inline size_t arena_rs(const Plan& p, int r) {
  int s, c;
  zrange(p.nz - p.Cz, p.nprocs, r, &s, &c);
  return ((size_t)p.nxh * c * p.ny + 15) / 16 * 16;
}
inline size_t arena_ws(const Plan& p, int r) {
  int s, c;
  range0(p.nxh, p.nprocs, r, &s, &c);
  return ((size_t)c * (p.nz - p.Cz) * p.ny + 15) / 16 * 16;
}

// where this rank's block starts inside rank r's receive buffer `slot` of the way to real space ([kx][zl_r][ky]:
// my kx slab) and of the way back ([rank][kxl_r][zl][ky]: behind the z slabs of the lower ranks)
inline cplx* peer_r_dst(const Plan& p, const Fused& f, int slot, int r) {
  int xs, xc, zs, zc;
  range0(p.nxh, p.nprocs, p.myrank, &xs, &xc);
  zrange(f.nph, p.nprocs, r, &zs, &zc);
  return f.peer_arena[r] + slot * arena_rs(p, r) + (size_t)xs * zc * p.ny;
}
inline cplx* peer_uz_dst(const Plan& p, const Fused& f, int slot, int r) {
  int xs, xc;
  range0(p.nxh, p.nprocs, r, &xs, &xc);
  size_t before = 0;
  for (int q = 0; q < p.myrank; ++q) {
    int zs, zc;
    zrange(f.nph, p.nprocs, q, &zs, &zc);
    before += (size_t)xc * zc * p.ny;
  }
  return f.peer_arena[r] + f.arena_nw * arena_rs(p, r) + slot * arena_ws(p, r) + before;
}
inline int slot_of(const std::vector<cplx*>& pool, const cplx* ptr) {
  for (size_t i = 0; i < pool.size(); ++i)
    if (pool[i] == ptr) return (int)i;
  return -1;
}

// lines per CTA of the tile kernels: 256 threads up to N = 512 (64 B pieces on the strided side --
// measured FASTER on B200 than 128 B pieces with twice the CTA footprint), 4 lines beyond
template <int N> struct TileNP {
  static constexpr int value = N <= 64 ? 32 : (N == 128 ? 16 : (N == 256 ? 8 : (N <= 1024 ? 4 : 2)));
};
template <int N> struct TileMinB {  // CTAs per SM the register budget is capped for
  static constexpr int value = N <= 512 ? 3 : (N == 1024 ? 1 : 1);
};

// Every tile kernel is persistent (grid = a multiple of the SM count, tiles strided by gridDim) and
// software-pipelined: while tile t is transformed, the 8 values per thread of tile t+gridDim are in
// flight as 16-byte cp.async copies into thread-private shared-memory slots (slot k of thread tid at
// stage[k*NT + tid]: conflict free, and no barrier is needed because only the issuing thread reads
// them).  The registers are the second pipeline stage.

constexpr int kMaxDF = 10;   // largest supported number of FC-Gram matching points

struct RkTerm {   // see ZfwdArgs
  const cplx* couple = nullptr;
  double ccoef = 0.0, cL = 0.0, sNL = -1.0;
  int lap = 1;
};

#define SX_FUSED_LAUNCH(p, stage, kfn, grid, threads, smem, ...)                                         \
  do {                                                                                                   \
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
    if (stage_mark((p), (stage))) return 1;                                                              \
    cudaStream_t st_ = (p).stream;                                                                       \
    SX_LAUNCH(kfn, grid, dim3(threads), (smem), st_, __VA_ARGS__);                                       \
    (p).launches++;                                                                                      \
    SX_KERNEL_CHECK();                                                                                   \
  } while (0)

// persistent grid: CTAs per SM from the occupancy calculator, times the SM count
template <class K> static int persistent_grid(Plan& p, K kfn, int threads, size_t smem, int ntiles, int* grid) {
  SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
#ifndef SX_EMU
  SX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem));
  if (per_sm < 1) per_sm = 1;
#endif
  const int g = per_sm * p.num_sms;
  *grid = ntiles < g ? ntiles : g;
  return 0;
}

#define SX_SIZE_SWITCH(n, CALL)                 \
  switch (n) {                                  \
    case 16: return CALL(16);                   \
    case 32: return CALL(32);                   \
    case 64: return CALL(64);                   \
    case 128: return CALL(128);                 \
    case 256: return CALL(256);                 \
    case 512: return CALL(512);                 \
    case 1024: return CALL(1024);               \
    case 2048: return CALL(2048);               \
  }                                             \
  SX_REQUIRE(false, "unsupported transform length (power of two in [16,2048])")


// stage entry points (one translation unit per kernel family)
int fused_zinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1);
int fused_yinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1);
int fused_yfwd(Plan& p, Fused& f, const cplx* in, cplx* out);
int fused_xpass(Plan& p, Fused& f, int nc, const double* kxg);
int fused_xcross(Plan& p, Fused& f, int npairs, const int* Pi, const int* Qi, const double* sgn, int xo);
int fused_xadvect(Plan& p, Fused& f, int ui, int qi, int xo, const double* kxg);
int fused_zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0, const cplx* frc,
                  const RkTerm& rk, double dt, double rmp);
int fused_project(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o, const double* zs, const double* ze);
// a_imposebc_and_project as one pencil kernel (conducting walls); -1: not applicable, compose the operators
int fused_aproject(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph);

}  // namespace sx
