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
};

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
int fused_zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0, const cplx* frc,
                  const RkTerm& rk, double dt, double rmp);
int fused_project(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o, const double* zs, const double* ze);

}  // namespace sx
