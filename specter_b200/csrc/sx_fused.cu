// The fused HD Runge-Kutta substep (include/hd/hd_rkstep2.f90:3-36 of the reference) as six
// memory-bound passes.  Every pass is one FFT axis with the neighbouring elementwise work and the
// slab transposition folded into its load / store pattern:
//
//   zinv_tile   v^(kz,ky,kx)      -> z-IFFT of v and of i kz v, physical rows only, written ky-fastest
//                                    in the exchange layout [rank][kx][z][ky]           (fftp.fpp:1060-1094)
//   yinv_tile   [kx][z][ky]       -> y-IFFT of (v, i ky v, dz v), written kx-fastest [z][y][kx]
//   xpass       [z][y][kx] x 9    -> 12 c2r lines (i kx applied on load), (u.grad)u / N^2, 3 r2c lines
//                                                                           (pseudospec_hd.f90:245-316)
//   yfwd_tile   [z][y][kx] x 3    -> y-FFT, written ky-fastest [kx][z][ky]
//   zfwd_rk     [kx][z][ky]       -> FC-Gram continuation + z-FFT (fftp.fpp:757-780), fc_filter
//                                    (pseudospec_hd.f90:1099-1112), Laplacian + RK update (hd_rkstep2.f90:14-32)
//   project     per (ky,kx) pencil: no-slip walls + Poisson/Laplace projection, nine z transforms in
//                                    shared memory (vboundary.f90:116-148, boundary_mod.fpp:197-402)
//
// Derivative sharing: d/dx and d/dy commute with the z-IFFT and the transposition, so only v and
// dz v (6 fields, not 12) cross the transposition; rows above the physical region never do (the
// products only use k <= pkend, pseudospec_hd.f90:255, and the forward continuation overwrites them).
//
// Thread mapping of the "tile" kernels: NP lines per CTA, lane-fastest over the NP lines
// (p = tid % NP, j = tid / NP, T = N/8 threads per line).  One side of the kernel then moves
// NP*16 B = 128 B full lines per quarter-warp and the other side 64 B pieces of NP different lines,
// which is the transposition.
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
      fft_regs<N, 1>(v, j, smem, si, tw);
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
    fft_regs<N, 1>(v, j, smem, si, tw);
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
  int nxh, nxp, nzf;
};

template <int N, int NP, int MINB, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_yinv_tile(YinvArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  cplx* slot = smem + (size_t)NP * N + threadIdx.x;
  const SIdxPencil si{p, NP};
  const int tiles_x = cdiv(a.nxp, NP), ntiles = tiles_x * a.nzf;
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
      fft_regs<N, 1>(v, j, smem, si, tw);
      if (store) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a.out1[dst + (size_t)(j + k * T) * a.nxp] = v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = PF ? slot[k * NT] : (load ? src[k * T] : cmake(0.0, 0.0));
    if (t + (int)gridDim.x < ntiles) issue(t + gridDim.x);
    fft_regs<N, 1>(v, j, smem, si, tw);
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
  int nxh, nxp, nzf;
};

template <int N, int NP, int MINB, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_yfwd_tile(YfwdArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  cplx* slot = smem + (size_t)NP * N + threadIdx.x;
  const int tiles_x = cdiv(a.nxh, NP), ntiles = tiles_x * a.nzf;
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
    fft_regs<N, -1>(v, j, smem, SIdxPencil{p, NP}, tw);
    if (kx < a.nxh) {
      cplx* dst = a.out + ((size_t)kx * a.nzf + zl) * N;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[j + k * T] = v[k];
    }
  }
}

// ------------------------------------------------------------------------------------------
// xpass (gradre): one CTA = LP pairs of adjacent y lines of one local z row.  Two real lines ride
// one complex FFT of length nx:  Z(k) = A(k) + i B(k) on the Hermitian-completed spectra.  FFTW's
// c2r ignores Im of the kx = 0 and kx = nx/2 entries; so do we (after the i kx factor, as the
// reference applies derivk before the transform).
// ------------------------------------------------------------------------------------------
struct XpassArgs {
  const cplx* V[12];  // q(NC), dy q(NC), dz q(NC) with q = (vx, vy, vz[, theta])   [zl][y][kx]
  cplx* X[4];
  const double* kx;  // GLOBAL kx(1:nx/2+1)
  int ny, nxp, nzf;
  double tmp;        // 1/(nx ny nz)^2
};

// One CTA = LP pairs of adjacent y lines; persistent over (z row, y group).  The 12 inverse
// transforms of a group are a software pipeline: while transform m runs, the two spectral rows of
// transform m+1 are in flight as cp.async copies into thread-private slots.  The three velocity
// lines are parked in thread-private shared memory, so the register file only holds one transform
// and one accumulator.
template <int N, int LP, bool PF, int MINB, int NC>
__global__ void __launch_bounds__(LP*(N / 8), MINB) k_xpass_gradre(XpassArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = LP * T, XS = sidx_elem_stride<N>();
  const int lp = threadIdx.x / T, t = threadIdx.x % T;
  const SIdxElem si{lp * XS};
  cplx* park = smem + (size_t)LP * XS + threadIdx.x;            // park[(c*8+k)*NT]
  cplx* slot = smem + (size_t)LP * XS + (size_t)24 * NT + threadIdx.x;  // slot[(2k+h)*NT]
  const int groups_y = cdiv(a.ny, 2 * LP), ngroups = groups_y * a.nzf;
  constexpr int NM = 3 + 3 * NC;  // inverse transforms per group: u(3), then d_x, d_y, d_z of each component
  auto field_of = [&](int m) -> const cplx* {
    if (m < 3) return a.V[m];
    const int c = (m - 3) / 3, d = (m - 3) % 3;
    return a.V[d * NC + c];
  };
  auto issue = [&](int g, int m) {
    if (!PF) return;
    const int y0 = ((g % groups_y) * LP + lp) * 2, zl = g / groups_y;
    if (y0 < a.ny) {
      const cplx* rowA = field_of(m) + ((size_t)zl * a.ny + y0) * a.nxp;
      const cplx* rowB = rowA + a.nxp;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = t + k * T;
        const int kx = e <= N / 2 ? e : N - e;
        cp_async16(slot + (2 * k) * NT, rowA + kx);
        cp_async16(slot + (2 * k + 1) * NT, rowB + kx);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  int g = blockIdx.x;
  if (g < ngroups) issue(g, 0);
  for (; g < ngroups; g += gridDim.x) {
    const int y0 = ((g % groups_y) * LP + lp) * 2, zl = g / groups_y;
    const bool active = y0 < a.ny;
    const size_t rowA = ((size_t)zl * a.ny + (active ? y0 : 0)) * a.nxp, rowB = rowA + a.nxp;
    cplx acc[8];
#pragma unroll 1
    for (int m = 0; m < NM; ++m) {
      const bool deriv = m >= 3 && (m - 3) % 3 == 0;
      cplx v[8];
      if (PF) cp_async_wait_all();
      const cplx* fA = field_of(m) + rowA;
      const cplx* fB = fA + a.nxp;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = t + k * T;
        const int kx = e <= N / 2 ? e : N - e;
        cplx A, B;
        if (PF) {
          A = slot[(2 * k) * NT];
          B = slot[(2 * k + 1) * NT];
        } else {
          A = active ? fA[kx] : cmake(0.0, 0.0);
          B = active ? fB[kx] : cmake(0.0, 0.0);
        }
        if (deriv) {
          const double kk = __ldg(&a.kx[kx]);
          A = cmake(-kk * A.y, kk * A.x);
          B = cmake(-kk * B.y, kk * B.x);
        }
        if (kx == 0 || kx == N / 2) { A.y = 0.0; B.y = 0.0; }
        if (e > N / 2) { A.y = -A.y; B.y = -B.y; }
        v[k] = cmake(A.x - B.y, A.y + B.x);
      }
      if (m < NM - 1) issue(g, m + 1);
      else if (g + (int)gridDim.x < ngroups) issue(g + gridDim.x, 0);
      fft_regs<N, 1>(v, t, smem, si, tw);
      if (m < 3) {
#pragma unroll
        for (int k = 0; k < 8; ++k) park[(m * 8 + k) * NT] = v[k];
      } else {
        const int c = (m - 3) / 3, d = (m - 3) % 3;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const cplx u = park[(d * 8 + k) * NT];
          if (d == 0) acc[k] = cmake(u.x * v[k].x, u.y * v[k].y);
          else acc[k] = cmake(acc[k].x + u.x * v[k].x, acc[k].y + u.y * v[k].y);
        }
        if (d == 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = cmake(acc[k].x * a.tmp, acc[k].y * a.tmp);
          // forward transform of the packed pair and split into the two half spectra
          fft_regs<N, -1>(acc, t, smem, si, tw);
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 8; ++k) smem[si(t + k * T)] = acc[k];
          __syncthreads();
          cplx* out = a.X[c];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int kk = t + k * T;
            if (kk <= N / 2 && active) {
              const cplx Zk = acc[k];
              const cplx Zn = smem[si((N - kk) & (N - 1))];
              out[rowA + kk] = cmake(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
              out[rowB + kk] = cmake(0.5 * (Zk.y + Zn.y), -0.5 * (Zk.x - Zn.x));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// xpass (cross products): X = sum over pairs s * (P x Q) / N^2 on the physical rows.  Used for the MHD
// nonlinear terms: omega x v - J x B = -(v x omega) + (B x J) (prodre pseudospec_hd.f90:357-399 and vector
// pseudospec_mhd.f90:87-102 as called at mhd_rkstep2.f90:29-37) and the electromotive force v x B (:45).
// The three lines of P are parked in thread-private shared memory, the lines of Q stream through the
// registers one transform at a time and feed three accumulators.
// ------------------------------------------------------------------------------------------
struct XcrossArgs {
  const cplx* P[2][3];
  const cplx* Q[2][3];
  double sgn[2];
  int npairs;
  cplx* X[3];
  int ny, nxp, nzf;
  double tmp;  // 1/(nx ny nz)^2
};

template <int N, int LP, int MINB>
__global__ void __launch_bounds__(LP*(N / 8), MINB) k_xpass_cross(XcrossArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = LP * T, XS = sidx_elem_stride<N>();
  const int lp = threadIdx.x / T, t = threadIdx.x % T;
  const SIdxElem si{lp * XS};
  cplx* park = smem + (size_t)LP * XS + threadIdx.x;                     // park[(c*8+k)*NT]
  cplx* slot = smem + (size_t)LP * XS + (size_t)24 * NT + threadIdx.x;   // slot[(2k+h)*NT]
  const int groups_y = cdiv(a.ny, 2 * LP), ngroups = groups_y * a.nzf;
  const int nm = 6 * a.npairs;
  auto field_of = [&](int m) -> const cplx* {
    const int pr = m / 6, q = m % 6;
    return q < 3 ? a.P[pr][q] : a.Q[pr][q - 3];
  };
  auto issue = [&](int g, int m) {
    const int y0 = ((g % groups_y) * LP + lp) * 2, zl = g / groups_y;
    if (y0 < a.ny) {
      const cplx* rowA = field_of(m) + ((size_t)zl * a.ny + y0) * a.nxp;
      const cplx* rowB = rowA + a.nxp;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = t + k * T;
        const int kx = e <= N / 2 ? e : N - e;
        cp_async16(slot + (2 * k) * NT, rowA + kx);
        cp_async16(slot + (2 * k + 1) * NT, rowB + kx);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  int g = blockIdx.x;
  if (g < ngroups) issue(g, 0);
  for (; g < ngroups; g += gridDim.x) {
    const int y0 = ((g % groups_y) * LP + lp) * 2, zl = g / groups_y;
    const bool active = y0 < a.ny;
    const size_t rowA = ((size_t)zl * a.ny + (active ? y0 : 0)) * a.nxp, rowB = rowA + a.nxp;
    cplx ax[8], ay[8], az[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ax[k] = ay[k] = az[k] = cmake(0.0, 0.0);
#pragma unroll 1
    for (int m = 0; m < nm; ++m) {
      const int q = m % 6;
      const double sg = a.sgn[m / 6];
      cplx v[8];
      cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = t + k * T;
        const int kx = e <= N / 2 ? e : N - e;
        cplx A = slot[(2 * k) * NT], B = slot[(2 * k + 1) * NT];
        if (kx == 0 || kx == N / 2) { A.y = 0.0; B.y = 0.0; }
        if (e > N / 2) { A.y = -A.y; B.y = -B.y; }
        v[k] = cmake(A.x - B.y, A.y + B.x);
      }
      if (m < nm - 1) issue(g, m + 1);
      else if (g + (int)gridDim.x < ngroups) issue(g + gridDim.x, 0);
      fft_regs<N, 1>(v, t, smem, si, tw);
      if (q < 3) {
#pragma unroll
        for (int k = 0; k < 8; ++k) park[(q * 8 + k) * NT] = v[k];
      } else if (q == 3) {   // Q_x: y += P_z Q_x, z -= P_y Q_x
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const cplx py = park[(8 + k) * NT], pz = park[(16 + k) * NT];
          ay[k] = cmake(fma(sg * pz.x, v[k].x, ay[k].x), fma(sg * pz.y, v[k].y, ay[k].y));
          az[k] = cmake(fma(-sg * py.x, v[k].x, az[k].x), fma(-sg * py.y, v[k].y, az[k].y));
        }
      } else if (q == 4) {   // Q_y: x -= P_z Q_y, z += P_x Q_y
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const cplx px = park[k * NT], pz = park[(16 + k) * NT];
          ax[k] = cmake(fma(-sg * pz.x, v[k].x, ax[k].x), fma(-sg * pz.y, v[k].y, ax[k].y));
          az[k] = cmake(fma(sg * px.x, v[k].x, az[k].x), fma(sg * px.y, v[k].y, az[k].y));
        }
      } else {               // Q_z: x += P_y Q_z, y -= P_x Q_z
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const cplx px = park[k * NT], py = park[(8 + k) * NT];
          ax[k] = cmake(fma(sg * py.x, v[k].x, ax[k].x), fma(sg * py.y, v[k].y, ax[k].y));
          ay[k] = cmake(fma(-sg * px.x, v[k].x, ay[k].x), fma(-sg * px.y, v[k].y, ay[k].y));
        }
      }
    }
    // forward transforms of the three packed pairs and split into the half spectra
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      cplx acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const cplx s3 = c == 0 ? ax[k] : (c == 1 ? ay[k] : az[k]);
        acc[k] = cmake(s3.x * a.tmp, s3.y * a.tmp);
      }
      fft_regs<N, -1>(acc, t, smem, si, tw);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) smem[si(t + k * T)] = acc[k];
      __syncthreads();
      cplx* out = a.X[c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = t + k * T;
        if (kk <= N / 2 && active) {
          const cplx Zk = acc[k];
          const cplx Zn = smem[si((N - kk) & (N - 1))];
          out[rowA + kk] = cmake(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
          out[rowB + kk] = cmake(0.5 * (Zk.y + Zn.y), -0.5 * (Zk.x - Zn.x));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// zfwd_rk: one CTA = NP adjacent ky pencils of one kx.  Reads the nonlinear term in the exchange
// layout, continues it, transforms, filters and performs the RK update of one velocity component.
// ------------------------------------------------------------------------------------------
constexpr int kMaxDF = 10;

// out = v0 + dt*( cL*(lap ? -k^2 v : v) + sNL*filter(NL^ + ccoef*couple) + f )*rmp
//   HD/BOUSS velocity: cL = nu, lap, sNL = -1 (hd_rkstep2.f90:14-32); BOUSS adds the buoyancy / heat-current
//   coupling before the filter (bouss_rkstep2.f90:9-24); MHD potential: v holds J, cL = -mu, no lap,
//   sNL = +1 (mhd_rkstep2.f90:69-74)
struct ZfwdArgs {
  const cplx* nl;     // exchange layout [rank][kxl][zl][ky], physical rows
  const cplx* v;      // spectral field the linear term is taken from
  cplx* vout;         // result (may alias v)
  const cplx* v0;     // RK base
  const cplx* f;      // forcing
  const cplx* couple; // optional spectral field added to the nonlinear term before the filter
  double ccoef, cL, sNL;
  int lap;
  const ZMap* zmap;
  const double *kx, *ky, *kz;     // kx LOCAL
  const double *fx, *fy, *fz;     // filter factors (fx LOCAL)
  const double* dir;  // [C][d]
  int ny, nxl, nph, C, d;
  double dt, rmp;
};

// continuation rows of a pencil-fastest tile from the stashed boundary values
// bnd[q*NP + p]: q in [0,d) = f(1..d), q in [d,2d) = f(n-C-d+1..n-C)
template <int N, int NP>
__device__ __forceinline__ void fc_continue_tile(cplx (&v)[8], int j, int p, const cplx* bnd, int nph, int C, int d,
                                                 const double* __restrict__ dir) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e >= nph) {
      const int ii = e - nph;
      double ax = 0.0, ay = 0.0;
      for (int jj = 0; jj < d; ++jj) {
        const double w1 = __ldg(&dir[ii * d + jj]);
        const double w2 = __ldg(&dir[(C - 1 - ii) * d + jj]);
        const cplx f1 = bnd[(d + jj) * NP + p];
        const cplx f2 = bnd[(d - 1 - jj) * NP + p];
        ax = fma(w2, f2.x, fma(w1, f1.x, ax));
        ay = fma(w2, f2.y, fma(w1, f1.y, ay));
      }
      v[k] = cmake(ax, ay);
    }
  }
}

template <int N, int NP>
__device__ __forceinline__ void stash_boundary_tile(const cplx (&v)[8], int j, int p, cplx* bnd, int nph, int d) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e < d) bnd[e * NP + p] = v[k];
    if (e >= nph - d && e < nph) bnd[(d + e - (nph - d)) * NP + p] = v[k];
  }
}

template <int N, int NP, int MINB, bool HOIST, bool PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_zfwd_rk(ZfwdArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  cplx* bnd = smem + (size_t)NP * N;
  cplx* slot = bnd + (size_t)2 * kMaxDF * NP + threadIdx.x;
  ZMap* zm = reinterpret_cast<ZMap*>(bnd + (size_t)2 * kMaxDF * NP + (size_t)NP * N);
  for (int z = threadIdx.x; z < a.nph; z += NT) zm[z] = a.zmap[z];
  __syncthreads();
  const int tiles_y = cdiv(a.ny, NP), ntiles = tiles_y * a.nxl;
  auto issue = [&](int t) {
    if (!PF) return;
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = j + k * T;
      if (ky < a.ny && z < a.nph) {
        const ZMap m = zm[z];
        cp_async16(slot + k * NT, a.nl + m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky);
      } else {
        slot[k * NT] = cmake(0.0, 0.0);
      }
    }
    cp_async_commit();
  };
  int t = blockIdx.x;
  if (t < ntiles) issue(t);
  for (; t < ntiles; t += gridDim.x) {
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
    const bool active = ky < a.ny;
    cplx v[8];
    if (PF) {
      cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = slot[k * NT];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int z = j + k * T;
        v[k] = cmake(0.0, 0.0);
        if (active && z < a.nph) {
          const ZMap m = zm[z];
          v[k] = a.nl[m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky];
        }
      }
    }
    if (t + (int)gridDim.x < ntiles) issue(t + gridDim.x);
    // the three spectral pencils of the RK update: issued before the transform so that their latency
    // is covered by it (HOIST), or loaded at the point of use
    const size_t base = ((size_t)kxl * a.ny + (active ? ky : 0)) * N;
    cplx L[8], B[8], F[8];
    if (HOIST) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        L[k] = a.v[base + e];
        B[k] = a.v0[base + e];
        F[k] = a.f[base + e];
      }
    }
    __syncthreads();  // bnd and the exchange buffer of the previous tile are free
    stash_boundary_tile<N, NP>(v, j, p, bnd, a.nph, a.d);
    __syncthreads();
    fc_continue_tile<N, NP>(v, j, p, bnd, a.nph, a.C, a.d, a.dir);
    fft_regs<N, -1>(v, j, smem, SIdxPencil{p, NP}, tw);
    if (active) {
      const double x = __ldg(&a.kx[kxl]), y = __ldg(&a.ky[ky]);
      const double f1 = __ldg(&a.fx[kxl]), f2 = __ldg(&a.fy[ky]);
      const double kh2 = x * x + y * y;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        const double z = __ldg(&a.kz[e]), f3 = __ldg(&a.fz[e]);
        const double lm = a.lap ? -(kh2 + z * z) : 1.0;
        cplx NL = v[k];
        if (a.couple != nullptr) NL = caxpy(a.ccoef, a.couple[base + e], NL);
        NL = cscale(cscale(cscale(NL, f1), f2), f3);
        const cplx Lk = HOIST ? L[k] : a.v[base + e], Bk = HOIST ? B[k] : a.v0[base + e], Fk = HOIST ? F[k] : a.f[base + e];
        a.vout[base + e] = cmake(Bk.x + a.dt * (a.cL * (lm * Lk.x) + a.sNL * NL.x + Fk.x) * a.rmp,
                                 Bk.y + a.dt * (a.cL * (lm * Lk.y) + a.sNL * NL.y + Fk.y) * a.rmp);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// project: v_imposebc_and_project on NPB pencils per CTA, element-fastest mapping.  Everything is
// local to a (ky,kx) pencil: 2 z-IFFTs + no-slip rows + 2 continued z-FFTs (goto_domain_w_boundaries /
// noslip_z / goto_3d_fourier), the Poisson particular solution, the wall values of v_z, the
// closed-form Neumann harmonic correction, the new p' and the final subtraction.
// ------------------------------------------------------------------------------------------
struct ProjArgs {
  cplx *vx, *vy, *vz, *pr;
  const double *kx, *ky, *kz, *zc, *dir;
  long npencils;
  int ny, nph, C, d, has_mean;
  double Lz, tmp_noslip, inv_nz;
  double mx0, my0, mx1, my1;  // nx*ny*v_wall (mean mode rows)
};

template <int N>
__device__ __forceinline__ void fc_continue_elem(cplx (&v)[8], int j, const cplx* bnd, int nph, int C, int d,
                                                 const double* __restrict__ dir) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e >= nph) {
      const int ii = e - nph;
      double ax = 0.0, ay = 0.0;
      for (int jj = 0; jj < d; ++jj) {
        const double w1 = __ldg(&dir[ii * d + jj]);
        const double w2 = __ldg(&dir[(C - 1 - ii) * d + jj]);
        const cplx f1 = bnd[d + jj];
        const cplx f2 = bnd[d - 1 - jj];
        ax = fma(w2, f2.x, fma(w1, f1.x, ax));
        ay = fma(w2, f2.y, fma(w1, f1.y, ay));
      }
      v[k] = cmake(ax, ay);
    }
  }
}

template <int N>
__device__ __forceinline__ void stash_boundary_elem(const cplx (&v)[8], int j, cplx* bnd, int nph, int d) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e < d) bnd[e] = v[k];
    if (e >= nph - d && e < nph) bnd[d + e - (nph - d)] = v[k];
  }
}

// continuation + forward transform of a register-resident pencil whose physical rows are final
template <int N, class SI>
__device__ __forceinline__ void fc_fft_fwd(cplx (&v)[8], int j, cplx* smem, const SI& si, cplx* bnd, int nph, int C,
                                           int d, const double* __restrict__ dir, const cplx* __restrict__ tw) {
  __syncthreads();  // bnd may still be read by a previous continuation
  stash_boundary_elem<N>(v, j, bnd, nph, d);
  __syncthreads();
  fc_continue_elem<N>(v, j, bnd, nph, C, d, dir);
  fft_regs<N, -1>(v, j, smem, si, tw);
}

template <int N, int NPB, bool PF, int MINB>
__global__ void __launch_bounds__(NPB*(N / 8), MINB) k_project(ProjArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NPB * T;
  constexpr int XS = sidx_elem_stride<N>();
  const int pl = threadIdx.x / T, j = threadIdx.x % T;
  const SIdxElem si{pl * XS};
  // shared: [NPB exchange buffers][2 parked fields, thread-private][prefetch slots][boundary stashes][wall values]
  cplx* park = smem + (size_t)NPB * XS + threadIdx.x;                      // park[(c*8+k)*NT]
  cplx* slot = smem + (size_t)NPB * XS + (size_t)16 * NT + threadIdx.x;    // slot[k*NT]
  cplx* bnd = smem + (size_t)NPB * XS + (size_t)24 * NT + (size_t)pl * 2 * kMaxDF;
  cplx* wall = smem + (size_t)NPB * XS + (size_t)24 * NT + (size_t)NPB * 2 * kMaxDF + (size_t)pl * 2;
  const int top = a.nph - 1;
  const long ngroups = (a.npencils + NPB - 1) / NPB;
  auto issue = [&](long g, const cplx* field) {
    if (!PF) return;
    const long pencil = g * NPB + pl;
    if (pencil < a.npencils) {
      const cplx* src = field + (size_t)pencil * N + j;
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async16(slot + k * NT, src + k * T);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) slot[k * NT] = cmake(0.0, 0.0);
    }
    cp_async_commit();
  };
  long g = blockIdx.x;
  if (g < ngroups) issue(g, a.vx);
  for (; g < ngroups; g += gridDim.x) {
    const long pencil = g * NPB + pl;
    const bool active = pencil < a.npencils;
    const long pc = active ? pencil : 0;
    const int ky_i = (int)(pc % a.ny), kx_i = (int)(pc / a.ny);
    const size_t base = (size_t)pc * N;
    const double x = __ldg(&a.kx[kx_i]), y = __ldg(&a.ky[ky_i]);
    const bool mean = a.has_mean && pencil == 0;
    const cplx pr0 = a.pr[base], prT = a.pr[base + top];

    // ---- no-slip rows of vx, vy in the mixed domain, back to Fourier (vboundary.f90:116-145) ----
    cplx v[8];
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const double kc = c == 0 ? x : y;
      if (PF) {
        cp_async_wait_all();
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = slot[k * NT];
        issue(g, c == 0 ? a.vy : a.vz);
      } else {
        const cplx* src = c == 0 ? a.vx : a.vy;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = active ? src[base + j + k * T] : cmake(0.0, 0.0);
      }
      fft_regs<N, 1>(v, j, smem, si, tw);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        v[k] = cscale(v[k], a.inv_nz);
        if (e == 0 || e == top) {
          const cplx P = e == 0 ? pr0 : prT;
          v[k] = cmake(-kc * P.y * a.tmp_noslip, kc * P.x * a.tmp_noslip);
          if (mean) v[k] = cmake(e == 0 ? (c == 0 ? a.mx0 : a.my0) : (c == 0 ? a.mx1 : a.my1), 0.0);
        }
      }
      fc_fft_fwd<N>(v, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, tw);
#pragma unroll
      for (int k = 0; k < 8; ++k) park[(c * 8 + k) * NT] = v[k];
    }
    // ---- particular solution and its gradient (boundary_mod.fpp:405-448, 249-259) ----
    cplx dd[8], cz[8];
    if (PF) {
      cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 8; ++k) cz[k] = slot[k * NT];
      if (g + (long)gridDim.x < ngroups) issue(g + gridDim.x, a.vx);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) cz[k] = active ? a.vz[base + j + k * T] : cmake(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      const double z = __ldg(&a.kz[e]);
      const double kk2 = x * x + y * y + z * z;
      cplx A = park[k * NT], B = park[(8 + k) * NT];
      cplx Cc = cz[k];
      const cplx s = cmake(x * A.x + y * B.x + z * Cc.x, x * A.y + y * B.y + z * Cc.y);
      cplx D = cmake(s.y / kk2, -s.x / kk2);
      if ((mean && e == 0) || !active) D = cmake(0.0, 0.0);
      A = cmake(A.x + x * D.y, A.y - x * D.x);
      B = cmake(B.x + y * D.y, B.y - y * D.x);
      Cc = cmake(Cc.x + z * D.y, Cc.y - z * D.x);
      park[k * NT] = A;
      park[(8 + k) * NT] = B;
      cz[k] = Cc;
      dd[k] = D;
      v[k] = cscale(Cc, a.inv_nz);
    }
    // ---- wall values of v_z (boundary_mod.fpp:275-338) ----
    fft_regs<N, 1>(v, j, smem, si, tw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      if (e == 0) wall[0] = v[k];
      if (e == top) wall[1] = v[k];
    }
    __syncthreads();
    const cplx bc1 = wall[0], bc2 = wall[1];
    // ---- laplace_z, Neumann-Neumann (boundary_mod.fpp:531-560, 635-675) ----
    const double kh = sqrt(x * x + y * y);
    cplx c1, c2;
    if (mean) {
      c1 = bc1;
      c2 = cmake(0.0, 0.0);
    } else {
      const double e1 = exp(-kh * a.Lz), tt = 1.0 / (kh * (1.0 - exp(-2.0 * kh * a.Lz)));
      c1 = cmake((bc2.x - bc1.x * e1) * tt, (bc2.y - bc1.y * e1) * tt);
      c2 = cmake((-bc1.x + bc2.x * e1) * tt, (-bc1.y + bc2.y * e1) * tt);
    }
    // p' = IFFT_z(d)/nz + phi  (boundary_mod.fpp:371-380); all nz rows like the reference
    fft_regs<N, 1>(dd, j, smem, si, tw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      const double z = __ldg(&a.zc[e]);
      cplx A, B;
      if (mean) {
        A = cmake(c1.x * z + c2.x, 0.0);
        B = cmake(c1.x, 0.0);
      } else {
        const double ep = exp(kh * (z - a.Lz)), em = exp(-kh * z);
        A = cmake(c1.x * ep + c2.x * em, c1.y * ep + c2.y * em);
        B = cmake(kh * (c1.x * ep - c2.x * em), kh * (c1.y * ep - c2.y * em));
      }
      if (active) a.pr[base + e] = cmake(dd[k].x * a.inv_nz + A.x, dd[k].y * a.inv_nz + A.y);
      dd[k] = A;   // phi
      v[k] = B;    // d(phi)/dz
    }
    // ---- subtract the harmonic correction (boundary_mod.fpp:385-399) ----
    fc_fft_fwd<N>(v, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, tw);
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a.vz[base + j + k * T] = cmake(cz[k].x - v[k].x, cz[k].y - v[k].y);
    }
    fc_fft_fwd<N>(dd, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, tw);
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        const cplx A = park[k * NT], B = park[(8 + k) * NT], h = dd[k];
        a.vx[base + e] = cmake(A.x + x * h.y, A.y - x * h.x);
        a.vy[base + e] = cmake(B.x + y * h.y, B.y - y * h.x);
      }
    }
    __syncthreads();  // wall / bnd are rewritten by the next group
  }
}

// ==========================================================================================
// host side
// ==========================================================================================
static void range0(int n, int nprocs, int r, int* sta, int* cnt) {  // `range` on [0,n)
  const int w = n / nprocs, m = n % nprocs;
  *sta = r * w + (r < m ? r : m);
  *cnt = w + (m > r ? 1 : 0);
}

static void release_fields(std::vector<cplx*>& a, const std::vector<cplx*>* alias = nullptr) {
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i] && (!alias || i >= alias->size() || (*alias)[i] != a[i])) cudaFree(a[i]);
  a.clear();
}

int fused_free(Plan& p) {
  Fused* f = p.fused;
  if (!f) return 0;
  release_fields(f->R, &f->W);
  release_fields(f->W);
  release_fields(f->V);
  release_fields(f->X);
  release_fields(f->Uz, &f->U);
  release_fields(f->U);
  if (f->d_zmap) cudaFree(f->d_zmap);
  delete f;
  p.fused = nullptr;
  return 0;
}

static int fused_init(Plan& p, Fused** out) {
  if (p.fused) { *out = p.fused; return 0; }
  SX_REQUIRE(p.Cz > 0 && p.oz > 0 && p.oz <= kMaxDF, "the fused substep needs a non-periodic z direction (0 < oz <= 10)");
  Fused* f = new Fused();
  p.fused = f;
  f->nph = p.nz - p.Cz;
  range0(f->nph, p.nprocs, p.myrank, &f->zf0, &f->nzf);
  f->nxp = (p.nxh + 7) / 8 * 8;
  f->wsize = (size_t)p.nxl * f->nph * p.ny;
  f->vsize = (size_t)f->nzf * p.ny * f->nxp;
  std::vector<ZMap> zm(f->nph);
  long long base = 0;
  for (int r = 0; r < p.nprocs; ++r) {
    int s, c;
    range0(f->nph, p.nprocs, r, &s, &c);
    for (int q = 0; q < c; ++q) zm[s + q] = ZMap{base, c, q};
    base += (long long)p.nxl * c * p.ny;
  }
  f->z_displ.resize(p.nprocs); f->z_count.resize(p.nprocs); f->x_displ.resize(p.nprocs); f->x_count.resize(p.nprocs);
  for (int r = 0; r < p.nprocs; ++r) {
    int s, c, xs, xc;
    range0(f->nph, p.nprocs, r, &s, &c);
    range0(p.nxh, p.nprocs, r, &xs, &xc);
    f->z_displ[r] = c ? (size_t)zm[s].base : 0;
    f->z_count[r] = (size_t)p.nxl * c * p.ny;
    f->x_displ[r] = (size_t)xs * f->nzf * p.ny;
    f->x_count[r] = (size_t)xc * f->nzf * p.ny;
  }
  SX_CUDA_CHECK(cudaMalloc((void**)&f->d_zmap, zm.size() * sizeof(ZMap)));
  SX_CUDA_CHECK(cudaMemcpy(f->d_zmap, zm.data(), zm.size() * sizeof(ZMap), cudaMemcpyHostToDevice));
  *out = f;
  return 0;
}

// grow the work-field pools: nw transposed inverse fields, nv real-side inputs of the x pass, nx nonlinear terms
static int fused_reserve(Plan& p, Fused& f, int nw, int nv, int nx) {
  const size_t rsize = (size_t)p.nxh * f.nzf * p.ny;  // y-stage side [kx][zl][ky]
  const size_t both = f.wsize > rsize ? f.wsize : rsize;
  while ((int)f.W.size() < nw) {
    cplx *w = nullptr, *r = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&w, both * sizeof(cplx)));
    if (p.nprocs == 1) r = w;
    else SX_CUDA_CHECK(cudaMalloc((void**)&r, rsize * sizeof(cplx)));
    f.W.push_back(w);
    f.R.push_back(r);
  }
  while ((int)f.V.size() < nv) {
    cplx* v = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&v, f.vsize * sizeof(cplx)));
    f.V.push_back(v);
  }
  while ((int)f.X.size() < nx) {
    cplx *x = nullptr, *u = nullptr, *uz = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&x, f.vsize * sizeof(cplx)));
    SX_CUDA_CHECK(cudaMalloc((void**)&u, both * sizeof(cplx)));
    if (p.nprocs == 1) uz = u;
    else SX_CUDA_CHECK(cudaMalloc((void**)&uz, f.wsize * sizeof(cplx)));
    f.X.push_back(x);
    f.U.push_back(u);
    f.Uz.push_back(uz);
  }
  return 0;
}

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

template <int N> static int run_zinv(Plan& p, Fused& f, const cplx* in, cplx* out0, cplx* out1) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
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
  if (f.nzf == 0) return 0;
  YinvArgs a{in, out0, out1, p.d_ky, p.nxh, f.nxp, f.nzf};
  const cplx* tw = p.tw_y;
  int grid;
  if (p.knob_pf & 2) {
    auto kfn = k_yinv_tile<N, NP, MINB, true>;
    const size_t smem = (size_t)2 * NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(f.nxp, NP) * f.nzf, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_yinv_tile<N, NP, MINB, false>;
    const size_t smem = (size_t)NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(f.nxp, NP) * f.nzf, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YINV, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
template <int N> static int run_yfwd(Plan& p, Fused& f, const cplx* in, cplx* out) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  if (f.nzf == 0) return 0;
  YfwdArgs a{in, out, p.nxh, f.nxp, f.nzf};
  const cplx* tw = p.tw_y;
  int grid;
  if (p.knob_pf & 4) {
    auto kfn = k_yfwd_tile<N, NP, MINB, true>;
    const size_t smem = (size_t)2 * NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.nxh, NP) * f.nzf, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YFWD, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_yfwd_tile<N, NP, MINB, false>;
    const size_t smem = (size_t)NP * N * sizeof(cplx);
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.nxh, NP) * f.nzf, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_YFWD, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
template <int N, int LP, bool PF, int MINB, int NC> static int run_xpass_v(Plan& p, Fused& f, const double* d_kx_global) {
  constexpr int T = N / 8;
  if (f.nzf == 0) return 0;
  XpassArgs a;
  for (int i = 0; i < 3 * NC; ++i) a.V[i] = f.V[i];
  for (int i = 0; i < NC; ++i) a.X[i] = f.X[i];
  a.kx = d_kx_global;
  a.ny = p.ny;
  a.nxp = f.nxp;
  a.nzf = f.nzf;
  const double Ntot = (double)p.nx * (double)p.ny * (double)p.nz;
  a.tmp = 1.0 / (Ntot * Ntot);
  const cplx* tw = p.tw_x;
  auto kfn = k_xpass_gradre<N, LP, PF, MINB, NC>;
  const size_t smem = ((size_t)LP * sidx_elem_stride<N>() + (size_t)(PF ? 40 : 24) * LP * T) * sizeof(cplx);
  int grid;
  if (persistent_grid(p, kfn, LP * T, smem, cdiv(p.ny, 2 * LP) * f.nzf, &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), LP * T, smem, a, tw);
  return 0;
}
template <int N, int NC> static int run_xpass(Plan& p, Fused& f, const double* d_kx_global) {
  constexpr int T = N / 8;
  constexpr int LP = T >= 128 ? 1 : 128 / T;
  if constexpr (N == 512 && NC == 3) {
    switch (p.knob_xp) {
      case 1: return run_xpass_v<N, 1, true, 4, NC>(p, f, d_kx_global);
      case 2: return run_xpass_v<N, 2, true, 2, NC>(p, f, d_kx_global);
      case 3: return run_xpass_v<N, 2, false, 3, NC>(p, f, d_kx_global);
      case 4: return run_xpass_v<N, 1, false, 4, NC>(p, f, d_kx_global);
      default: break;
    }
  }
  if constexpr (N == 512) {   // measured best on B200 (profiles/r1h_knobs.md): one pair per CTA, direct loads
    if (p.knob_xp == 0) return run_xpass_v<N, 1, false, 6, NC>(p, f, d_kx_global);
  }
  return run_xpass_v<N, LP, true, (N <= 1024 ? 2 : 1), NC>(p, f, d_kx_global);
}
// X[xo..xo+2] = sum_pairs sgn * (V[P] x V[Q]) / N^2; Pi/Qi index the first of three consecutive V fields
template <int N> static int run_xcross(Plan& p, Fused& f, int npairs, const int* Pi, const int* Qi, const double* sgn, int xo) {
  constexpr int T = N / 8;
  constexpr int LP = T >= 128 ? 1 : 128 / T;
  if (f.nzf == 0) return 0;
  XcrossArgs a;
  for (int q = 0; q < 2; ++q)
    for (int c = 0; c < 3; ++c) {
      a.P[q][c] = f.V[Pi[q < npairs ? q : 0] + c];
      a.Q[q][c] = f.V[Qi[q < npairs ? q : 0] + c];
    }
  a.sgn[0] = sgn[0];
  a.sgn[1] = npairs > 1 ? sgn[1] : 0.0;
  a.npairs = npairs;
  for (int c = 0; c < 3; ++c) a.X[c] = f.X[xo + c];
  a.ny = p.ny;
  a.nxp = f.nxp;
  a.nzf = f.nzf;
  const double Ntot = (double)p.nx * (double)p.ny * (double)p.nz;
  a.tmp = 1.0 / (Ntot * Ntot);
  const cplx* tw = p.tw_x;
  auto kfn = k_xpass_cross<N, LP, (N <= 1024 ? 2 : 1)>;
  const size_t smem = ((size_t)LP * sidx_elem_stride<N>() + (size_t)40 * LP * T) * sizeof(cplx);
  int grid;
  if (persistent_grid(p, kfn, LP * T, smem, cdiv(p.ny, 2 * LP) * f.nzf, &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), LP * T, smem, a, tw);
  return 0;
}
struct RkTerm {   // see ZfwdArgs
  const cplx* couple = nullptr;
  double ccoef = 0.0, cL = 0.0, sNL = -1.0;
  int lap = 1;
};
template <int N> static int run_zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0,
                                        const cplx* frc, const RkTerm& rk, double dt, double rmp) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  ZfwdArgs a{nl, v, vout, v0, frc, rk.couple, rk.ccoef, rk.cL, rk.sNL, rk.lap, f.d_zmap, p.d_kx, p.d_ky, p.d_kz,
             p.d_fx, p.d_fy, p.d_fz, p.d_dir, p.ny, p.nxl, f.nph, p.Cz, p.oz, dt, rmp};
  const cplx* tw = p.tw_z;
  const size_t smem = ((size_t)2 * NP * N + (size_t)2 * kMaxDF * NP) * sizeof(cplx) + (size_t)N * sizeof(ZMap);
  int grid;
  if (N == 512 && p.knob_zf == 1) {
    auto kfn = k_zfwd_rk<N, NP, (N == 512 ? 2 : MINB), true, true>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZFWD_RK, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
    return 0;
  }
  if (p.knob_pf & 8) {
    auto kfn = k_zfwd_rk<N, NP, MINB, false, true>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZFWD_RK, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    auto kfn = k_zfwd_rk<N, NP, MINB, false, false>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZFWD_RK, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
template <int N, int NPB, bool PF, int MINB> static int run_project_v(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o,
                                        const double* zs, const double* ze) {
  constexpr int T = N / 8;
  double tmp = 1.0 / (double)o;
  if (o != p.ord) tmp = (double)(o + 1) * tmp;   // vboundary.f90:195-196
  const double sc = (double)p.nx * (double)p.ny;
  ProjArgs a{vx, vy, vz, pr, p.d_kx, p.d_ky, p.d_kz, p.d_z, p.d_dir, (long)p.ny * p.nxl,
             p.ny, f.nph, p.Cz, p.oz, p.ista == 1 ? 1 : 0, p.Lz, tmp, 1.0 / (double)p.nz,
             sc * (zs ? zs[0] : 0.0), sc * (zs ? zs[1] : 0.0), sc * (ze ? ze[0] : 0.0), sc * (ze ? ze[1] : 0.0)};
  const cplx* tw = p.tw_z;
  auto kfn = k_project<N, NPB, PF, MINB>;
  const size_t smem = ((size_t)NPB * sidx_elem_stride<N>() + (size_t)24 * NPB * T + (size_t)NPB * 2 * kMaxDF + (size_t)NPB * 2) * sizeof(cplx);
  const int ngroups = (int)((a.npencils + NPB - 1) / NPB);
  int grid;
  if (persistent_grid(p, kfn, NPB * T, smem, ngroups, &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_PROJECT, kfn, dim3(grid), NPB * T, smem, a, tw);
  return 0;
}
template <int N> static int run_project(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o,
                                        const double* zs, const double* ze) {
  constexpr int T = N / 8;
  constexpr int NPB = T >= 128 ? 1 : 128 / T;
  if constexpr (N == 512) {
    switch (p.knob_pj) {
      case 1: return run_project_v<N, 2, true, 3>(p, f, vx, vy, vz, pr, o, zs, ze);
      case 2: return run_project_v<N, 1, false, 8>(p, f, vx, vy, vz, pr, o, zs, ze);
      case 3: return run_project_v<N, 2, false, 4>(p, f, vx, vy, vz, pr, o, zs, ze);
      case 4: return run_project_v<N, 1, true, 8>(p, f, vx, vy, vz, pr, o, zs, ze);
      default: break;
    }
  }
  if constexpr (N == 512) {   // measured best on B200 (profiles/r1h_knobs.md): one pencil per CTA
    if (p.knob_pj == 0) return run_project_v<N, 1, true, 6>(p, f, vx, vy, vz, pr, o, zs, ze);
  }
  return run_project_v<N, NPB, true, (N <= 512 ? 3 : 1)>(p, f, vx, vy, vz, pr, o, zs, ze);
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

static int zinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1) {
#define C_(N) run_zinv<N>(p, f, in, o0, o1)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}
static int yinv(Plan& p, Fused& f, const cplx* in, cplx* o0, cplx* o1) {
#define C_(N) run_yinv<N>(p, f, in, o0, o1)
  SX_SIZE_SWITCH(p.ny, C_);
#undef C_
}
static int yfwd(Plan& p, Fused& f, const cplx* in, cplx* out) {
#define C_(N) run_yfwd<N>(p, f, in, out)
  SX_SIZE_SWITCH(p.ny, C_);
#undef C_
}
template <int NC> static int xpass(Plan& p, Fused& f, const double* kxg) {
#define C_(N) run_xpass<N, NC>(p, f, kxg)
  SX_SIZE_SWITCH(p.nx, C_);
#undef C_
}
static int xcross(Plan& p, Fused& f, int npairs, const int* Pi, const int* Qi, const double* sgn, int xo) {
#define C_(N) run_xcross<N>(p, f, npairs, Pi, Qi, sgn, xo)
  SX_SIZE_SWITCH(p.nx, C_);
#undef C_
}
static int zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0, const cplx* frc,
                   const RkTerm& rk, double dt, double rmp) {
#define C_(N) run_zfwd_rk<N>(p, f, nl, v, vout, v0, frc, rk, dt, rmp)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}
static int project(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o, const double* zs, const double* ze) {
#define C_(N) run_project<N>(p, f, vx, vy, vz, pr, o, zs, ze)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}

// W[slot] (z side, [rank][kxl][zl][ky]) -> R[slot] (xy side, [kx][zl][ky]); event slots 0..15
static int to_real_begin(Plan& p, Fused& f, int slot) {
  if (p.nprocs == 1) return 0;
  return exchange_begin(p, slot, f.W[slot], f.R[slot], f.z_displ.data(), f.z_count.data(), f.x_displ.data(), f.x_count.data());
}
// U[slot] (xy side) -> Uz[slot] (z side); event slots 16..23
static int to_spec_begin(Plan& p, Fused& f, int slot) {
  if (p.nprocs == 1) return 0;
  return exchange_begin(p, 16 + slot, f.U[slot], f.Uz[slot], f.x_displ.data(), f.x_count.data(), f.z_displ.data(), f.z_count.data());
}
static int ex_wait(Plan& p, int ev) { return p.nprocs == 1 ? 0 : exchange_wait(p, ev); }

static int fused_begin(Plan& p, Fused** fp, int nw, int nv, int nx) {
  if (fused_init(p, fp)) return 1;
  SX_REQUIRE(comm_ready(p), "multi-rank plan without a communicator: call sx_plan_set_comm or sx_plan_set_comm_callbacks");
  return fused_reserve(p, **fp, nw, nv, nx);
}

// inverse half shared by HD and BOUSS: q_c -> V[c], V[NC+c] (dy), V[2NC+c] (dz)
template <int NC> static int gradient_fields_to_real(Plan& p, Fused& f, const cplx* const* q) {
  for (int c = 0; c < NC; ++c) {
    if (zinv(p, f, q[c], f.W[2 * c], f.W[2 * c + 1])) return 1;
    if (to_real_begin(p, f, 2 * c) || to_real_begin(p, f, 2 * c + 1)) return 1;
  }
  for (int c = 0; c < NC; ++c) {
    if (ex_wait(p, 2 * c) || ex_wait(p, 2 * c + 1)) return 1;
    if (yinv(p, f, f.R[2 * c], f.V[c], f.V[NC + c])) return 1;
    if (yinv(p, f, f.R[2 * c + 1], f.V[2 * NC + c], nullptr)) return 1;
  }
  return 0;
}
static int nonlinear_to_spectral_begin(Plan& p, Fused& f, int nx) {
  for (int c = 0; c < nx; ++c) {
    if (yfwd(p, f, f.X[c], f.U[c])) return 1;
    if (to_spec_begin(p, f, c)) return 1;
  }
  return 0;
}

// hd_rkstep2.f90:3-36.  st[0..2] v, st[3] pr, st[4..6] f, st[7..9] RK base.
int hd_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, const double* zs, const double* ze) {
  Fused* fp;
  if (fused_begin(p, &fp, 6, 9, 3)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  if (gradient_fields_to_real<3>(p, f, st)) return 1;
  if (xpass<3>(p, f, p.d_kxg)) return 1;
  if (nonlinear_to_spectral_begin(p, f, 3)) return 1;
  RkTerm rk;
  rk.cL = nu;
  for (int c = 0; c < 3; ++c) {
    if (ex_wait(p, 16 + c)) return 1;
    if (zfwd_rk(p, f, f.Uz[c], st[c], st[c], st[7 + c], st[4 + c], rk, dt, rmp)) return 1;
  }
  return project(p, f, st[0], st[1], st[2], st[3], o, zs, ze);
}

int s_imposebc(Plan& p, cplx* th);
int theta_roundtrip(Plan& p, cplx* th, cplx* out);
int a_imposebc_and_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph);

// bouss_rkstep2.f90:3-59.  st[0..2] v, 3 pr, 4..6 f, 7..9 C1..C3, 10 th, 11 fs, 12 C7, 13 scratch.
int bouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                        const double* zs, const double* ze) {
  Fused* fp;
  if (fused_begin(p, &fp, 8, 12, 4)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  const cplx* q[4] = {st[0], st[1], st[2], st[10]};
  if (gradient_fields_to_real<4>(p, f, q)) return 1;
  if (xpass<4>(p, f, p.d_kxg)) return 1;          // gradre (3) and advect (1) in one pass
  if (nonlinear_to_spectral_begin(p, f, 4)) return 1;
  // theta first, into the scratch field: it reads the not yet updated v_z (heat current), and v_z below reads
  // the not yet updated theta (buoyancy)
  RkTerm rt;
  rt.cL = kappa; rt.couple = st[2]; rt.ccoef = -xtemp;
  if (ex_wait(p, 16 + 3)) return 1;
  if (zfwd_rk(p, f, f.Uz[3], st[10], st[13], st[12], st[11], rt, dt, rmp)) return 1;
  for (int c = 0; c < 3; ++c) {
    RkTerm rk;
    rk.cL = nu;
    if (c == 2) { rk.couple = st[10]; rk.ccoef = -xmom; }
    if (ex_wait(p, 16 + c)) return 1;
    if (zfwd_rk(p, f, f.Uz[c], st[c], st[c], st[7 + c], st[4 + c], rk, dt, rmp)) return 1;
  }
  if (project(p, f, st[0], st[1], st[2], st[3], o, zs, ze)) return 1;
  // s_imposebc, fc_filter and the theta round trip (bouss_rkstep2.f90:53-59); all pencil-local
  return s_imposebc(p, st[13]) || op_fc_filter(p, st[13]) || theta_roundtrip(p, st[13], st[10]);
}

// mhd_rkstep2.f90:3-84.  st[0..2] v, 3 pr, 4..6 f, 7..9 C1..C3, 10..12 a, 13 ph, 14..16 m, 17..19 C9..C11.
int mhd_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double mu, const double* b0) {
  Fused* fp;
  if (fused_begin(p, &fp, 12, 12, 6)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  cplx *B[3], *Wv[3];
  for (int c = 0; c < 3; ++c)
    if (plan_cwork(p, 9 + c, &B[c]) || plan_cwork(p, 12 + c, &Wv[c])) return 1;
  cplx *ax = st[10], *ay = st[11], *az = st[12];
  // B = curl A (+ uniform field at the mean mode), J = curl B written over A, omega = curl v   (:6-20, prodre)
  if (op_curlk(p, ay, az, B[0], 1) || op_curlk(p, ax, az, B[1], 2) || op_curlk(p, ax, ay, B[2], 3)) return 1;
  if (p.ista == 1)
    for (int c = 0; c < 3; ++c)
      if (op_set_elem(p, B[c], 0, (b0 ? b0[c] : 0.0) * N, 0.0)) return 1;
  if (op_curlk(p, B[1], B[2], ax, 1) || op_curlk(p, B[0], B[2], ay, 2) || op_curlk(p, B[0], B[1], az, 3)) return 1;
  if (op_curlk(p, st[1], st[2], Wv[0], 1) || op_curlk(p, st[0], st[2], Wv[1], 2) || op_curlk(p, st[0], st[1], Wv[2], 3)) return 1;
  // twelve plain fields to real space: v -> V[0..2], omega -> V[3..5], B -> V[6..8], J -> V[9..11]
  const cplx* q[12] = {st[0], st[1], st[2], Wv[0], Wv[1], Wv[2], B[0], B[1], B[2], ax, ay, az};
  for (int c = 0; c < 12; ++c) {
    if (zinv(p, f, q[c], f.W[c], nullptr)) return 1;
    if (to_real_begin(p, f, c)) return 1;
  }
  for (int c = 0; c < 12; ++c) {
    if (ex_wait(p, c)) return 1;
    if (yinv(p, f, f.R[c], f.V[c], nullptr)) return 1;
  }
  // X[0..2] = omega x v - J x B = -(v x omega) + (B x J);  X[3..5] = v x B
  {
    const int Pi[2] = {0, 6}, Qi[2] = {3, 9};
    const double sg[2] = {-1.0, 1.0};
    if (xcross(p, f, 2, Pi, Qi, sg, 0)) return 1;
    const int Pe[1] = {0}, Qe[1] = {6};
    const double se[1] = {1.0};
    if (xcross(p, f, 1, Pe, Qe, se, 3)) return 1;
  }
  if (nonlinear_to_spectral_begin(p, f, 6)) return 1;
  for (int c = 0; c < 3; ++c) {
    RkTerm rk;
    rk.cL = nu;
    if (ex_wait(p, 16 + c)) return 1;
    if (zfwd_rk(p, f, f.Uz[c], st[c], st[c], st[7 + c], st[4 + c], rk, dt, rmp)) return 1;
  }
  for (int c = 0; c < 3; ++c) {
    RkTerm rk;
    rk.cL = -mu; rk.lap = 0; rk.sNL = 1.0;
    if (ex_wait(p, 16 + 3 + c)) return 1;
    if (zfwd_rk(p, f, f.Uz[3 + c], st[10 + c], st[10 + c], st[17 + c], st[14 + c], rk, dt, rmp)) return 1;
  }
  if (project(p, f, st[0], st[1], st[2], st[3], o, nullptr, nullptr)) return 1;
  return a_imposebc_and_project(p, ax, ay, az, st[13]);
}

}  // namespace sx
