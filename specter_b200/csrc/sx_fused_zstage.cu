// Fused substep, z stage in ONE kernel: for a tile of NP adjacent ky pencils of one kx, the three velocity components
// go through FC-Gram continuation + z-FFT + fc_filter + Laplacian + RK update (what k_zfwd_rk does per component:
// fftp.fpp:757-780, pseudospec_hd.f90:1099-1112, hd_rkstep2.f90:14-32) and, still on chip, through the whole
// v_imposebc_and_project (vboundary.f90:116-148, boundary_mod.fpp:197-402) that k_project_bulk does per pencil.
//
// Against the two-kernel form (3 x zfwd_rk + project) this removes the write and the re-read of the updated velocity
// between the kernels (6 F of the 62 F the substep moved, F = one full-field read or write), and it needs fewer
// transforms per pencil:
//   * transforms run in PAIRS (fft_regs2: two pencils per thread share every barrier and the twiddle powers):
//     (NL_x, NL_y), NL_z, (IFFT v_x, IFFT v_y), (FFT v_x, FFT v_y), (IFFT d, FFT e) -- nine transforms, five rounds;
//   * the harmonic correction phi = c1 e^{kh (z - Lz)} + c2 e^{-kh z} and its derivative are linear in the two
//     exponentials, and the continuation and the transform are linear, so phi^ = c1 E+^ + c2 E-^ and
//     phi'^ = kh (c1 E+^ - c2 E-^) with E-^ = FFT(cont(e^{-kh z})): ONE real-input transform instead of the two
//     complex ones of boundary_mod.fpp:385-399.  E+ is the mirror image of E- on the grid (z_k = k dz, Lz = z_top) and
//     the FC-Gram continuation commutes with that reflection (fftp.fpp:760-770 is symmetric under ii -> C-ii+1 with
//     the two boundary stencils swapped), hence E+^(k) = conj(e^{2 pi i k top / N} E-^(k)): no second transform;
//   * the wall values of v_z are two dot products with a block reduction (rows 0 and top of the backward transform).
// The mean pencil (kx = ky = 0) has phi = Re(c1) z + Re(c2): its phi^ multiplies kx = ky = 0 and drops out; its
// phi'^ = Re(c1) FFT(cont(1)) takes the place of E-^ in the last pair.
//
// Thread mapping as in the tile kernels: lane-fastest over the NP pencils (p = tid % NP, j = tid / NP), eight
// elements j + k T per thread; the nonlinear-term tiles (64-byte pieces in the exchange layout) arrive through
// thread-private cp.async slots one to two transforms ahead, the spectral pencils of the RK update are read straight
// into registers (the two CTAs of an SM cover each other's load latency).
#include "sx_fused.h"

namespace sx {

struct ZstageArgs {
  const cplx* nl[3];      // nonlinear terms, exchange layout [rank][kxl][zl][ky], physical rows
  cplx* v[3];             // velocity: linear-term input and result
  const cplx* v0[3];      // RK base
  const cplx* f[3];       // forcing
  const cplx* couple[3];  // optional field added (times ccoef) to the nonlinear term before the filter
  double ccoef[3];
  cplx* pr;               // p' in the mixed domain: wall rows read, all rows written
  double cL, sNL;
  int lap;
  const ZMap* zmap;
  const double *kx, *ky, *kz;   // kx LOCAL
  const double *fx, *fy, *fz;   // filter factors (fx LOCAL)
  const double *dir, *zc;
  double dkz;                   // kz(e) = (e < N/2 ? e : e - N) * dkz (specter.fpp:772-789)
  int ny, nxl, nph, C, d, has_mean;
  double dt, rmp, Lz, tmp_noslip, inv_nz;
  double mx0, my0, mx1, my1;    // nx*ny*v_wall (mean mode rows)
  cplx phT[8];                  // exp(+2 pi i k top / 8)
};

// Geometry: a CTA works on NP adjacent ky pencils with TWO thread groups of NP * N/8 threads; in every round each
// group runs one transform per pencil (group 0: v_z / v_x / d, group 1: idle / v_y / e^{-kh z}), so the pair of
// transforms of a round is spread over twice the warps instead of doubling the work of each thread: at 128 registers
// two 256-thread CTAs (N = 512, NP = 2) keep 16 warps per SM busy, which is what hides the shared-memory and global
// latencies of a transform-bound kernel (the first version paired the transforms inside each thread at 8 warps per SM:
// 7.3 ms against 5.8 ms of the separate kernels, stalls short_scoreboard + long_scoreboard + wait = 68 %, profiles/r2e).
template <int N, int NP> struct ZstageGeo {
  static constexpr int T = N / 8, NG = NP * T, NT = 2 * NG, NWG = (NG + 31) / 32;
  static constexpr int XS = N + N / 8 + 8;        // exchange buffer of one pencil: element-fastest, one pad slot per 8 elements
  static constexpr int POFF = NP >= 8 ? 1 : 8 / NP;   // extra offset of pencil p: the NP lanes of a quarter-warp hit distinct banks
  static constexpr int NRED = NP >= 32 ? T : NWG;     // rows of wall-sum partials per pencil
  static constexpr size_t cplx_elems = (size_t)2 * NP * XS     // exchange buffers of the two groups
                                       + (size_t)4 * 8 * NG     // units PC, SZ, SX, SY: element k of thread t of a group at [k * NG + t]
                                       + (size_t)2 * 2 * kMaxDF * NP   // boundary rows of the two groups' arrays
                                       + (size_t)2 * NRED * NP;  // wall-row partial sums
  static size_t smem_bytes(int C) { return (cplx_elems + (size_t)2 * C * NP) * sizeof(cplx); }
};

// Rounds of a tile (every barrier is executed by both groups):
//   round 0: g0 NL_z                      -> continuation, FFT, RK update -> v_z parked (g1 only keeps the barriers)
//   round 1: g0 NL_x, g1 NL_y             -> continuation, FFT, RK update
//   round 2: g0 v_x,  g1 v_y              -> backward FFT, no-slip wall rows
//   round 3: g0 v_x,  g1 v_y              -> continuation, FFT; g1: Poisson particular solution d and the gradient
//                                            subtracted from v_x, v_y; g0: v_z minus its part, wall rows of v_z; c1 / c2
//   round 4: g0 d,    g1 e^{-kh z}        -> g0 backward FFT (p'), g1 continued FFT -> the final fields
// The loop body holds ONE copy of the forward transform (a backward transform is the forward one between two
// conjugations): five unrolled instances made 230 KB of instructions and 41 % stall_no_instructions (profiles/r2c).
template <int N, int NP, int MINB, int L2PF>
__global__ void __launch_bounds__(2 * NP*(N / 8), MINB) k_zstage(ZstageArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  typedef ZstageGeo<N, NP> G;
  constexpr int T = G::T, NG = G::NG, XS = G::XS;
  const int g = threadIdx.x / NG, tg = threadIdx.x - g * NG;     // group, thread inside the group
  const int p = tg % NP, j = tg / NP;
  const int lane = threadIdx.x & 31, wg = tg >> 5;
  TwRegs<N> twr;
  twr.load(tw, j);
  const SIdxElem si{(g * NP + p) * XS + (p % 8) * G::POFF};   // the offset stays inside the 8 spare slots of XS
  cplx* ex = smem;
  cplx* PC = smem + (size_t)2 * NP * XS + tg;   // v_z park
  cplx* SZ = PC + (size_t)8 * NG;               // NL_z prefetch slot, then d on its way from g1 to g0
  cplx* SX = SZ + (size_t)8 * NG;               // NL_x prefetch slot, then v_x
  cplx* SY = SX + (size_t)8 * NG;               // NL_y prefetch slot, then v_y
  cplx* bnd = smem + (size_t)2 * NP * XS + (size_t)4 * 8 * NG;   // [(g * 2d + q) * NP + p]
  cplx* cv = bnd + (size_t)4 * kMaxDF * NP;                       // [(g * C + ii) * NP + p]
  cplx* red = cv + (size_t)2 * a.C * NP;                          // [(row * NP + p) * 2 + {0,1}]
  const int top = a.nph - 1;
  // phase of the top-wall row at this thread's first element, e^{+2 pi i j top / N}
  cplx phj;
  {
    double sn, cs;
    sincospi(2.0 * (double)(((long)j * top) % N) / (double)N, &sn, &cs);
    phj = cmake(cs, sn);
  }
  const double zj = __ldg(&a.zc[j]), z7 = __ldg(&a.zc[j + 7 * T]), dzT = __ldg(&a.zc[T]) - __ldg(&a.zc[0]);
  const int tiles_y = cdiv(a.ny, NP), ntiles = tiles_y * a.nxl;
  // one nonlinear term of tile t into a slot of this group (one commit group); a rolled loop on purpose (code size)
  auto issue = [&](cplx* slot, int t, const cplx* __restrict__ nl) {
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      const int z = j + k * T;
      if (ky < a.ny && z < a.nph) {
        const ZMap m = a.zmap[z];
        cp_async16(slot + k * NG, nl + m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky);
      } else {
        slot[k * NG] = cmake(0.0, 0.0);
      }
    }
    cp_async_commit();
  };

  int t = blockIdx.x;
  if (t < ntiles && g == 0) issue(SZ, t, a.nl[2]);
  for (; t < ntiles; t += gridDim.x) {
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
    const bool active = ky < a.ny;
    const int tn = t + gridDim.x;
    const size_t base = ((size_t)kxl * a.ny + (active ? ky : 0)) * N;
    const double x = __ldg(&a.kx[kxl]), y = __ldg(&a.ky[active ? ky : 0]);
    const double f1 = __ldg(&a.fx[kxl]), f2 = __ldg(&a.fy[active ? ky : 0]);
    const double kh2 = x * x + y * y;
    const double kh = sqrt(kh2);
    const bool mean = a.has_mean && kxl == 0 && ky == 0;
    if (L2PF && threadIdx.x < 9 && tn < ntiles) {
      // the spectral pencils of the next tile's RK updates (NP adjacent ky pencils are one contiguous range)
      const int ky0 = (tn % tiles_y) * NP, kxn = tn / tiles_y;
      const int np = a.ny - ky0 < NP ? a.ny - ky0 : NP;
      const size_t tb = ((size_t)kxn * a.ny + ky0) * N;
      const int c = threadIdx.x / 3, q = threadIdx.x % 3;
      const cplx* fld = q == 0 ? a.v[c] : (q == 1 ? a.v0[c] : a.f[c]);
      l2_prefetch(fld + tb, (unsigned)((size_t)np * N * sizeof(cplx)));
    }
    cplx c1 = cmake(0.0, 0.0), c2 = cmake(0.0, 0.0);   // laplace_z coefficients, set in round 3
    double st = 0.0, em = 0.0;                          // e^{-kh dz T}, e^{-kh z_j}
    cplx v[8];

#pragma unroll 1
    for (int r = 0; r < 5; ++r) {
      const bool work = !(r == 0 && g == 1);            // this group's transform of the round carries data
      const bool cont = (r == 0 && g == 0) || r == 1 || r == 3 || (r == 4 && g == 1);
      const bool inv = r == 2 || (r == 4 && g == 0);
      const int c = r == 0 ? 2 : g;                     // component of the RK rounds
      cplx L[8];
      // ---- inputs of the round ---------------------------------------------------------------------------
      if (r <= 1) {
        if (work) {
          const cplx* __restrict__ lin = a.v[c];        // linear-term pencil: travels under the transform
#pragma unroll
          for (int k = 0; k < 8; ++k) L[k] = lin[base + j + k * T];
        }
        cp_async_wait_all();
        if (r == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = g == 0 ? SZ[k * NG] : cmake(0.0, 0.0);
          issue(g == 0 ? SX : SY, t, a.nl[g]);          // the v_x / v_y terms travel under this round
        } else {
          const cplx* src = g == 0 ? SX : SY;
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = src[k * NG];
        }
      } else if (r == 4) {
        if (g == 0) {
          if (tn < ntiles) issue(SZ, tn, a.nl[2]);      // the next tile's first term travels under the last round
        } else {
          // the real sequence whose continued transform carries the harmonic correction: e^{-kh z} (mean pencil:
          // the constant Re(c1) = phi'), physical rows; geometric in the thread's stride
          double q = em;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[k] = cmake(mean ? c1.x : q, 0.0);
            q *= st;
          }
        }
      }
      if (inv) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k].y = -v[k].y;
      }
      if (r != 2) {
        // FC-Gram continuation (fftp.fpp:757-772): boundary rows to shared memory, C * NP threads per group form one
        // row each, every thread picks up the rows it owns
        if (cont) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            if (e < a.d) bnd[(g * 2 * a.d + e) * NP + p] = v[k];
            if (e >= a.nph - a.d && e < a.nph) bnd[(g * 2 * a.d + a.d + e - (a.nph - a.d)) * NP + p] = v[k];
          }
        }
        __syncthreads();
        if (cont) {
#pragma unroll 1
          for (int u = tg; u < a.C * NP; u += NG) {
            const int ii = u / NP, pp = u - ii * NP;
            const cplx* bq = bnd + (size_t)g * 2 * a.d * NP + pp;
            double ax = 0.0, ay = 0.0;
            for (int jj = 0; jj < a.d; ++jj) {
              const double w1 = __ldg(&a.dir[ii * a.d + jj]);
              const double w2 = __ldg(&a.dir[(a.C - 1 - ii) * a.d + jj]);
              const cplx q1 = bq[(a.d + jj) * NP], q2 = bq[(a.d - 1 - jj) * NP];
              ax = fma(w2, q2.x, fma(w1, q1.x, ax));
              ay = fma(w2, q2.y, fma(w1, q1.y, ay));
            }
            cv[(g * a.C + ii) * NP + pp] = cmake(ax, ay);
          }
        }
        __syncthreads();
        if (cont) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            if (e >= a.nph) v[k] = cv[(g * a.C + e - a.nph) * NP + p];
          }
        }
      }
      fft_regs<N, -1>(v, j, ex, si, twr);
      if (inv) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k].y = -v[k].y;
      }
      // ---- results of the round ----------------------------------------------------------------------------
      if (r <= 1) {
        if (work) {
          // RK update (k_zfwd_rk, BATCH == 2 association)
          const cplx* __restrict__ cpl = a.couple[c];
          const cplx* __restrict__ frc = a.f[c];
          const cplx* __restrict__ rkb = a.v0[c];
          if (cpl != nullptr) {
            const double cc = a.ccoef[c];
            cplx Q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) Q[k] = cpl[base + j + k * T];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = caxpy(cc, Q[k], v[k]);
          }
          // register budget (128): the forcing travels while the linear and nonlinear terms are combined, the RK base
          // is requested once the linear-term pencil is dead
          cplx F[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) F[k] = frc[base + j + k * T];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            const double z = (double)(e < N / 2 ? e : e - N) * a.dkz;   // the product the host table holds
            const double f3 = __ldg(&a.fz[e]);
            const double lm = a.lap ? -(kh2 + z * z) : 1.0;
            const cplx NL = cscale(cscale(cscale(v[k], f1), f2), f3);
            v[k] = cmake(a.cL * (lm * L[k].x) + a.sNL * NL.x, a.cL * (lm * L[k].y) + a.sNL * NL.y);
          }
          cplx B[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) B[k] = rkb[base + j + k * T];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = cmake((v[k].x + F[k].x) * a.dt * a.rmp, (v[k].y + F[k].y) * a.dt * a.rmp);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = cmake(B[k].x + v[k].x, B[k].y + v[k].y);
          if (r == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) PC[k * NG] = v[k];
          }
        }
      } else if (r == 2) {
        // no-slip rows of v_x, v_y in the mixed domain (vboundary.f90:116-145)
        const double kc = g == 0 ? x : y;
        const cplx pr0 = a.pr[base], prT = a.pr[base + top];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          v[k] = cscale(v[k], a.inv_nz);
          if (e == 0 || e == top) {
            const cplx P = e == 0 ? pr0 : prT;
            v[k] = cmake(-kc * P.y * a.tmp_noslip, kc * P.x * a.tmp_noslip);
            if (mean) v[k] = cmake(e == 0 ? (g == 0 ? a.mx0 : a.my0) : (g == 0 ? a.mx1 : a.my1), 0.0);
          }
        }
      } else if (r == 3) {
        // particular solution and its gradient (boundary_mod.fpp:405-448, 249-259)
        if (g == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) SX[k * NG] = v[k];
        }
        __syncthreads();
        if (g == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            const double z = (double)(e < N / 2 ? e : e - N) * a.dkz;
            const double kk2 = x * x + y * y + z * z;
            const cplx A = SX[k * NG], Cc = PC[k * NG];
            const cplx s = cmake(x * A.x + y * v[k].x + z * Cc.x, x * A.y + y * v[k].y + z * Cc.y);
            const double ik2 = 1.0 / kk2;    // one reciprocal instead of two divisions per element (1 ulp)
            cplx D = cmake(s.y * ik2, -s.x * ik2);
            if ((mean && e == 0) || !active) D = cmake(0.0, 0.0);
            SX[k * NG] = cmake(A.x + x * D.y, A.y - x * D.x);
            SY[k * NG] = cmake(v[k].x + y * D.y, v[k].y - y * D.x);
            SZ[k * NG] = D;      // the NL_z slot is free until g0 has read d back and issues the next tile's term
          }
        }
        __syncthreads();
        cplx s0 = cmake(0.0, 0.0), s1 = cmake(0.0, 0.0);
        if (g == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            const double z = (double)(e < N / 2 ? e : e - N) * a.dkz;
            const cplx D = SZ[k * NG];
            cplx Cc = PC[k * NG];
            Cc = cmake(Cc.x + z * D.y, Cc.y - z * D.x);
            PC[k * NG] = Cc;
            v[k] = D;
            // wall values of v_z (boundary_mod.fpp:275-338): rows 0 and top of IFFT_z(v_z)/nz
            s0 = cadd(s0, Cc);
            s1 = cadd(s1, cmul(Cc, a.phT[k]));
          }
          s1 = cmul(s1, phj);
        }
        if (NP < 32) {   // executed by both groups (group 1 carries zeros): warp-uniform, and uniform under the CPU emulation
#pragma unroll
          for (int o = NP; o < 32; o <<= 1) {
            s0.x += __shfl_xor_sync(0xffffffffu, s0.x, o);
            s0.y += __shfl_xor_sync(0xffffffffu, s0.y, o);
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o);
            s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
          }
        }
        // one row of partials per warp of group 0 (NP >= 32: a warp is one element row j of 32 pencils)
        if (g == 0 && (NP >= 32 || lane < NP)) {
          const int row = NP >= 32 ? j : wg;
          red[(row * NP + p) * 2] = s0;
          red[(row * NP + p) * 2 + 1] = s1;
        }
        __syncthreads();   // partial sums visible
        cplx bc1 = cmake(0.0, 0.0), bc2 = cmake(0.0, 0.0);
#pragma unroll 1
        for (int q = 0; q < G::NRED; ++q) {
          bc1 = cadd(bc1, red[(q * NP + p) * 2]);
          bc2 = cadd(bc2, red[(q * NP + p) * 2 + 1]);
        }
        bc1 = cscale(bc1, a.inv_nz);
        bc2 = cscale(bc2, a.inv_nz);
        // laplace_z, Neumann-Neumann (boundary_mod.fpp:531-560, 635-675)
        if (mean) {
          c1 = bc1;
          c2 = cmake(0.0, 0.0);
        } else {
          const double e1 = exp(-kh * a.Lz), tt = 1.0 / (kh * (1.0 - e1 * e1));
          c1 = cmake((bc2.x - bc1.x * e1) * tt, (bc2.y - bc1.y * e1) * tt);
          c2 = cmake((-bc1.x + bc2.x * e1) * tt, (-bc1.y + bc2.y * e1) * tt);
          st = exp(-kh * dzT);
          em = exp(-kh * zj);
        }
      } else if (g == 0) {
        // p' = IFFT_z(d)/nz + phi (boundary_mod.fpp:371-380)
        double ep[8];
        if (!mean) {
          ep[7] = exp(kh * (z7 - a.Lz));
#pragma unroll
          for (int k = 6; k >= 0; --k) ep[k] = ep[k + 1] * st;
        }
        double q = em;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          cplx ph;
          if (mean) {
            const double z = __ldg(&a.zc[e]);
            ph = cmake(c1.x * z + c2.x, 0.0);
          } else {
            ph = cmake(c1.x * ep[k] + c2.x * q, c1.y * ep[k] + c2.y * q);
            q *= st;
          }
          if (active) a.pr[base + e] = cmake(v[k].x * a.inv_nz + ph.x, v[k].y * a.inv_nz + ph.y);
        }
      } else if (active) {
        // the harmonic correction subtracted (boundary_mod.fpp:385-399): the final fields
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          cplx h, hz;   // phi^, phi'^
          if (mean) {
            h = cmake(0.0, 0.0);
            hz = v[k];
          } else {
            const cplx Em = v[k];
            const cplx Ep = cconj(cmul(cmul(a.phT[k], phj), Em));
            const cplx t1 = cmul(c1, Ep), t2 = cmul(c2, Em);
            h = cadd(t1, t2);
            hz = cscale(csub(t1, t2), kh);
          }
          const cplx A = SX[k * NG], B = SY[k * NG], Cc = PC[k * NG];
          a.v[0][base + e] = cmake(A.x + x * h.y, A.y - x * h.x);
          a.v[1][base + e] = cmake(B.x + y * h.y, B.y - y * h.x);
          a.v[2][base + e] = cmake(Cc.x - hz.x, Cc.y - hz.y);
        }
      }
    }
    __syncthreads();   // the parks are rewritten by the next tile's rounds
  }
}

template <int N, int NP, int MINB, int L2PF>
static int run_zstage_v(Plan& p, Fused& f, const ZstageArgs& a) {
  typedef ZstageGeo<N, NP> G;
  const cplx* tw = p.tw_z;
  auto kfn = k_zstage<N, NP, MINB, L2PF>;
  const size_t smem = G::smem_bytes(p.Cz);
  int grid;
  if (persistent_grid(p, kfn, G::NT, smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_ZSTAGE, kfn, dim3(grid), G::NT, smem, a, tw);
  return 0;
}

template <int N> static int run_zstage(Plan& p, Fused& f, ZstageArgs& a) {
  // pencils per tile: two at the long lengths (two CTAs per SM at 86 KB each cover each other's load latency and
  // barriers); the short lengths keep the tile kernels' geometry
  if constexpr (N == 512) {
    switch (p.knob_zs) {
      case 2: return run_zstage_v<N, 4, 1, 0>(p, f, a);
      case 3: return run_zstage_v<N, 4, 1, 1>(p, f, a);
      case 4: return run_zstage_v<N, 2, 2, 1>(p, f, a);
      default: return run_zstage_v<N, 2, 2, 0>(p, f, a);
    }
  } else if constexpr (N >= 1024) {
    return run_zstage_v<N, (N == 1024 ? 2 : 1), 1, 0>(p, f, a);
  } else {
    return run_zstage_v<N, (TileNP<N>::value > 1 ? TileNP<N>::value / 2 : 1), 1, 0>(p, f, a);
  }
}

bool zstage_enabled(const Plan& p) { return p.knob_zs != 1; }

// velocity part of a substep: nl[c] -> RK update of v[c] (optional coupling fields) -> v_imposebc_and_project
int fused_zstage(Plan& p, Fused& f, const cplx* const* nl, cplx* const* v, const cplx* const* v0, const cplx* const* frc,
                 const RkTerm* rk, cplx* pr, int o, double dt, double rmp, const double* zs, const double* ze) {
  ZstageArgs a;
  for (int c = 0; c < 3; ++c) {
    a.nl[c] = nl[c];
    a.v[c] = v[c];
    a.v0[c] = v0[c];
    a.f[c] = frc[c];
    a.couple[c] = rk[c].couple;
    a.ccoef[c] = rk[c].ccoef;
    SX_REQUIRE(rk[c].cL == rk[0].cL && rk[c].sNL == rk[0].sNL && rk[c].lap == rk[0].lap, "z stage: the three components share the linear term");
  }
  a.pr = pr;
  a.cL = rk[0].cL;
  a.sNL = rk[0].sNL;
  a.lap = rk[0].lap;
  a.zmap = f.d_zmap;
  a.kx = p.d_kx; a.ky = p.d_ky; a.kz = p.d_kz;
  a.fx = p.d_fx; a.fy = p.d_fy; a.fz = p.d_fz;
  a.dir = p.d_dir;
  a.zc = p.d_z;
  a.dkz = p.Dkz;
  a.ny = p.ny; a.nxl = p.nxl; a.nph = f.nph; a.C = p.Cz; a.d = p.oz;
  a.has_mean = p.ista == 1 ? 1 : 0;
  a.dt = dt; a.rmp = rmp; a.Lz = p.Lz;
  double tmp = 1.0 / (double)o;
  if (o != p.ord) tmp = (double)(o + 1) * tmp;   // vboundary.f90:195-196
  a.tmp_noslip = tmp;
  a.inv_nz = 1.0 / (double)p.nz;
  const double sc = (double)p.nx * (double)p.ny;
  a.mx0 = sc * (zs ? zs[0] : 0.0); a.my0 = sc * (zs ? zs[1] : 0.0);
  a.mx1 = sc * (ze ? ze[0] : 0.0); a.my1 = sc * (ze ? ze[1] : 0.0);
  static const double r8[8][2] = {{1, 0}, {0.70710678118654752440, 0.70710678118654752440}, {0, 1}, {-0.70710678118654752440, 0.70710678118654752440},
                                  {-1, 0}, {-0.70710678118654752440, -0.70710678118654752440}, {0, -1}, {0.70710678118654752440, -0.70710678118654752440}};
  for (int k = 0; k < 8; ++k) {
    const int q = (int)(((long)k * (f.nph - 1)) % 8);
    a.phT[k] = cmake(r8[q][0], r8[q][1]);
  }
#define C_(N) run_zstage<N>(p, f, a)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}

}  // namespace sx
