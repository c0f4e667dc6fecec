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
  int ny, nxl, nph, C, d, has_mean;
  double dt, rmp, Lz, tmp_noslip, inv_nz;
  double mx0, my0, mx1, my1;    // nx*ny*v_wall (mean mode rows)
  cplx phT[8];                  // exp(+2 pi i k top / 8)
};

template <int N, int NP> struct ZstageGeo {
  static constexpr int T = N / 8, NT = NP * T, NW = NT / 32, PW = NP < 32 ? NP : 32;
  static constexpr size_t cplx_elems = (size_t)2 * N * NP      // two exchange buffers
                                       + (size_t)3 * 8 * NT     // two nonlinear-term slots + the v_z park
                                       + (size_t)2 * 2 * kMaxDF * NP   // boundary stashes of two continuations
                                       + (size_t)2 * NW * PW;   // wall-row partial sums
  static size_t smem_bytes() { return cplx_elems * sizeof(cplx) + (size_t)N * sizeof(ZMap); }
};

// stash / continue with an explicit pencil count (the tile kernels' helpers, sx_fused_zfwd.cu, on a given stash)
template <int N, int NP>
__device__ __forceinline__ void zs_stash(const cplx (&v)[8], int j, int p, cplx* bnd, int nph, int d) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e < d) bnd[e * NP + p] = v[k];
    if (e >= nph - d && e < nph) bnd[(d + e - (nph - d)) * NP + p] = v[k];
  }
}
template <int N, int NP>
__device__ __forceinline__ void zs_continue(cplx (&v)[8], int j, int p, const cplx* bnd, int nph, int C, int d,
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

template <int N, int NP, int MINB, int L2PF>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_zstage(ZstageArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  typedef ZstageGeo<N, NP> G;
  constexpr int T = G::T, NT = G::NT, NW = G::NW, PW = G::PW;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TwRegs<N> twr;
  twr.load(tw, j);
  const SIdxPencil si{p, NP};
  cplx* ex0 = smem;
  cplx* ex1 = ex0 + (size_t)N * NP;
  cplx* slotA = ex1 + (size_t)N * NP + threadIdx.x;   // slot[k * NT]
  cplx* slotB = slotA + (size_t)8 * NT;
  cplx* park = slotB + (size_t)8 * NT;                // v_z, thread-private
  cplx* bnd0 = smem + (size_t)2 * N * NP + (size_t)3 * 8 * NT;
  cplx* bnd1 = bnd0 + (size_t)2 * kMaxDF * NP;
  cplx* red = bnd1 + (size_t)2 * kMaxDF * NP;
  ZMap* zm = reinterpret_cast<ZMap*>(red + (size_t)2 * NW * PW);
  for (int z = threadIdx.x; z < a.nph; z += NT) zm[z] = a.zmap[z];
  __syncthreads();
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
  auto issue = [&](cplx* slot, int t, int c) {   // nonlinear term c of tile t into a slot (one commit group)
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = j + k * T;
      if (ky < a.ny && z < a.nph) {
        const ZMap m = zm[z];
        cp_async16(slot + k * NT, a.nl[c] + m.base + ((long long)kxl * m.nzl + m.zl) * a.ny + ky);
      } else {
        slot[k * NT] = cmake(0.0, 0.0);
      }
    }
    cp_async_commit();
  };
  // RK update of one component (k_zfwd_rk, BATCH == 2 association): v holds the transformed nonlinear term
  auto rk_update = [&](cplx (&v)[8], const cplx (&L)[8], int c, size_t base, double f1, double f2, double kh2) {
    if (a.couple[c] != nullptr) {
      cplx Q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) Q[k] = a.couple[c][base + j + k * T];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = caxpy(a.ccoef[c], Q[k], v[k]);
    }
    cplx F[8], B[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      F[k] = a.f[c][base + j + k * T];
      B[k] = a.v0[c][base + j + k * T];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      const double z = __ldg(&a.kz[e]), f3 = __ldg(&a.fz[e]);
      const double lm = a.lap ? -(kh2 + z * z) : 1.0;
      const cplx NL = cscale(cscale(cscale(v[k], f1), f2), f3);
      v[k] = cmake(a.cL * (lm * L[k].x) + a.sNL * NL.x, a.cL * (lm * L[k].y) + a.sNL * NL.y);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = cmake((v[k].x + F[k].x) * a.dt * a.rmp, (v[k].y + F[k].y) * a.dt * a.rmp);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = cmake(B[k].x + v[k].x, B[k].y + v[k].y);
  };

  int t = blockIdx.x;
  if (t < ntiles) {
    issue(slotA, t, 2);
    issue(slotB, t, 0);
  }
  for (; t < ntiles; t += gridDim.x) {
    const int ky = (t % tiles_y) * NP + p, kxl = t / tiles_y;
    const bool active = ky < a.ny;
    const int tn = t + gridDim.x;
    const size_t base = ((size_t)kxl * a.ny + (active ? ky : 0)) * N;
    const double x = __ldg(&a.kx[kxl]), y = __ldg(&a.ky[active ? ky : 0]);
    const double f1 = __ldg(&a.fx[kxl]), f2 = __ldg(&a.fy[active ? ky : 0]);
    const double kh2 = x * x + y * y;
    const bool mean = a.has_mean && kxl == 0 && ky == 0;
    const cplx pr0 = a.pr[base], prT = a.pr[base + top];
    if (L2PF && threadIdx.x < 9 && tn < ntiles) {
      // the spectral pencils of the next tile's RK updates (NP adjacent ky pencils are one contiguous range)
      const int ky0 = (tn % tiles_y) * NP, kxn = tn / tiles_y;
      const int np = a.ny - ky0 < NP ? a.ny - ky0 : NP;
      const size_t tb = ((size_t)kxn * a.ny + ky0) * N;
      const int c = threadIdx.x / 3, q = threadIdx.x % 3;
      const cplx* fld = q == 0 ? a.v[c] : (q == 1 ? a.v0[c] : a.f[c]);
      l2_prefetch(fld + tb, (unsigned)((size_t)np * N * sizeof(cplx)));
    }

    cplx v[8], w[8];
    // ---- v_z: continuation + transform + RK update, parked ------------------------------------------------
    {
      cplx L[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) L[k] = a.v[2][base + j + k * T];
      cp_async_wait_all();   // slot A (v_z term) and slot B (v_x term) of this tile
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = slotA[k * NT];
      issue(slotA, t, 1);    // the v_y term travels under this transform
      zs_stash<N, NP>(v, j, p, bnd0, a.nph, a.d);
      __syncthreads();
      zs_continue<N, NP>(v, j, p, bnd0, a.nph, a.C, a.d, a.dir);
      fft_regs<N, -1>(v, j, ex0, si, twr);
      rk_update(v, L, 2, base, f1, f2, kh2);
#pragma unroll
      for (int k = 0; k < 8; ++k) park[k * NT] = v[k];
    }
    // ---- v_x, v_y: continuation + transform + RK update ---------------------------------------------------
    {
      cplx L[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) L[k] = a.v[0][base + j + k * T];
      cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = slotB[k * NT];
        w[k] = slotA[k * NT];
      }
      if (tn < ntiles) {     // the next tile's v_z and v_x terms: a whole tile ahead
        issue(slotA, tn, 2);
        issue(slotB, tn, 0);
      }
      zs_stash<N, NP>(v, j, p, bnd0, a.nph, a.d);
      zs_stash<N, NP>(w, j, p, bnd1, a.nph, a.d);
      __syncthreads();
      zs_continue<N, NP>(v, j, p, bnd0, a.nph, a.C, a.d, a.dir);
      zs_continue<N, NP>(w, j, p, bnd1, a.nph, a.C, a.d, a.dir);
      fft_regs2<N, -1, -1>(v, w, j, ex0, ex1, si, twr);
      rk_update(v, L, 0, base, f1, f2, kh2);
#pragma unroll
      for (int k = 0; k < 8; ++k) L[k] = a.v[1][base + j + k * T];
      rk_update(w, L, 1, base, f1, f2, kh2);
    }
    // ---- no-slip rows of v_x, v_y in the mixed domain, back to Fourier (vboundary.f90:116-145) ------------
    fft_regs2<N, 1, 1>(v, w, j, ex0, ex1, si, twr);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      v[k] = cscale(v[k], a.inv_nz);
      w[k] = cscale(w[k], a.inv_nz);
      if (e == 0 || e == top) {
        const cplx P = e == 0 ? pr0 : prT;
        v[k] = cmake(-x * P.y * a.tmp_noslip, x * P.x * a.tmp_noslip);
        w[k] = cmake(-y * P.y * a.tmp_noslip, y * P.x * a.tmp_noslip);
        if (mean) {
          v[k] = cmake(e == 0 ? a.mx0 : a.mx1, 0.0);
          w[k] = cmake(e == 0 ? a.my0 : a.my1, 0.0);
        }
      }
    }
    zs_stash<N, NP>(v, j, p, bnd0, a.nph, a.d);
    zs_stash<N, NP>(w, j, p, bnd1, a.nph, a.d);
    __syncthreads();
    zs_continue<N, NP>(v, j, p, bnd0, a.nph, a.C, a.d, a.dir);
    zs_continue<N, NP>(w, j, p, bnd1, a.nph, a.C, a.d, a.dir);
    fft_regs2<N, -1, -1>(v, w, j, ex0, ex1, si, twr);
    // ---- particular solution and its gradient (boundary_mod.fpp:405-448, 249-259) -------------------------
    cplx A[8], B[8];   // v_x, v_y minus the gradient of the particular solution
    cplx s0 = cmake(0.0, 0.0), s1 = cmake(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      const double z = __ldg(&a.kz[e]);
      const double kk2 = x * x + y * y + z * z;
      cplx Cc = park[k * NT];
      const cplx s = cmake(x * v[k].x + y * w[k].x + z * Cc.x, x * v[k].y + y * w[k].y + z * Cc.y);
      cplx D = cmake(s.y / kk2, -s.x / kk2);
      if ((mean && e == 0) || !active) D = cmake(0.0, 0.0);
      A[k] = cmake(v[k].x + x * D.y, v[k].y - x * D.x);
      B[k] = cmake(w[k].x + y * D.y, w[k].y - y * D.x);
      Cc = cmake(Cc.x + z * D.y, Cc.y - z * D.x);
      park[k * NT] = Cc;
      v[k] = D;
      // wall values of v_z (boundary_mod.fpp:275-338): rows 0 and top of IFFT_z(v_z)/nz
      s0 = cadd(s0, Cc);
      s1 = cadd(s1, cmul(Cc, a.phT[k]));
    }
    s1 = cmul(s1, phj);
#pragma unroll
    for (int o = NP; o < 32; o <<= 1) {
      s0.x += __shfl_xor_sync(0xffffffffu, s0.x, o);
      s0.y += __shfl_xor_sync(0xffffffffu, s0.y, o);
      s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o);
      s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
    }
    if (lane < PW) {
      red[(warp * PW + lane) * 2] = s0;
      red[(warp * PW + lane) * 2 + 1] = s1;
    }
    __syncthreads();   // partial sums visible
    cplx bc1 = cmake(0.0, 0.0), bc2 = cmake(0.0, 0.0);
    {
      const int pw = p % PW;
#pragma unroll
      for (int q = 0; q < NW; ++q) {
        // with NP >= 32 a warp holds ONE element row of 32 pencils: only the warps of this pencil's group count
        if (NP > 32 && (q * 32) % NP / 32 != p / 32) continue;
        bc1 = cadd(bc1, red[(q * PW + pw) * 2]);
        bc2 = cadd(bc2, red[(q * PW + pw) * 2 + 1]);
      }
    }
    bc1 = cscale(bc1, a.inv_nz);
    bc2 = cscale(bc2, a.inv_nz);
    // ---- laplace_z, Neumann-Neumann (boundary_mod.fpp:531-560, 635-675) -----------------------------------
    const double kh = sqrt(x * x + y * y);
    cplx c1, c2;
    if (mean) {
      c1 = bc1;
      c2 = cmake(0.0, 0.0);
    } else {
      const double e1 = exp(-kh * a.Lz), tt = 1.0 / (kh * (1.0 - e1 * e1));
      c1 = cmake((bc2.x - bc1.x * e1) * tt, (bc2.y - bc1.y * e1) * tt);
      c2 = cmake((-bc1.x + bc2.x * e1) * tt, (-bc1.y + bc2.y * e1) * tt);
    }
    // the real sequence whose continued transform carries the harmonic correction: e^{-kh z} (mean pencil: the
    // constant Re(c1) = phi'), physical rows; geometric in the thread's stride
    double st = 0.0, em = 0.0;
    if (!mean) {
      st = exp(-kh * dzT);
      em = exp(-kh * zj);
    }
    {
      double q = em;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        w[k] = cmake(mean ? c1.x : q, 0.0);
        q *= st;
      }
    }
    zs_stash<N, NP>(w, j, p, bnd1, a.nph, a.d);
    __syncthreads();
    zs_continue<N, NP>(w, j, p, bnd1, a.nph, a.C, a.d, a.dir);
    // p' = IFFT_z(d)/nz + phi (boundary_mod.fpp:371-380) next to the transform of the exponential
    fft_regs2<N, 1, -1>(v, w, j, ex0, ex1, si, twr);
    {
      double ep[8];
      if (!mean) {
        ep[7] = exp(kh * (z7 - a.Lz));
#pragma unroll
        for (int k = 6; k >= 0; --k) ep[k] = ep[k + 1] * st;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        cplx ph;
        if (mean) {
          const double z = __ldg(&a.zc[e]);
          ph = cmake(c1.x * z + c2.x, 0.0);
        } else {
          ph = cmake(c1.x * ep[k] + c2.x * em, c1.y * ep[k] + c2.y * em);
          em *= st;
        }
        if (active) a.pr[base + e] = cmake(v[k].x * a.inv_nz + ph.x, v[k].y * a.inv_nz + ph.y);
      }
    }
    // ---- subtract the harmonic correction (boundary_mod.fpp:385-399) --------------------------------------
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = j + k * T;
        const cplx Cc = park[k * NT];
        cplx h, hz;   // phi^, phi'^
        if (mean) {
          h = cmake(0.0, 0.0);
          hz = w[k];
        } else {
          const cplx Em = w[k];
          const cplx Ep = cconj(cmul(cmul(a.phT[k], phj), Em));
          const cplx t1 = cmul(c1, Ep), t2 = cmul(c2, Em);
          h = cadd(t1, t2);
          hz = cscale(csub(t1, t2), kh);
        }
        a.v[0][base + e] = cmake(A[k].x + x * h.y, A[k].y - x * h.x);
        a.v[1][base + e] = cmake(B[k].x + y * h.y, B[k].y - y * h.x);
        a.v[2][base + e] = cmake(Cc.x - hz.x, Cc.y - hz.y);
      }
    }
  }
}

template <int N, int NP, int MINB, int L2PF>
static int run_zstage_v(Plan& p, Fused& f, const ZstageArgs& a) {
  typedef ZstageGeo<N, NP> G;
  const cplx* tw = p.tw_z;
  auto kfn = k_zstage<N, NP, MINB, L2PF>;
  const size_t smem = G::smem_bytes();
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
