// Fused substep: v_imposebc_and_project, one (ky,kx) pencil at a time.
#include "sx_fused.h"

namespace sx {

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
                                           int d, const double* __restrict__ dir, const TwRegs<N>& twr) {
  __syncthreads();  // bnd may still be read by a previous continuation
  stash_boundary_elem<N>(v, j, bnd, nph, d);
  __syncthreads();
  fc_continue_elem<N>(v, j, bnd, nph, C, d, dir);
  fft_regs<N, -1>(v, j, smem, si, twr);
}

template <int N, int NPB, bool PF, int MINB>
__global__ void __launch_bounds__(NPB*(N / 8), MINB) k_project(ProjArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NPB * T;
  constexpr int XS = sidx_elem_stride<N>();
  const int pl = threadIdx.x / T, j = threadIdx.x % T;
  TwRegs<N> twr;
  twr.load(tw, j);
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
      fft_regs<N, 1>(v, j, smem, si, twr);
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
      fc_fft_fwd<N>(v, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, twr);
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
    fft_regs<N, 1>(v, j, smem, si, twr);
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
    fft_regs<N, 1>(dd, j, smem, si, twr);
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
    fc_fft_fwd<N>(v, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, twr);
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a.vz[base + j + k * T] = cmake(cz[k].x - v[k].x, cz[k].y - v[k].y);
    }
    fc_fft_fwd<N>(dd, j, smem, si, bnd, a.nph, a.C, a.d, a.dir, twr);
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

// pencil arguments plus the stride factors of the top-wall row of the backward transform
struct ProjBulkArgs {
  ProjArgs a;
  cplx phT[8];   // exp(+2 pi i k top / 8), k = 0..7: stride factors of the top-wall row of the backward transform
};

// ------------------------------------------------------------------------------------------
// project, paired version: one pencil per CTA (N/8 threads), persistent.  The three spectral pencils arrive as 8 KB
// cp.async.bulk copies (TMA) into three shared-memory slots, which then park the intermediate fields (every thread only
// touches its own elements) and are refilled for the next pencil while the results are formed; the wall values of v_z
// are two dot products with a block reduction (boundary_mod.fpp:275-338 only uses rows 1 and nz-Cz); e^{-kh z} and
// e^{kh (z - Lz)} on the uniform grid are geometric sequences in the thread's stride.  Against the round-1 kernel
// (profiles/r2_zstage_experiment.md):
//   * SIX transforms instead of seven, in THREE rounds of two (fft_regs2: both pencils of a round share every barrier
//     and the twiddle powers): (IFFT v_x, IFFT v_y), (FFT v_x, FFT v_y), (IFFT d, FFT e).  The harmonic correction
//     phi = c1 e^{kh (z - Lz)} + c2 e^{-kh z} and its derivative are linear in the two exponentials, and so are the
//     continuation and the transform: phi^ = c1 E+^ + c2 E-^, phi'^ = kh (c1 E+^ - c2 E-^) with E-^ = FFT(cont(e^{-kh z})),
//     ONE real-input transform instead of the two complex ones of boundary_mod.fpp:385-399.  E+ is the mirror image of
//     E- on the grid (z_k = k dz, Lz = z_top) and the FC-Gram continuation commutes with that reflection (fftp.fpp:760-770
//     is symmetric under ii -> C-ii+1 with the two boundary stencils swapped), so E+^(k) = conj(e^{2 pi i k top / N} E-^(k)).
//     The mean pencil (kx = ky = 0) has phi = Re(c1) z + Re(c2): its phi^ multiplies kx = ky = 0 and drops out, and
//     phi'^ = Re(c1) FFT(cont(1)) takes the place of E-^;
//   * ONE copy of the forward transform in a rolled loop over the rounds (a backward transform is the forward one
//     between two conjugations): the unrolled round-1 kernel (k_project_bulk, seven transform instances) was 152 KB of
//     instructions with 9 % of its stall samples on instruction fetches, and the merged z-stage experiment lost a
//     factor of three to them;
//   * one reciprocal instead of two divisions per element in the Poisson step.
// ------------------------------------------------------------------------------------------
template <int N> struct ProjPairGeo {
  static constexpr int T = N / 8, XS = sidx_elem_stride<N>(), NW = (T + 31) / 32;
  static size_t smem_bytes(int C) {
    return ((size_t)2 * XS + (size_t)3 * N + (size_t)4 * kMaxDF + (size_t)2 * C + (size_t)2 * NW) * sizeof(cplx) + 3 * 8;
  }
};

template <int N, int MINB>
__global__ void __launch_bounds__(N / 8, MINB) k_project_pair(ProjBulkArgs pa, double dkz, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  const ProjArgs& a = pa.a;
  typedef ProjPairGeo<N> G;
  constexpr int T = G::T, XS = G::XS, NW = G::NW;
  constexpr unsigned PBYTES = N * sizeof(cplx);
  const int j = threadIdx.x;
  const bool lead = j == 0;
  TwRegs<N> twr;
  twr.load(tw, j);
  const SIdxElem si{0};
  cplx* ex0 = smem;
  cplx* ex1 = ex0 + XS;
  cplx* sx = ex1 + XS;
  cplx* sy = sx + N;
  cplx* sz = sy + N;
  cplx* bnd = sz + N;                 // [a * 2d + q]: boundary rows of the two arrays of a round
  cplx* cv = bnd + 4 * kMaxDF;        // [a * C + ii]: their continuation rows
  cplx* red = cv + 2 * a.C;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(red + 2 * NW);
  const int top = a.nph - 1;
  // phase of the top-wall row at this thread's first element
  cplx phj;
  {
    double sn, cs;
    sincospi(2.0 * (double)(((long)j * top) % N) / (double)N, &sn, &cs);
    phj = cmake(cs, sn);
  }
  const double zj = __ldg(&a.zc[j]), z7 = __ldg(&a.zc[j + 7 * T]), dzT = __ldg(&a.zc[T]) - __ldg(&a.zc[0]);
  if (lead) {
    for (int b = 0; b < 3; ++b) mbar_init(bar + b, 1);
    mbar_init_fence();
  }
  __syncthreads();
  int g = blockIdx.x;   // pencil index: ny * nxl < 2^31
  const int npencils = (int)a.npencils;
  if (lead && g < npencils) {
    bulk_load(sx, a.vx + (size_t)g * N, PBYTES, bar);
    bulk_load(sy, a.vy + (size_t)g * N, PBYTES, bar + 1);
    bulk_load(sz, a.vz + (size_t)g * N, PBYTES, bar + 2);
  }
  unsigned phase = 0;
  for (; g < npencils; g += gridDim.x) {
    const int gn = g + gridDim.x;
    const int kx_i = g / a.ny, ky_i = g - kx_i * a.ny;
    const size_t base = (size_t)g * N;
    const double x = __ldg(&a.kx[kx_i]), y = __ldg(&a.ky[ky_i]);
    const double kh = sqrt(x * x + y * y);
    const bool mean = a.has_mean && g == 0;
    const cplx pr0 = a.pr[base], prT = a.pr[base + top];
    cplx c1 = cmake(0.0, 0.0), c2 = cmake(0.0, 0.0);
    double st = 0.0, em = 0.0;
    cplx v[8], w[8];
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
      // ---- inputs of the round ----
      if (r == 0) {
        mbar_wait(bar, phase);
        mbar_wait(bar + 1, phase);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = cconj(sx[j + k * T]);
          w[k] = cconj(sy[j + k * T]);
        }
      } else if (r == 2) {
        // d for its backward transform, and the real sequence whose continued transform carries the harmonic correction:
        // e^{-kh z} (mean pencil: the constant Re(c1) = phi') on the physical rows, geometric in the thread's stride
        double q = em;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k].y = -v[k].y;
          w[k] = cmake(mean ? c1.x : q, 0.0);
          q *= st;
        }
      }
      if (r > 0) {
        // FC-Gram continuation (fftp.fpp:757-772) of w, and of v in round 1: boundary rows to shared memory, the first
        // C threads form one row each for both arrays, every thread picks up the rows it owns
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          int q = -1;
          if (e < a.d) q = e;
          if (e >= a.nph - a.d && e < a.nph) q = a.d + e - (a.nph - a.d);
          if (q >= 0) {
            bnd[q] = v[k];
            bnd[2 * a.d + q] = w[k];
          }
        }
        __syncthreads();
#pragma unroll 1
        for (int ii = j; ii < a.C; ii += T) {
          double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
          for (int jj = 0; jj < a.d; ++jj) {
            const double w1 = __ldg(&a.dir[ii * a.d + jj]);
            const double w2 = __ldg(&a.dir[(a.C - 1 - ii) * a.d + jj]);
            const cplx f1 = bnd[a.d + jj], f2 = bnd[a.d - 1 - jj];
            const cplx g1 = bnd[3 * a.d + jj], g2 = bnd[3 * a.d - 1 - jj];
            ax = fma(w2, f2.x, fma(w1, f1.x, ax));
            ay = fma(w2, f2.y, fma(w1, f1.y, ay));
            bx = fma(w2, g2.x, fma(w1, g1.x, bx));
            by = fma(w2, g2.y, fma(w1, g1.y, by));
          }
          cv[ii] = cmake(ax, ay);
          cv[a.C + ii] = cmake(bx, by);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          if (e >= a.nph) {
            if (r == 1) v[k] = cv[e - a.nph];
            w[k] = cv[a.C + e - a.nph];
          }
        }
      }
      fft_regs2<N, -1, -1>(v, w, j, ex0, ex1, si, twr);
      // ---- results of the round ----
      if (r == 0) {
        // no-slip rows of vx, vy in the mixed domain (vboundary.f90:116-145); v, w hold the conjugates
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          v[k] = cmake(v[k].x * a.inv_nz, -v[k].y * a.inv_nz);
          w[k] = cmake(w[k].x * a.inv_nz, -w[k].y * a.inv_nz);
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
      } else if (r == 1) {
        // particular solution and its gradient (boundary_mod.fpp:405-448, 249-259)
        mbar_wait(bar + 2, phase);
        phase ^= 1;
        cplx s0 = cmake(0.0, 0.0), s1 = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          const double z = (double)(e < N / 2 ? e : e - N) * dkz;   // the product the host table holds
          const double kk2 = x * x + y * y + z * z;
          cplx Cc = sz[e];
          const cplx s = cmake(x * v[k].x + y * w[k].x + z * Cc.x, x * v[k].y + y * w[k].y + z * Cc.y);
          const double ik2 = 1.0 / kk2;
          cplx D = cmake(s.y * ik2, -s.x * ik2);
          if (mean && e == 0) D = cmake(0.0, 0.0);
          sx[e] = cmake(v[k].x + x * D.y, v[k].y - x * D.x);
          sy[e] = cmake(w[k].x + y * D.y, w[k].y - y * D.x);
          Cc = cmake(Cc.x + z * D.y, Cc.y - z * D.x);
          sz[e] = Cc;   // parked in the slots (own elements only)
          v[k] = D;
          // wall values of v_z (boundary_mod.fpp:275-338): rows 0 and top of IFFT_z(v_z)/nz
          s0 = cadd(s0, Cc);
          s1 = cadd(s1, cmul(Cc, pa.phT[k]));
        }
        s1 = cmul(s1, phj);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0.x += __shfl_xor_sync(0xffffffffu, s0.x, o);
          s0.y += __shfl_xor_sync(0xffffffffu, s0.y, o);
          s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o);
          s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
        }
        if ((j & 31) == 0) {
          red[2 * (j >> 5)] = s0;
          red[2 * (j >> 5) + 1] = s1;
        }
        __syncthreads();   // partial sums visible
        cplx bc1 = red[0], bc2 = red[1];
#pragma unroll
        for (int q = 1; q < NW; ++q) {
          bc1 = cadd(bc1, red[2 * q]);
          bc2 = cadd(bc2, red[2 * q + 1]);
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
      } else {
        // p' = IFFT_z(d)/nz + phi (boundary_mod.fpp:371-380); v holds the conjugate of the backward transform
        {
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
            a.pr[base + e] = cmake(v[k].x * a.inv_nz + ph.x, -v[k].y * a.inv_nz + ph.y);
          }
        }
        // the harmonic correction subtracted (boundary_mod.fpp:385-399).  The parked fields come back to registers
        // first, so that the slots can be refilled for the next pencil while the results are formed and stored.
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = sz[j + k * T];
        cplx A[8], B[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          A[k] = sx[j + k * T];
          B[k] = sy[j + k * T];
        }
        __syncthreads();   // every thread is done with the three slots
        if (lead && gn < npencils) {
          bulk_load(sx, a.vx + (size_t)gn * N, PBYTES, bar);
          bulk_load(sy, a.vy + (size_t)gn * N, PBYTES, bar + 1);
          bulk_load(sz, a.vz + (size_t)gn * N, PBYTES, bar + 2);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          cplx h, hz;   // phi^, phi'^
          if (mean) {
            h = cmake(0.0, 0.0);
            hz = w[k];
          } else {
            const cplx Em = w[k];
            const cplx Ep = cconj(cmul(cmul(pa.phT[k], phj), Em));
            const cplx t1 = cmul(c1, Ep), t2 = cmul(c2, Em);
            h = cadd(t1, t2);
            hz = cscale(csub(t1, t2), kh);
          }
          a.vx[base + e] = cmake(A[k].x + x * h.y, A[k].y - x * h.x);
          a.vy[base + e] = cmake(B[k].x + y * h.y, B[k].y - y * h.x);
          a.vz[base + e] = cmake(v[k].x - hz.x, v[k].y - hz.y);
        }
      }
    }
  }
}

template <int N, int MINB> static int run_project_pair(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o,
                                                       const double* zs, const double* ze) {
  constexpr int T = N / 8;
  double tmp = 1.0 / (double)o;
  if (o != p.ord) tmp = (double)(o + 1) * tmp;   // vboundary.f90:195-196
  const double sc = (double)p.nx * (double)p.ny;
  ProjBulkArgs pa;
  pa.a = ProjArgs{vx, vy, vz, pr, p.d_kx, p.d_ky, p.d_kz, p.d_z, p.d_dir, (long)p.ny * p.nxl,
                  p.ny, f.nph, p.Cz, p.oz, p.ista == 1 ? 1 : 0, p.Lz, tmp, 1.0 / (double)p.nz,
                  sc * (zs ? zs[0] : 0.0), sc * (zs ? zs[1] : 0.0), sc * (ze ? ze[0] : 0.0), sc * (ze ? ze[1] : 0.0)};
  static const double r8[8][2] = {{1, 0}, {0.70710678118654752440, 0.70710678118654752440}, {0, 1}, {-0.70710678118654752440, 0.70710678118654752440},
                                  {-1, 0}, {-0.70710678118654752440, -0.70710678118654752440}, {0, -1}, {0.70710678118654752440, -0.70710678118654752440}};
  for (int k = 0; k < 8; ++k) {
    const int q = (int)(((long)k * (f.nph - 1)) % 8);
    pa.phT[k] = cmake(r8[q][0], r8[q][1]);
  }
  const cplx* tw = p.tw_z;
  const double dkz = p.Dkz;
  auto kfn = k_project_pair<N, MINB>;
  const size_t smem = ProjPairGeo<N>::smem_bytes(p.Cz);
  int grid;
  if (persistent_grid(p, kfn, T, smem, (int)pa.a.npencils, &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_PROJECT, kfn, dim3(grid), T, smem, pa, dkz, tw);
  return 0;
}

// ------------------------------------------------------------------------------------------
// a_imposebc_and_project for conducting walls at both ends (bboundary.f90:100-189 with int_conducting_z :192-236,
// sol_project(...,0,0,0) boundary_mod.fpp:197-402 and conducting_z :239-290 -> neumann_reconstruct
// fcgram_mod.f90:456-499), one (ky,kx) pencil per CTA like k_project_pair: TWELVE transforms in six paired rounds
// instead of the thirteen stand-alone z transforms and eight elementwise passes of the per-operator composition
// (10.8 of the 40.6 ms of an MHD 512^3 substep, profiles/r2l_mhd512_launches.md).
//   round 0: IFFT (a_x, a_y)                     -> wall rows <- 0
//   round 1: FFT  (a_x, a_y), continued          -> d = -i k.a / k^2, a -= i k d (parked)
//   round 2: IFFT d, FFT cont(e^{-kh z})         -> wall values of d, Dirichlet laplace_z, ph = d + phi, a -= grad(phi)
//   round 3: IFFT (a_x, a_y)                     -> wall rows <- second-order Neumann reconstruction
//   round 4: FFT a_x (continued), IFFT a_z       -> a_x final; wall rows of a_z <- first-order reconstruction
//   round 5: FFT (a_y, a_z), continued           -> final
// The harmonic correction uses the same single real-input transform as k_project_pair.
// ------------------------------------------------------------------------------------------
struct AprojArgs {
  cplx *ax, *ay, *az, *ph;
  const double *kx, *ky, *zc, *dir;
  long npencils;
  int ny, nph, C, d, has_mean;
  double Lz, inv_nz, dkz;
  double neu1[kMaxDF], neu2[kMaxDF];   // neumann_reconstruct weights of order 1 and 2 (fcgram_mod.f90:355-361)
  cplx phT[8];
};

template <int N, int MINB>
__global__ void __launch_bounds__(N / 8, MINB) k_aproject_pair(AprojArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  typedef ProjPairGeo<N> G;
  constexpr int T = G::T, XS = G::XS, NW = G::NW;
  constexpr unsigned PBYTES = N * sizeof(cplx);
  const int j = threadIdx.x;
  const bool lead = j == 0;
  TwRegs<N> twr;
  twr.load(tw, j);
  const SIdxElem si{0};
  cplx* ex0 = smem;
  cplx* ex1 = ex0 + XS;
  cplx* sx = ex1 + XS;
  cplx* sy = sx + N;
  cplx* sz = sy + N;
  cplx* bnd = sz + N;                 // [arr * 2d + q]
  cplx* cv = bnd + 4 * kMaxDF;        // [arr * C + ii]
  cplx* red = cv + 2 * a.C;           // wall values of d
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(red + 2 * NW);
  const int top = a.nph - 1;
  cplx phj;
  {
    double sn, cs;
    sincospi(2.0 * (double)(((long)j * top) % N) / (double)N, &sn, &cs);
    phj = cmake(cs, sn);
  }
  const double zj = __ldg(&a.zc[j]), z7 = __ldg(&a.zc[j + 7 * T]), dzT = __ldg(&a.zc[T]) - __ldg(&a.zc[0]);
  if (lead) {
    for (int b = 0; b < 3; ++b) mbar_init(bar + b, 1);
    mbar_init_fence();
  }
  __syncthreads();
  int g = blockIdx.x;
  const int npencils = (int)a.npencils;
  if (lead && g < npencils) {
    bulk_load(sx, a.ax + (size_t)g * N, PBYTES, bar);
    bulk_load(sy, a.ay + (size_t)g * N, PBYTES, bar + 1);
    bulk_load(sz, a.az + (size_t)g * N, PBYTES, bar + 2);
  }
  unsigned phase = 0;
  for (; g < npencils; g += gridDim.x) {
    const int gn = g + gridDim.x;
    const int kx_i = g / a.ny, ky_i = g - kx_i * a.ny;
    const size_t base = (size_t)g * N;
    const double x = __ldg(&a.kx[kx_i]), y = __ldg(&a.ky[ky_i]);
    const double kh = sqrt(x * x + y * y);
    const bool mean = a.has_mean && g == 0;
    cplx v[8], w[8];
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
      // which arrays are continued (c) / reconstructed at the walls with weights of order (o) before this round's transform
      const bool cvv = r == 1 || r == 4 || r == 5, cvw = r == 1 || r == 2 || r == 5;
      const int ov = 0, ow = r == 5 ? 1 : 0;
      // ---- inputs of the round ----
      if (r == 0) {
        mbar_wait(bar, phase);
        mbar_wait(bar + 1, phase);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = cconj(sx[j + k * T]);
          w[k] = cconj(sy[j + k * T]);
        }
      } else if (r == 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          sy[j + k * T] = w[k];               // a_y (mixed domain, walls reconstructed) waits for round 5
          w[k] = cconj(sz[j + k * T]);        // a_z for its backward transform
        }
      } else if (r == 5) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = sy[j + k * T];
      }
      if (cvv || cvw) {
        // boundary rows to shared memory; optional Neumann reconstruction of the wall rows (their prescribed datum is
        // zero: neu(d) * 0 + sum_k neu(k) f(.)); then the continuation rows, one per thread for both arrays
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          int q = -1;
          if (e < a.d) q = e;
          if (e >= a.nph - a.d && e < a.nph) q = a.d + e - (a.nph - a.d);
          if (q >= 0) {
            bnd[q] = v[k];
            bnd[2 * a.d + q] = w[k];
          }
        }
        __syncthreads();
        if (ov || ow) {
          if (j < 4) {   // thread 2 * arr + wall
            const int arr = j >> 1, upper = j & 1, o = arr == 0 ? ov : ow;
            if (o) {
              const double* wt = o == 1 ? a.neu1 : a.neu2;
              cplx* b = bnd + arr * 2 * a.d;
              double sx_ = 0.0, sy_ = 0.0;
              for (int k = 1; k < a.d; ++k) {
                const cplx f = upper ? b[a.d + k - 1] : b[a.d - k];
                sx_ += wt[k - 1] * f.x;
                sy_ += wt[k - 1] * f.y;
              }
              b[upper ? 2 * a.d - 1 : 0] = cmake(sx_, sy_);
            }
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            if (e == 0 || e == top) {
              if (ov) v[k] = bnd[e == 0 ? 0 : 2 * a.d - 1];
              if (ow) w[k] = bnd[2 * a.d + (e == 0 ? 0 : 2 * a.d - 1)];
            }
          }
        }
#pragma unroll 1
        for (int ii = j; ii < a.C; ii += T) {
          double ax_ = 0.0, ay_ = 0.0, bx_ = 0.0, by_ = 0.0;
          for (int jj = 0; jj < a.d; ++jj) {
            const double w1 = __ldg(&a.dir[ii * a.d + jj]);
            const double w2 = __ldg(&a.dir[(a.C - 1 - ii) * a.d + jj]);
            const cplx f1 = bnd[a.d + jj], f2 = bnd[a.d - 1 - jj];
            const cplx g1 = bnd[3 * a.d + jj], g2 = bnd[3 * a.d - 1 - jj];
            ax_ = fma(w2, f2.x, fma(w1, f1.x, ax_));
            ay_ = fma(w2, f2.y, fma(w1, f1.y, ay_));
            bx_ = fma(w2, g2.x, fma(w1, g1.x, bx_));
            by_ = fma(w2, g2.y, fma(w1, g1.y, by_));
          }
          cv[ii] = cmake(ax_, ay_);
          cv[a.C + ii] = cmake(bx_, by_);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          if (e >= a.nph) {
            if (cvv) v[k] = cv[e - a.nph];
            if (cvw) w[k] = cv[a.C + e - a.nph];
          }
        }
      }
      fft_regs2<N, -1, -1>(v, w, j, ex0, ex1, si, twr);
      // ---- results of the round ----
      if (r == 0 || r == 3) {
        // mixed domain (goto_domain_w_boundaries): scale, wall rows <- 0 (int_conducting_z / conducting_z); in round 3
        // the second-order reconstruction of both wall rows follows from the neighbouring rows
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          v[k] = cmake(v[k].x * a.inv_nz, -v[k].y * a.inv_nz);
          w[k] = cmake(w[k].x * a.inv_nz, -w[k].y * a.inv_nz);
          if (e == 0 || e == top) v[k] = w[k] = cmake(0.0, 0.0);
        }
        if (r == 3) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            int q = -1;
            if (e < a.d) q = e;
            if (e >= a.nph - a.d && e < a.nph) q = a.d + e - (a.nph - a.d);
            if (q >= 0) {
              bnd[q] = v[k];
              bnd[2 * a.d + q] = w[k];
            }
          }
          __syncthreads();
          if (j < 4) {
            const int arr = j >> 1, upper = j & 1;
            cplx* b = bnd + arr * 2 * a.d;
            double sx_ = 0.0, sy_ = 0.0;
            for (int k = 1; k < a.d; ++k) {
              const cplx f = upper ? b[a.d + k - 1] : b[a.d - k];
              sx_ += a.neu2[k - 1] * f.x;
              sy_ += a.neu2[k - 1] * f.y;
            }
            cv[j] = cmake(sx_, sy_);
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = j + k * T;
            if (e == 0 || e == top) {
              v[k] = cv[e == 0 ? 0 : 1];
              w[k] = cv[e == 0 ? 2 : 3];
            }
          }
        }
      } else if (r == 1) {
        // particular solution and its gradient (boundary_mod.fpp:405-448, 249-259); a_z(1,1,1) = 0 first (bboundary.f90:147-149)
        mbar_wait(bar + 2, phase);
        phase ^= 1;
        double q = 0.0, st = 0.0;
        if (!mean) {
          st = exp(-kh * dzT);
          q = exp(-kh * zj);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          const double z = (double)(e < N / 2 ? e : e - N) * a.dkz;
          const double kk2 = x * x + y * y + z * z;
          cplx Cc = sz[e];
          if (mean && e == 0) Cc = cmake(0.0, 0.0);
          const cplx s = cmake(x * v[k].x + y * w[k].x + z * Cc.x, x * v[k].y + y * w[k].y + z * Cc.y);
          const double ik2 = 1.0 / kk2;
          cplx D = cmake(s.y * ik2, -s.x * ik2);
          if (mean && e == 0) D = cmake(0.0, 0.0);
          sx[e] = cmake(v[k].x + x * D.y, v[k].y - x * D.x);
          sy[e] = cmake(w[k].x + y * D.y, w[k].y - y * D.x);
          sz[e] = cmake(Cc.x + z * D.y, Cc.y - z * D.x);
          v[k] = cmake(D.x, -D.y);              // conjugate: backward transform in round 2
          w[k] = cmake(mean ? 1.0 : q, 0.0);    // e^{-kh z} (mean pencil: the constant 1), physical rows
          q *= st;
        }
      } else if (r == 2) {
        // C1 = IFFT_z(d)/nz (v holds its conjugate); wall values, Dirichlet laplace_z (boundary_mod.fpp:499-528, 635-675)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          v[k] = cmake(v[k].x * a.inv_nz, -v[k].y * a.inv_nz);
          if (e == 0) red[0] = v[k];
          if (e == top) red[1] = v[k];
        }
        __syncthreads();
        const cplx bc1 = cmake(-red[0].x, -red[0].y), bc2 = cmake(-red[1].x, -red[1].y);
        cplx c1, c2;
        double st = 0.0, em = 0.0;
        if (mean) {
          c1 = cmake((bc2.x - bc1.x) / a.Lz, (bc2.y - bc1.y) / a.Lz);
          c2 = bc1;
        } else {
          const double e1 = exp(-kh * a.Lz), tt = 1.0 / (1.0 - e1 * e1);
          c1 = cmake((bc2.x - bc1.x * e1) * tt, (bc2.y - bc1.y * e1) * tt);
          c2 = cmake((bc1.x - bc2.x * e1) * tt, (bc1.y - bc2.y * e1) * tt);
          st = exp(-kh * dzT);
          em = exp(-kh * zj);
        }
        {
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
            cplx phi;
            if (mean) {
              const double z = __ldg(&a.zc[e]);
              phi = cmake(c1.x * z + c2.x, 0.0);
            } else {
              phi = cmake(c1.x * ep[k] + c2.x * q, c1.y * ep[k] + c2.y * q);
              q *= st;
            }
            a.ph[base + e] = cadd(v[k], phi);   // d = C1 + C2 (boundary_mod.fpp:371-380)
          }
        }
        // a -= grad(phi) in Fourier space, then straight on to the backward transforms of round 3
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          cplx h, hz;
          if (mean) {
            h = cmake(0.0, 0.0);
            hz = cscale(w[k], c1.x);
          } else {
            const cplx Em = w[k];
            const cplx Ep = cconj(cmul(cmul(a.phT[k], phj), Em));
            const cplx t1 = cmul(c1, Ep), t2 = cmul(c2, Em);
            h = cadd(t1, t2);
            hz = cscale(csub(t1, t2), kh);
          }
          const cplx A = sx[e], B = sy[e], Cc = sz[e];
          v[k] = cmake(A.x + x * h.y, -(A.y - x * h.x));     // conjugates: backward transforms next
          w[k] = cmake(B.x + y * h.y, -(B.y - y * h.x));
          sz[e] = cmake(Cc.x - hz.x, Cc.y - hz.y);
        }
      } else if (r == 4) {
        // a_x is final; a_z reaches the mixed domain: scale, wall rows <- 0 (reconstructed with the first-order weights
        // at the start of round 5)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          a.ax[base + e] = v[k];
          w[k] = cmake(w[k].x * a.inv_nz, -w[k].y * a.inv_nz);
          if (e == 0 || e == top) w[k] = cmake(0.0, 0.0);
        }
      } else {
        __syncthreads();   // every thread is done with the three slots: refill them for the next pencil
        if (lead && gn < npencils) {
          bulk_load(sx, a.ax + (size_t)gn * N, PBYTES, bar);
          bulk_load(sy, a.ay + (size_t)gn * N, PBYTES, bar + 1);
          bulk_load(sz, a.az + (size_t)gn * N, PBYTES, bar + 2);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a.ay[base + j + k * T] = v[k];
          a.az[base + j + k * T] = w[k];
        }
      }
    }
  }
}

// returns -1 when the fused kernel does not apply (other wall kinds, short pencils): the caller composes the operators
int fused_aproject(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph) {
  if (p.b_bczsta != 0 || p.b_bczend != 0 || p.Cz <= 0 || p.nz < 256 || p.nz > 2048 || p.knob_pj == 9) return -1;
  if (load_neumann(p)) return 1;
  AprojArgs a;
  a.ax = ax; a.ay = ay; a.az = az; a.ph = ph;
  a.kx = p.d_kx; a.ky = p.d_ky; a.zc = p.d_z; a.dir = p.d_dir;
  a.npencils = (long)p.ny * p.nxl;
  a.ny = p.ny; a.nph = p.nphys(); a.C = p.Cz; a.d = p.oz; a.has_mean = p.ista == 1 ? 1 : 0;
  a.Lz = p.Lz; a.inv_nz = 1.0 / (double)p.nz; a.dkz = p.Dkz;
  for (int k = 0; k < kMaxDF; ++k) {
    a.neu1[k] = k < p.oz ? p.h_neu[k] : 0.0;
    a.neu2[k] = k < p.oz ? p.h_neu2[k] : 0.0;
  }
  static const double r8[8][2] = {{1, 0}, {0.70710678118654752440, 0.70710678118654752440}, {0, 1}, {-0.70710678118654752440, 0.70710678118654752440},
                                  {-1, 0}, {-0.70710678118654752440, -0.70710678118654752440}, {0, -1}, {0.70710678118654752440, -0.70710678118654752440}};
  for (int k = 0; k < 8; ++k) {
    const int q = (int)(((long)k * (p.nphys() - 1)) % 8);
    a.phT[k] = cmake(r8[q][0], r8[q][1]);
  }
  const cplx* tw = p.tw_z;
#define SX_APROJ_(N, MINB)                                                                          \
  case N: {                                                                                         \
    auto kfn = k_aproject_pair<N, MINB>;                                                            \
    const size_t smem = ProjPairGeo<N>::smem_bytes(p.Cz);                                           \
    int grid;                                                                                       \
    if (persistent_grid(p, kfn, N / 8, smem, (int)a.npencils, &grid)) return 1;                     \
    SX_FUSED_LAUNCH(p, ST_PROJECT, kfn, dim3(grid), N / 8, smem, a, tw);                            \
    return 0;                                                                                       \
  }
  switch (p.nz) {
    SX_APROJ_(256, 4)
    SX_APROJ_(512, 4)
    SX_APROJ_(1024, 2)
    SX_APROJ_(2048, 1)
  }
#undef SX_APROJ_
  return -1;
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
  // paired six-transform kernel from one warp per pencil upwards, four CTAs per SM (five spill at 204 registers: 2.93 ms
  // against 2.49 ms, profiles/r2g); SX_PJ=9: the slot kernel the short pencils use
  if constexpr (N >= 256 && N <= 2048) {
    if ((p.knob_pj == 0 && N >= p.knob_tma_min) || p.knob_pj == 10)
      return run_project_pair<N, (N <= 512 ? 4 : (N == 1024 ? 2 : 1))>(p, f, vx, vy, vz, pr, o, zs, ze);
  }
  return run_project_v<N, NPB, true, (N <= 512 ? 3 : 1)>(p, f, vx, vy, vz, pr, o, zs, ze);
}

int fused_project(Plan& p, Fused& f, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int o, const double* zs, const double* ze) {
#define C_(N) run_project<N>(p, f, vx, vy, vz, pr, o, zs, ze)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}

}  // namespace sx
