// Fused substep: the x pass (12 c2r + products + 3 r2c per line pair) for gradre/advect and the cross products.
#include "sx_fused.h"

namespace sx {

// ------------------------------------------------------------------------------------------
// xpass (gradre): one CTA = LP pairs of adjacent y lines of one local z row.  Two real lines ride
// one complex FFT of length nx:  Z(k) = A(k) + i B(k) on the Hermitian-completed spectra.  FFTW's
// c2r ignores Im of the kx = 0 and kx = nx/2 entries; so do we (after the i kx factor, as the
// reference applies derivk before the transform).
// ------------------------------------------------------------------------------------------
struct XpassArgs {
  const cplx* V[12];  // q(NC), dy q(NC), dz q(NC) with q = (vx, vy, vz[, theta])   [zl][y][kx]
  const cplx* U[3];   // advecting velocity when it is not q(0..2) (scalar advection, NC = 1); U[0] == nullptr: V[0..2]
  cplx* X[4];
  const double* kx;  // GLOBAL kx(1:nx/2+1)
  int ny, nxp, nzf;
  double tmp;        // 1/(nx ny nz)^2
  double dkx;        // kx(i) = (i-1) dkx for i <= nx/2, kx(nx/2+1) = -(nx/2) dkx (specter.fpp:772-789)
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
  TwRegs<N> twr;
  twr.load(tw, t);
  const SIdxElem si{lp * XS};
  cplx* park = smem + (size_t)LP * XS + threadIdx.x;            // park[(c*8+k)*NT]
  cplx* slot = smem + (size_t)LP * XS + (size_t)24 * NT + threadIdx.x;  // slot[(2k+h)*NT]
  const int groups_y = cdiv(a.ny, 2 * LP), ngroups = groups_y * a.nzf;
  constexpr int NM = 3 + 3 * NC;  // inverse transforms per group: u(3), then d_x, d_y, d_z of each component
  auto field_of = [&](int m) -> const cplx* {
    if (m < 3) return a.U[0] != nullptr ? a.U[m] : a.V[m];
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
      fft_regs<N, 1>(v, t, smem, si, twr);
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
          fft_regs<N, -1>(acc, t, smem, si, twr);
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
// xpass (gradre), bulk-copy version: one CTA = one pair of adjacent y lines (N/8 threads), persistent.
// The two half-spectrum rows of every inverse transform are contiguous in [zl][y][kx], so one elected
// thread streams them into a ring of S shared-memory stages with cp.async.bulk (TMA) S-1 transforms ahead:
// the bytes in flight per SM no longer depend on the number of resident warps, and the copies never touch
// the load/store pipe.  The transforms run derivative-direction-major (u_d, then d_d q_c for every c), so
// the velocity line lives in registers for exactly NC transforms and the NC accumulators stay in
// registers: no shared-memory parking.  The sums are formed in the same order as before (d = x, y, z).
// ------------------------------------------------------------------------------------------
template <int N> struct XpassBulk {
  static constexpr int NXP = (N / 2 + 1 + 7) / 8 * 8;   // == Fused::nxp
  static constexpr int XS = sidx_elem_stride<N>();
  static constexpr size_t smem_bytes(int S) { return ((size_t)XS + (size_t)S * 2 * NXP) * sizeof(cplx) + (size_t)S * 8; }
};

__device__ __forceinline__ void accum(cplx (&acc)[8], const cplx (&u)[8], const cplx (&v)[8], bool first) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double px = u[k].x * v[k].x, py = u[k].y * v[k].y;
    acc[k] = first ? cmake(px, py) : cmake(acc[k].x + px, acc[k].y + py);
  }
}
__device__ __forceinline__ void accum(cplx (&)[1], const cplx (&)[8], const cplx (&)[8], bool) {}

template <int N, int NC, int S, int MINB>
__global__ void __launch_bounds__(N / 8, MINB) k_xpass_gradre_bulk(XpassArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, XS = XpassBulk<N>::XS, NXP = XpassBulk<N>::NXP;
  constexpr int NM = 3 * (NC + 1);   // inverse transforms per group, direction-major: u_d, d_d q_0 .. d_d q_{NC-1}
  constexpr unsigned BYTES = 2 * NXP * sizeof(cplx);
  const int t = threadIdx.x;
  TwRegs<N> twr;
  twr.load(tw, t);
  const SIdxElem si{0};
  cplx* stage = smem + XS;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(stage + (size_t)S * 2 * NXP);
  const int groups_y = a.ny / 2, ngroups = groups_y * a.nzf;
  const int mine = ((int)blockIdx.x < ngroups) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto row_of = [&](int gi) -> size_t {
    const int g = blockIdx.x + gi * gridDim.x;
    return ((size_t)(g / groups_y) * a.ny + (size_t)(g % groups_y) * 2) * NXP;
  };
  // producer state (only meaningful in the lead thread): next load = transform (pd, pi) of group pg into stage ps.
  // All counters are advanced incrementally: no divisions in the per-transform path.
  int pg = 0, pd = 0, pi = 0, ps = 0;
  size_t prow = mine > 0 ? row_of(0) : 0;
  auto issue_next = [&]() {
    if (pg >= mine) return;
    const cplx* field = pi == 0 ? (a.U[0] != nullptr ? a.U[pd] : a.V[pd]) : a.V[pd * NC + pi - 1];
    bulk_load(stage + (size_t)ps * 2 * NXP, field + prow, BYTES, bar + ps);
    if (++ps == S) ps = 0;
    if (++pi == NC + 1) {
      pi = 0;
      if (++pd == 3) {
        pd = 0;
        if (++pg < mine) prow = row_of(pg);
      }
    }
  };
  if (t == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar + s, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (t == 0)
    for (int q = 0; q < S - 1; ++q) issue_next();
  int cs = 0;            // consumer stage
  unsigned cph = 0;      // and its mbarrier parity
  for (int gi = 0; gi < mine; ++gi) {
    const size_t rowA = row_of(gi), rowB = rowA + NXP;
    cplx acc0[8], acc1[8], acc2[8], acc3[NC > 3 ? 8 : 1], u[8];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
#pragma unroll 1
      for (int i = 0; i <= NC; ++i) {
        // the stage consumed by the previous transform was read before that transform's barriers: refill it
        if (t == 0) issue_next();
        mbar_wait(bar + cs, cph);
        const cplx* sA = stage + (size_t)cs * 2 * NXP;
        const cplx* sB = sA + NXP;
        if (++cs == S) {
          cs = 0;
          cph ^= 1;
        }
        const bool deriv = d == 0 && i > 0;
        cplx v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = t + k * T;
          const int kx = e <= N / 2 ? e : N - e;
          cplx A = sA[kx], B = sB[kx];
          if (deriv) {
            const double kk = (double)(kx == N / 2 ? -kx : kx) * a.dkx;   // the product the host table holds
            A = cmake(-kk * A.y, kk * A.x);
            B = cmake(-kk * B.y, kk * B.x);
          }
          if (kx == 0 || kx == N / 2) { A.y = 0.0; B.y = 0.0; }
          if (e > N / 2) { A.y = -A.y; B.y = -B.y; }
          v[k] = cmake(A.x - B.y, A.y + B.x);
        }
        fft_regs<N, 1>(v, t, smem, si, twr);
        if (i == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) u[k] = v[k];
        } else {
          const bool f0 = d == 0;
          if (i == 1) accum(acc0, u, v, f0);
          else if (i == 2) accum(acc1, u, v, f0);
          else if (i == 3) accum(acc2, u, v, f0);
          else if (NC > 3) accum(acc3, u, v, f0);
        }
      }
    }
    // forward transforms of the packed pairs and split into the two half spectra
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
      cplx w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        cplx q = acc0[k];
        if (c == 1) q = acc1[k];
        if (c == 2) q = acc2[k];
        if (NC > 3 && c == 3) q = acc3[NC > 3 ? k : 0];
        w[k] = cmake(q.x * a.tmp, q.y * a.tmp);
      }
      fft_regs<N, -1>(w, t, smem, si, twr);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) smem[si(t + k * T)] = w[k];
      __syncthreads();
      cplx* out = a.X[c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = t + k * T;
        if (kk <= N / 2) {
          const cplx Zk = w[k];
          const cplx Zn = smem[si((N - kk) & (N - 1))];
          out[rowA + kk] = cmake(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
          out[rowB + kk] = cmake(0.5 * (Zk.y + Zn.y), -0.5 * (Zk.x - Zn.x));
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
  TwRegs<N> twr;
  twr.load(tw, t);
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
      fft_regs<N, 1>(v, t, smem, si, twr);
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
      fft_regs<N, -1>(acc, t, smem, si, twr);
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
// xpass (cross products), bulk-copy version: the ring of k_xpass_gradre_bulk (one elected thread streams the two
// half-spectrum rows of every inverse transform into S shared-memory stages with cp.async.bulk, S-1 transforms ahead)
// for X = sum over pairs s (P x Q) / N^2.  The six lines of a pair are transformed in the order of the cycle
// Px - Qy - Pz - Qx - Py - Qz (- Px) of the products a cross product needs, so that only the previous line and Px
// have to be kept: Px waits in a thread-private shared-memory park (written once, read twice), the previous line and
// the three accumulators live in registers.
//   z += Px Qy;  x -= Pz Qy;  y += Pz Qx;  z -= Py Qx;  x += Py Qz;  y -= Px Qz
// ------------------------------------------------------------------------------------------
template <int N, int S, int MINB>
__global__ void __launch_bounds__(N / 8, MINB) k_xpass_cross_bulk(XcrossArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, XS = XpassBulk<N>::XS, NXP = XpassBulk<N>::NXP;
  constexpr unsigned BYTES = 2 * NXP * sizeof(cplx);
  const int t = threadIdx.x;
  TwRegs<N> twr;
  twr.load(tw, t);
  const SIdxElem si{0};
  cplx* stage = smem + XS;
  cplx* park = stage + (size_t)S * 2 * NXP + t;    // park[k * T]: Px of the current pair
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(stage + (size_t)S * 2 * NXP + N);
  const int groups_y = a.ny / 2, ngroups = groups_y * a.nzf;
  const int mine = ((int)blockIdx.x < ngroups) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto row_of = [&](int gi) -> size_t {
    const int g = blockIdx.x + gi * gridDim.x;
    return ((size_t)(g / groups_y) * a.ny + (size_t)(g % groups_y) * 2) * NXP;
  };
  // producer state (lead thread): next load = line pm of pair pp of group pg into stage ps
  int pg = 0, pp = 0, pm = 0, ps = 0;
  size_t prow = mine > 0 ? row_of(0) : 0;
  auto issue_next = [&]() {
    if (pg >= mine) return;
    // cycle order Px Qy Pz Qx Py Qz
    const cplx* field = pm == 0 ? a.P[pp][0] : pm == 1 ? a.Q[pp][1] : pm == 2 ? a.P[pp][2] : pm == 3 ? a.Q[pp][0]
                      : pm == 4 ? a.P[pp][1] : a.Q[pp][2];
    bulk_load(stage + (size_t)ps * 2 * NXP, field + prow, BYTES, bar + ps);
    if (++ps == S) ps = 0;
    if (++pm == 6) {
      pm = 0;
      if (++pp == a.npairs) {
        pp = 0;
        if (++pg < mine) prow = row_of(pg);
      }
    }
  };
  if (t == 0) {
    for (int q = 0; q < S; ++q) mbar_init(bar + q, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (t == 0)
    for (int q = 0; q < S - 1; ++q) issue_next();
  int cs = 0;
  unsigned cph = 0;
  for (int gi = 0; gi < mine; ++gi) {
    const size_t rowA = row_of(gi), rowB = rowA + NXP;
    cplx ax[8], ay[8], az[8], prev[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ax[k] = ay[k] = az[k] = prev[k] = cmake(0.0, 0.0);
#pragma unroll 1
    for (int pr = 0; pr < a.npairs; ++pr) {
      const double sg = a.sgn[pr];
#pragma unroll 1
      for (int m = 0; m < 6; ++m) {
        if (t == 0) issue_next();   // the stage consumed by the previous transform was read before its barriers
        mbar_wait(bar + cs, cph);
        const cplx* sA = stage + (size_t)cs * 2 * NXP;
        const cplx* sB = sA + NXP;
        if (++cs == S) {
          cs = 0;
          cph ^= 1;
        }
        cplx v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = t + k * T;
          const int kx = e <= N / 2 ? e : N - e;
          cplx A = sA[kx], B = sB[kx];
          if (kx == 0 || kx == N / 2) { A.y = 0.0; B.y = 0.0; }
          if (e > N / 2) { A.y = -A.y; B.y = -B.y; }
          v[k] = cmake(A.x - B.y, A.y + B.x);
        }
        fft_regs<N, 1>(v, t, smem, si, twr);
        // products with the previous line of the cycle (and with Px at its end)
        if (m == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) park[k * T] = v[k];
        } else if (m == 1) {      // Qy: z += Px Qy
#pragma unroll
          for (int k = 0; k < 8; ++k) az[k] = cmake(fma(sg * prev[k].x, v[k].x, az[k].x), fma(sg * prev[k].y, v[k].y, az[k].y));
        } else if (m == 2) {      // Pz: x -= Pz Qy
#pragma unroll
          for (int k = 0; k < 8; ++k) ax[k] = cmake(fma(-sg * v[k].x, prev[k].x, ax[k].x), fma(-sg * v[k].y, prev[k].y, ax[k].y));
        } else if (m == 3) {      // Qx: y += Pz Qx
#pragma unroll
          for (int k = 0; k < 8; ++k) ay[k] = cmake(fma(sg * prev[k].x, v[k].x, ay[k].x), fma(sg * prev[k].y, v[k].y, ay[k].y));
        } else if (m == 4) {      // Py: z -= Py Qx
#pragma unroll
          for (int k = 0; k < 8; ++k) az[k] = cmake(fma(-sg * v[k].x, prev[k].x, az[k].x), fma(-sg * v[k].y, prev[k].y, az[k].y));
        } else {                  // Qz: x += Py Qz, y -= Px Qz
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const cplx px = park[k * T];
            ax[k] = cmake(fma(sg * prev[k].x, v[k].x, ax[k].x), fma(sg * prev[k].y, v[k].y, ax[k].y));
            ay[k] = cmake(fma(-sg * px.x, v[k].x, ay[k].x), fma(-sg * px.y, v[k].y, ay[k].y));
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) prev[k] = v[k];
      }
    }
    // forward transforms of the three packed pairs and split into the half spectra
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      cplx w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const cplx s3 = c == 0 ? ax[k] : (c == 1 ? ay[k] : az[k]);
        w[k] = cmake(s3.x * a.tmp, s3.y * a.tmp);
      }
      fft_regs<N, -1>(w, t, smem, si, twr);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) smem[si(t + k * T)] = w[k];
      __syncthreads();
      cplx* out = a.X[c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = t + k * T;
        if (kk <= N / 2) {
          const cplx Zk = w[k];
          const cplx Zn = smem[si((N - kk) & (N - 1))];
          out[rowA + kk] = cmake(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
          out[rowB + kk] = cmake(0.5 * (Zk.y + Zn.y), -0.5 * (Zk.x - Zn.x));
        }
      }
    }
  }
}

// ui < 0: gradre / advect of q = V[0..NC) by its own first three components into X[0..NC);
// ui >= 0 (NC = 1): the scalar q = V[qi], dy q = V[qi+1], dz q = V[qi+2] advected by V[ui..ui+2] into X[xo]
static void xpass_fields(Plan& p, Fused& f, XpassArgs& a, int NC, int ui, int qi, int xo) {
  const size_t zo = f.vz0() * p.ny * f.nxp;   // z window of the [zl][y][kx] arrays
  for (int i = 0; i < 12; ++i) a.V[i] = nullptr;
  for (int i = 0; i < 4; ++i) a.X[i] = nullptr;
  for (int i = 0; i < 3; ++i) a.U[i] = ui >= 0 ? f.V[ui + i] + zo : nullptr;
  for (int i = 0; i < 3 * NC; ++i) a.V[i] = f.V[(ui >= 0 ? qi : 0) + i] + zo;
  for (int i = 0; i < NC; ++i) a.X[i] = f.X[(ui >= 0 ? xo : 0) + i] + zo;
}
template <int N, int LP, bool PF, int MINB, int NC> static int run_xpass_v(Plan& p, Fused& f, const double* d_kx_global, int ui, int qi, int xo) {
  constexpr int T = N / 8;
  if (f.zc() == 0) return 0;
  XpassArgs a;
  xpass_fields(p, f, a, NC, ui, qi, xo);
  a.kx = d_kx_global;
  a.ny = p.ny;
  a.nxp = f.nxp;
  a.nzf = f.zc();
  const double Ntot = (double)p.nx * (double)p.ny * (double)p.nz;
  a.tmp = 1.0 / (Ntot * Ntot);
  a.dkx = p.Dkx;
  const cplx* tw = p.tw_x;
  auto kfn = k_xpass_gradre<N, LP, PF, MINB, NC>;
  const size_t smem = ((size_t)LP * sidx_elem_stride<N>() + (size_t)(PF ? 40 : 24) * LP * T) * sizeof(cplx);
  int grid;
  if (persistent_grid(p, kfn, LP * T, smem, cdiv(p.ny, 2 * LP) * f.zc(), &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), LP * T, smem, a, tw);
  return 0;
}
template <int N, int NC, int S, int MINB> static int run_xpass_bulk(Plan& p, Fused& f, const double* d_kx_global, int ui, int qi, int xo) {
  constexpr int T = N / 8;
  if (f.zc() == 0) return 0;
  SX_REQUIRE(f.nxp == XpassBulk<N>::NXP && p.ny % 2 == 0, "xpass (bulk): unexpected row padding");
  XpassArgs a;
  xpass_fields(p, f, a, NC, ui, qi, xo);
  a.kx = d_kx_global;
  a.ny = p.ny;
  a.nxp = f.nxp;
  a.nzf = f.zc();
  const double Ntot = (double)p.nx * (double)p.ny * (double)p.nz;
  a.tmp = 1.0 / (Ntot * Ntot);
  a.dkx = p.Dkx;
  const cplx* tw = p.tw_x;
  auto kfn = k_xpass_gradre_bulk<N, NC, S, MINB>;
  const size_t smem = XpassBulk<N>::smem_bytes(S);
  int grid;
  if (persistent_grid(p, kfn, T, smem, (p.ny / 2) * f.zc(), &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), T, smem, a, tw);
  return 0;
}
template <int N, int NC> static int run_xpass(Plan& p, Fused& f, const double* d_kx_global, int ui = -1, int qi = 0, int xo = 0) {
  constexpr int T = N / 8;
  constexpr int LP = T >= 128 ? 1 : 128 / T;
  // bulk-copy ring (TMA) + register accumulators from one warp per line pair upwards: 255 registers (accumulators, velocity
  // line and one transform live in the register file), four two-warp CTAs per SM at N = 512, ring of three stages.
  // SX_XP=9: the cp.async slot kernel the short lines use.
  // SX_XP=10 forces it from N = 64 (the emulation tests run it on small grids).
  if constexpr (N >= 64 && N <= 2048) {
    if (p.knob_xp == 0 ? N >= 256 : p.knob_xp == 10)
      return run_xpass_bulk<N, NC, 3, (N <= 512 ? 4 : (N == 1024 ? 2 : 1))>(p, f, d_kx_global, ui, qi, xo);
  }
  return run_xpass_v<N, LP, true, (N <= 1024 ? 2 : 1), NC>(p, f, d_kx_global, ui, qi, xo);
}
// X[xo..xo+2] = sum_pairs sgn * (V[P] x V[Q]) / N^2; Pi/Qi index the first of three consecutive V fields
template <int N> static int run_xcross(Plan& p, Fused& f, int npairs, const int* Pi, const int* Qi, const double* sgn, int xo) {
  constexpr int T = N / 8;
  constexpr int LP = T >= 128 ? 1 : 128 / T;
  if (f.zc() == 0) return 0;
  XcrossArgs a;
  for (int q = 0; q < 2; ++q)
    for (int c = 0; c < 3; ++c) {
      a.P[q][c] = f.V[Pi[q < npairs ? q : 0] + c] + f.vz0() * p.ny * f.nxp;
      a.Q[q][c] = f.V[Qi[q < npairs ? q : 0] + c] + f.vz0() * p.ny * f.nxp;
    }
  a.sgn[0] = sgn[0];
  a.sgn[1] = npairs > 1 ? sgn[1] : 0.0;
  a.npairs = npairs;
  for (int c = 0; c < 3; ++c) a.X[c] = f.X[xo + c] + f.vz0() * p.ny * f.nxp;
  a.ny = p.ny;
  a.nxp = f.nxp;
  a.nzf = f.zc();
  const double Ntot = (double)p.nx * (double)p.ny * (double)p.nz;
  a.tmp = 1.0 / (Ntot * Ntot);
  const cplx* tw = p.tw_x;
  // bulk-copy ring from one warp per line pair upwards (SX_XP=9: the previous kernel with cp.async slots and parked P lines)
  if constexpr (N >= 64 && N <= 2048) {
    if ((p.knob_xp == 0 ? N >= 256 : p.knob_xp == 10) && f.nxp == XpassBulk<N>::NXP && p.ny % 2 == 0) {
      constexpr int S = 3, MINB = N <= 512 ? 4 : (N == 1024 ? 2 : 1);
      auto kfn = k_xpass_cross_bulk<N, S, MINB>;
      const size_t smem = ((size_t)XpassBulk<N>::XS + (size_t)S * 2 * XpassBulk<N>::NXP + (size_t)N) * sizeof(cplx) + (size_t)S * 8;
      int grid;
      if (persistent_grid(p, kfn, T, smem, (p.ny / 2) * f.zc(), &grid)) return 1;
      SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), T, smem, a, tw);
      return 0;
    }
  }
  auto kfn = k_xpass_cross<N, LP, (N <= 1024 ? 2 : 1)>;
  const size_t smem = ((size_t)LP * sidx_elem_stride<N>() + (size_t)40 * LP * T) * sizeof(cplx);
  int grid;
  if (persistent_grid(p, kfn, LP * T, smem, cdiv(p.ny, 2 * LP) * f.zc(), &grid)) return 1;
  SX_FUSED_LAUNCH(p, ST_XPASS, kfn, dim3(grid), LP * T, smem, a, tw);
  return 0;
}
template <int NC> static int xpass_nc(Plan& p, Fused& f, const double* kxg) {
#define C_(N) run_xpass<N, NC>(p, f, kxg)
  SX_SIZE_SWITCH(p.nx, C_);
#undef C_
}
int fused_xpass(Plan& p, Fused& f, int nc, const double* kxg) {
  if (nc == 3) return xpass_nc<3>(p, f, kxg);
  if (nc == 4) return xpass_nc<4>(p, f, kxg);
  SX_REQUIRE(false, "xpass: 3 or 4 advected components");
}
// scalar advection: X[xo] = (V[ui..ui+2] . grad q) / N^2 with q = V[qi], dy q = V[qi+1], dz q = V[qi+2]
// (advect, pseudospec_phd.f90:58-110, next to a cross-product x pass whose velocity is already in V)
int fused_xadvect(Plan& p, Fused& f, int ui, int qi, int xo, const double* kxg) {
#define C_(N) run_xpass<N, 1>(p, f, kxg, ui, qi, xo)
  SX_SIZE_SWITCH(p.nx, C_);
#undef C_
}
int fused_xcross(Plan& p, Fused& f, int npairs, const int* Pi, const int* Qi, const double* sgn, int xo) {
#define C_(N) run_xcross<N>(p, f, npairs, Pi, Qi, sgn, xo)
  SX_SIZE_SWITCH(p.nx, C_);
#undef C_
}

}  // namespace sx
