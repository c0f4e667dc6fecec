// Fused substep: FC-Gram continuation + z-FFT + fc_filter + RK update (see sx_fused.cu for the pass structure).
#include "sx_fused.h"

namespace sx {

// ------------------------------------------------------------------------------------------
// zfwd_rk: one CTA = NP adjacent ky pencils of one kx.  Reads the nonlinear term in the exchange
// layout, continues it, transforms, filters and performs the RK update of one velocity component.
// ------------------------------------------------------------------------------------------

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

// BATCH: how the three (four) spectral pencils of the RK update are loaded.  true (N = 512, 128 registers, two CTAs per
// SM): the linear-term pencil travels under the transform, the forcing and the RK base are loaded together after it
// (16 loads in flight per thread; 1.074 -> 0.988 ms per launch, profiles/r1k_session5.md).  false: at the point of use.
// Variants that lost in round 1 (all three pencils hoisted, L2 prefetch of the next tile's pencils, tensor-map box loads
// of the nonlinear term) are in the history and in profiles/r1i_session4.md.
template <int N, int NP, int MINB, bool PF, bool BATCH>
__global__ void __launch_bounds__(NP*(N / 8), MINB) k_zfwd_rk(ZfwdArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8, NT = NP * T;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  TwRegs<N> twr;
  twr.load(tw, j);
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
    const size_t base = ((size_t)kxl * a.ny + (active ? ky : 0)) * N;
    cplx L[8];
    if (BATCH) {
#pragma unroll
      for (int k = 0; k < 8; ++k) L[k] = a.v[base + j + k * T];
    }
    __syncthreads();  // bnd and the exchange buffer of the previous tile are free
    stash_boundary_tile<N, NP>(v, j, p, bnd, a.nph, a.d);
    __syncthreads();
    fc_continue_tile<N, NP>(v, j, p, bnd, a.nph, a.C, a.d, a.dir);
    fft_regs<N, -1>(v, j, smem, SIdxPencil{p, NP}, twr);
    if (active) {
      const double x = __ldg(&a.kx[kxl]), y = __ldg(&a.ky[ky]);
      const double f1 = __ldg(&a.fx[kxl]), f2 = __ldg(&a.fy[ky]);
      const double kh2 = x * x + y * y;
      if (BATCH) {
        // same arithmetic, association and order as the one-field-at-a-time form below; only the loads move
        if (a.couple != nullptr) {
          cplx Q[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) Q[k] = a.couple[base + j + k * T];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = caxpy(a.ccoef, Q[k], v[k]);
        }
        cplx B[8], F[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          F[k] = a.f[base + j + k * T];
          B[k] = a.v0[base + j + k * T];
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
        for (int k = 0; k < 8; ++k) a.vout[base + j + k * T] = cmake(B[k].x + v[k].x, B[k].y + v[k].y);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = j + k * T;
          const double z = __ldg(&a.kz[e]), f3 = __ldg(&a.fz[e]);
          const double lm = a.lap ? -(kh2 + z * z) : 1.0;
          cplx NL = v[k];
          if (a.couple != nullptr) NL = caxpy(a.ccoef, a.couple[base + e], NL);
          NL = cscale(cscale(cscale(NL, f1), f2), f3);
          const cplx Lk = a.v[base + e], Bk = a.v0[base + e], Fk = a.f[base + e];
          a.vout[base + e] = cmake(Bk.x + a.dt * (a.cL * (lm * Lk.x) + a.sNL * NL.x + Fk.x) * a.rmp,
                                   Bk.y + a.dt * (a.cL * (lm * Lk.y) + a.sNL * NL.y + Fk.y) * a.rmp);
        }
      }
    }
  }
}

template <int N> static int run_zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0,
                                        const cplx* frc, const RkTerm& rk, double dt, double rmp) {
  constexpr int NP = TileNP<N>::value, MINB = TileMinB<N>::value;
  ZfwdArgs a{nl, v, vout, v0, frc, rk.couple, rk.ccoef, rk.cL, rk.sNL, rk.lap, f.d_zmap, p.d_kx, p.d_ky, p.d_kz,
             p.d_fx, p.d_fy, p.d_fz, p.d_dir, p.ny, p.nxl, f.nph, p.Cz, p.oz, dt, rmp};
  const cplx* tw = p.tw_z;
  const size_t smem = ((size_t)2 * NP * N + (size_t)2 * kMaxDF * NP) * sizeof(cplx) + (size_t)N * sizeof(ZMap);
  int grid;
  if (!(p.knob_pf & 8)) {   // SX_TILE_PF bit 3 off: no cp.async prefetch of the nonlinear-term tile
    auto kfn = k_zfwd_rk<N, NP, MINB, false, false>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZFWD_RK, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  } else {
    // N = 512: 128 registers, two CTAs per SM, batched loads of the RK pencils (profiles/r1i_session4.md)
    constexpr bool B512 = N == 512;
    auto kfn = k_zfwd_rk<N, NP, (B512 ? 2 : MINB), true, B512>;
    if (persistent_grid(p, kfn, NP * (N / 8), smem, cdiv(p.ny, NP) * p.nxl, &grid)) return 1;
    SX_FUSED_LAUNCH(p, ST_ZFWD_RK, kfn, dim3(grid), NP * (N / 8), smem, a, tw);
  }
  return 0;
}
int fused_zfwd_rk(Plan& p, Fused& f, const cplx* nl, const cplx* v, cplx* vout, const cplx* v0, const cplx* frc,
                  const RkTerm& rk, double dt, double rmp) {
#define C_(N) run_zfwd_rk<N>(p, f, nl, v, vout, v0, frc, rk, dt, rmp)
  SX_SIZE_SWITCH(p.nz, C_);
#undef C_
}

}  // namespace sx
