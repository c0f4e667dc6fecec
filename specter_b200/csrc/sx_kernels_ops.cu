// Stand-alone (per-operator) spectral, real-space and boundary kernels.  These back the
// per-operator C-ABI entry points that mirror the reference's pseudo / boundary modules;
// the fused RK substep (sx_rkstep.cu) does the same arithmetic inside the FFT passes.
// All arrays: spectral/mixed (nz, ny, nxl) z-fastest; real (nx, ny, nzl) x-fastest.
#include "sx_plan.h"

namespace sx {

struct Dims {
  int nz, ny, nxl;
  size_t n;
};
static inline Dims dims_of(const Plan& p) { return Dims{p.nz, p.ny, p.nxl, p.csize()}; }

#define SX_GRID_STRIDE(idx, n) \
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (n); idx += (size_t)gridDim.x * blockDim.x)

static inline unsigned ew_grid(size_t n, int threads = 256) {
  size_t g = (n + threads - 1) / threads;
  const size_t cap = 148u * 16u;  // B200: 148 SMs, grid-stride beyond that
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

// ---- derivk / laplak / curlk / fc_filter  (pseudospec_hd.f90:28-206, 1082-1115) ----
__global__ void k_derivk(Dims d, const cplx* __restrict__ a, cplx* __restrict__ b,
                         const double* __restrict__ kv, int dir) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double kk = __ldg(&kv[dir == 1 ? i : (dir == 2 ? j : k)]);
    const cplx v = a[idx];
    b[idx] = cmake(-kk * v.y, kk * v.x);
  }
}

__global__ void k_laplak(Dims d, const cplx* __restrict__ a, cplx* __restrict__ b,
                         const double* __restrict__ kx, const double* __restrict__ ky,
                         const double* __restrict__ kz) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]), z = __ldg(&kz[k]);
    const double kk2 = x * x + y * y + z * z;
    const cplx v = a[idx];
    b[idx] = cmake(-kk2 * v.x, -kk2 * v.y);
  }
}

// curlk(a,b,c,dir): dir=1: c = i ky b - i kz a ; dir=2: c = i kz a - i kx b ; dir=3: c = i kx b - i ky a
__global__ void k_curlk(Dims d, const cplx* __restrict__ a, const cplx* __restrict__ b,
                        cplx* __restrict__ c, const double* __restrict__ kx,
                        const double* __restrict__ ky, const double* __restrict__ kz, int dir) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const cplx A = a[idx], B = b[idx];
    double ka, kb;
    if (dir == 1) { ka = __ldg(&kz[k]); kb = __ldg(&ky[j]); }
    else if (dir == 2) { ka = __ldg(&kz[k]); kb = __ldg(&kx[i]); }
    else { ka = __ldg(&ky[j]); kb = __ldg(&kx[i]); }
    const cplx c1 = cmake(-ka * A.y, ka * A.x);  // i ka a
    const cplx c2 = cmake(-kb * B.y, kb * B.x);  // i kb b
    c[idx] = (dir == 2) ? csub(c1, c2) : csub(c2, c1);
  }
}

__global__ void k_fc_filter(Dims d, cplx* __restrict__ a, const double* __restrict__ fx,
                            const double* __restrict__ fy, const double* __restrict__ fz) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double f1 = __ldg(&fx[i]), f2 = __ldg(&fy[j]), f3 = __ldg(&fz[k]);
    cplx v = a[idx];
    v = cscale(cscale(cscale(v, f1), f2), f3);
    a[idx] = v;
  }
}

__global__ void k_copy(size_t n, const cplx* __restrict__ a, cplx* __restrict__ b) {
  SX_GRID_STRIDE(idx, n) b[idx] = a[idx];
}

__global__ void k_scale_copy(size_t n, const cplx* __restrict__ a, cplx* __restrict__ b, double s) {
  SX_GRID_STRIDE(idx, n) b[idx] = cscale(a[idx], s);
}
__global__ void k_add(size_t n, cplx* __restrict__ a, const cplx* __restrict__ b) {
  SX_GRID_STRIDE(idx, n) a[idx] = cadd(a[idx], b[idx]);
}

__global__ void k_scale_phys(Dims d, cplx* __restrict__ a, int nphys, double s) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    if (k < nphys) a[idx] = cscale(a[idx], s);
  }
}

// RK update (hd_rkstep2.f90:19-32): v = v0 + dt*(nu*v - nl + f)*rmp, v holds lap(v) on entry
__global__ void k_rk_axpy(size_t n, cplx* __restrict__ v, const cplx* __restrict__ v0,
                          const cplx* __restrict__ nl, const cplx* __restrict__ f, double dt,
                          double nu, double rmp) {
  SX_GRID_STRIDE(idx, n) {
    const cplx L = v[idx], B = v0[idx], NL = nl[idx], F = f[idx];
    v[idx] = cmake(B.x + dt * (nu * L.x - NL.x + F.x) * rmp, B.y + dt * (nu * L.y - NL.y + F.y) * rmp);
  }
}

// gradre products (pseudospec_hd.f90:254-312): r[4*dir+0] = A_dir, r[4*dir+1..3] = d_dir A_{x,y,z}
struct R12 { const double* r[12]; };
__global__ void k_gradre_products(size_t n, R12 in, double* __restrict__ rx, double* __restrict__ ry,
                                  double* __restrict__ rz, double tmp) {
  SX_GRID_STRIDE(idx, n) {
    double sx_ = 0.0, sy_ = 0.0, sz_ = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double u = in.r[4 * d][idx];
      sx_ += u * in.r[4 * d + 1][idx];
      sy_ += u * in.r[4 * d + 2][idx];
      sz_ += u * in.r[4 * d + 3][idx];
    }
    rx[idx] = sx_ * tmp;
    ry[idx] = sy_ * tmp;
    rz[idx] = sz_ * tmp;
  }
}

// cross product a x b (prodre pseudospec_hd.f90:363-371 with a=curl, b=field; vector pseudospec_mhd.f90:87-102)
__global__ void k_cross_products(size_t n, const double* __restrict__ a1, const double* __restrict__ a2,
                                 const double* __restrict__ a3, const double* __restrict__ b1,
                                 const double* __restrict__ b2, const double* __restrict__ b3,
                                 double* __restrict__ rx, double* __restrict__ ry,
                                 double* __restrict__ rz, double tmp) {
  SX_GRID_STRIDE(idx, n) {
    const double A1 = a1[idx], A2 = a2[idx], A3 = a3[idx];
    const double B1 = b1[idx], B2 = b2[idx], B3 = b3[idx];
    rx[idx] = (A2 * B3 - B2 * A3) * tmp;
    ry[idx] = (A3 * B1 - B3 * A1) * tmp;
    rz[idx] = (A1 * B2 - B1 * A2) * tmp;
  }
}

// ---- projection pieces (boundary_mod.fpp:197-448) -----------------------------------
// d = -i (k.v)/k^2 (0 at the mean mode); v -= i k d ; C1 = (bctarget ? c : d)/nz
__global__ void k_proj_inh(Dims d, cplx* __restrict__ a, cplx* __restrict__ b, cplx* __restrict__ c,
                           cplx* __restrict__ dd, cplx* __restrict__ C1, const double* __restrict__ kx,
                           const double* __restrict__ ky, const double* __restrict__ kz, int bctarget,
                           int has_mean, double inv_nz) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]), z = __ldg(&kz[k]);
    const double kk2 = x * x + y * y + z * z;
    cplx A = a[idx], B = b[idx], C = c[idx];
    const cplx s = cmake(x * A.x + y * B.x + z * C.x, x * A.y + y * B.y + z * C.y);
    cplx D = cmake(s.y / kk2, -s.x / kk2);  // -i*s/kk2
    if (has_mean && idx == 0) D = cmake(0.0, 0.0);
    // v -= i k D
    A = cmake(A.x + x * D.y, A.y - x * D.x);
    B = cmake(B.x + y * D.y, B.y - y * D.x);
    C = cmake(C.x + z * D.y, C.y - z * D.x);
    a[idx] = A; b[idx] = B; c[idx] = C; dd[idx] = D;
    C1[idx] = cscale(bctarget ? C : D, inv_nz);
  }
}

// laplace_z (boundary_mod.fpp:451-678): closed-form harmonic solution per (ky,kx) pencil from
// the wall values of C1 (after its z-IFFT).  Supported: Dirichlet (0,0), Neumann (1,1),
// Robin (2,2), Dirichlet/Robin (0,2).  C2in holds IFFT_z(i kz C1) for the Robin cases.
__global__ void k_laplace_z(Dims d, const cplx* __restrict__ C1, const cplx* __restrict__ C2in,
                            cplx* __restrict__ C2, cplx* __restrict__ C3,
                            const double* __restrict__ kx, const double* __restrict__ ky,
                            const double* __restrict__ zc, int top, int bctarget, int bczsta,
                            int bczend, int has_mean, double Lz) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]);
    const double kh = sqrt(x * x + y * y);
    const size_t pb = ji * d.nz;
    const cplx w0 = C1[pb], w1 = C1[pb + top];
    cplx bc1, bc2;
    if (bczsta == 0) bc1 = bctarget ? w0 : cmake(-w0.x, -w0.y);
    else { const cplx g = C2in[pb]; bc1 = cmake(g.x - kh * w0.x, g.y - kh * w0.y); }
    if (bczend == 0) bc2 = bctarget ? w1 : cmake(-w1.x, -w1.y);
    else { const cplx g = C2in[pb + top]; bc2 = cmake(-(g.x + kh * w1.x), -(g.y + kh * w1.y)); }
    const int ks = bczsta + bctarget, ke = bczend + bctarget;
    const double z = __ldg(&zc[k]);
    const bool mean = has_mean && ji == 0;
    cplx c1, c2;
    if (ks == 0 && ke == 0) {
      if (mean) { c1 = cmake((bc2.x - bc1.x) / Lz, (bc2.y - bc1.y) / Lz); c2 = bc1; }
      else {
        const double e1 = exp(-kh * Lz), t = 1.0 / (1 - exp(-2 * kh * Lz));
        c1 = cmake((bc2.x - bc1.x * e1) * t, (bc2.y - bc1.y * e1) * t);
        c2 = cmake((bc1.x - bc2.x * e1) * t, (bc1.y - bc2.y * e1) * t);
      }
    } else if (ks == 1 && ke == 1) {
      if (mean) { c1 = bc1; c2 = cmake(0.0, 0.0); }
      else {
        const double e1 = exp(-kh * Lz), t = 1.0 / (kh * (1 - exp(-2 * kh * Lz)));
        c1 = cmake((bc2.x - bc1.x * e1) * t, (bc2.y - bc1.y * e1) * t);
        c2 = cmake((-bc1.x + bc2.x * e1) * t, (-bc1.y + bc2.y * e1) * t);
      }
    } else if (ks == 2 && ke == 2) {
      if (mean) { c1 = bc1; c2 = cmake(0.0, 0.0); }
      else {
        const double t = 1.0 / (2 * kh);
        c1 = cscale(bc2, t);
        c2 = cscale(bc1, t);
      }
    } else {  // ks == 0 && ke == 2
      if (mean) { c1 = bc2; c2 = bc1; }
      else {
        const double e1 = exp(-kh * Lz), t = 1.0 / (2 * kh);
        c1 = cscale(bc2, t);
        c2 = cmake((bc1.x * 2 * kh - bc2.x * e1) * t, (bc1.y * 2 * kh - bc2.y * e1) * t);
      }
    }
    cplx A, B;
    if (mean) {
      A = cmake(c1.x * z + c2.x, 0.0);
      B = cmake(c1.x, 0.0);
    } else {
      const double ep = exp(kh * (z - Lz)), em = exp(-kh * z);
      A = cmake(c1.x * ep + c2.x * em, c1.y * ep + c2.y * em);
      B = cmake(kh * (c1.x * ep - c2.x * em), kh * (c1.y * ep - c2.y * em));
    }
    C2[idx] = A;
    C3[idx] = B;
  }
}

// d = d_inh + d_hom in the mixed domain (boundary_mod.fpp:361-381)
__global__ void k_pr_combine(size_t n, cplx* __restrict__ dd, const cplx* __restrict__ C1,
                             const cplx* __restrict__ C2, int bctarget, double inv_nz) {
  SX_GRID_STRIDE(idx, n) {
    const cplx h = C2[idx];
    if (bctarget) { const cplx v = dd[idx]; dd[idx] = cmake(v.x * inv_nz + h.x, v.y * inv_nz + h.y); }
    else { const cplx v = C1[idx]; dd[idx] = cmake(v.x + h.x, v.y + h.y); }
  }
}

// a -= i kx C2 ; b -= i ky C2 ; c -= C3  (boundary_mod.fpp:388-399)
__global__ void k_apply_hom(Dims d, cplx* __restrict__ a, cplx* __restrict__ b, cplx* __restrict__ c,
                            const cplx* __restrict__ C2, const cplx* __restrict__ C3,
                            const double* __restrict__ kx, const double* __restrict__ ky) {
  SX_GRID_STRIDE(idx, d.n) {
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]);
    const cplx h = C2[idx], g = C3[idx];
    cplx A = a[idx], B = b[idx], C = c[idx];
    a[idx] = cmake(A.x + x * h.y, A.y - x * h.x);
    b[idx] = cmake(B.x + y * h.y, B.y - y * h.x);
    c[idx] = cmake(C.x - g.x, C.y - g.y);
  }
}

// noslip_z both walls (vboundary.f90:154-211), vx,vy,pr in the mixed domain
__global__ void k_noslip(Dims d, cplx* __restrict__ vx, cplx* __restrict__ vy, const cplx* __restrict__ pr,
                         const double* __restrict__ kx, const double* __restrict__ ky, int top, double tmp,
                         int has_mean, double m_x0, double m_y0, double m_x1, double m_y1) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(t, 2 * npen) {
    const int wall = (int)(t / npen);
    const size_t ji = t % npen;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const size_t idx = ji * d.nz + (wall ? top : 0);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]);
    const cplx P = pr[idx];
    cplx ax = cmake(-x * P.y * tmp, x * P.x * tmp);
    cplx ay = cmake(-y * P.y * tmp, y * P.x * tmp);
    if (has_mean && ji == 0) {
      ax = cmake(wall ? m_x1 : m_x0, 0.0);
      ay = cmake(wall ? m_y1 : m_y0, 0.0);
    }
    vx[idx] = ax;
    vy[idx] = ay;
  }
}

// ---- reductions (pseudospec_hd.f90:602-631; boundary_mod.fpp:758-798) ----------------
// mode 0: sum w_i |a|^2 ; mode 1: sum w_i Re(a conj b).  row<0: all physical rows, else one row.
__global__ void k_reduce_phys(Dims d, const cplx* __restrict__ a, const cplx* __restrict__ b, int mode,
                              int nphys, int row, int first_is_mean, double scale,
                              double* __restrict__ partial) {
  SX_DYN_SMEM(double, sh);
  double acc = 0.0;
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const bool use = row < 0 ? (k < nphys) : (k == row);
    if (use) {
      const int i = (int)(idx / ((size_t)d.nz * d.ny));
      const double w = (first_is_mean && i == 0) ? 1.0 : 2.0;
      const cplx A = a[idx];
      double q;
      if (mode == 0) q = A.x * A.x + A.y * A.y;
      else { const cplx B = b[idx]; q = A.x * B.x + A.y * B.y; }
      acc += w * q * scale;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---- launchers -----------------------------------------------------------------------
#define SX_EW_LAUNCH(p, kernel, n, ...)                                        \
  do {                                                                         \
    auto kfn = kernel;                                                         \
    cudaStream_t st_ = (p).stream;                                             \
    if (stage_mark((p), ST_EW)) return 1;                                      \
    SX_LAUNCH(kfn, dim3(ew_grid(n)), dim3(256), 0, st_, __VA_ARGS__);          \
    (p).launches++;                                                            \
    SX_KERNEL_CHECK();                                                         \
  } while (0)

int op_derivk(Plan& p, const cplx* a, cplx* b, int dir) {
  SX_REQUIRE(dir >= 1 && dir <= 3, "derivk: dir must be 1..3");
  const Dims d = dims_of(p);
  const double* kv = dir == 1 ? p.d_kx : (dir == 2 ? p.d_ky : p.d_kz);
  SX_EW_LAUNCH(p, k_derivk, d.n, d, a, b, kv, dir);
  return 0;
}
int op_laplak(Plan& p, const cplx* a, cplx* b) {
  const Dims d = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky, *kz = p.d_kz;
  SX_EW_LAUNCH(p, k_laplak, d.n, d, a, b, kx, ky, kz);
  return 0;
}
int op_curlk(Plan& p, const cplx* a, const cplx* b, cplx* c, int dir) {
  SX_REQUIRE(dir >= 1 && dir <= 3, "curlk: dir must be 1..3");
  const Dims d = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky, *kz = p.d_kz;
  SX_EW_LAUNCH(p, k_curlk, d.n, d, a, b, c, kx, ky, kz, dir);
  return 0;
}
int op_fc_filter(Plan& p, cplx* a) {
  const Dims d = dims_of(p);
  const double *fx = p.d_fx, *fy = p.d_fy, *fz = p.d_fz;
  SX_EW_LAUNCH(p, k_fc_filter, d.n, d, a, fx, fy, fz);
  return 0;
}
int op_copy(Plan& p, const cplx* a, cplx* b) {
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_copy, n, n, a, b);
  return 0;
}
int op_add(Plan& p, cplx* a, const cplx* b) {
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_add, n, n, a, b);
  return 0;
}
int op_scale_copy(Plan& p, const cplx* a, cplx* b, double s) {   // b = s a on all nz rows (specter.fpp:1010-1020)
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_scale_copy, n, n, a, b, s);
  return 0;
}
int op_scale_phys(Plan& p, cplx* a, double s) {
  const Dims d = dims_of(p);
  const int nph = p.nphys();
  SX_EW_LAUNCH(p, k_scale_phys, d.n, d, a, nph, s);
  return 0;
}
int op_rk_axpy(Plan& p, cplx* v, const cplx* v0, const cplx* nl, const cplx* f, double dt, double nu,
               double rmp) {
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_rk_axpy, n, n, v, v0, nl, f, dt, nu, rmp);
  return 0;
}
int op_gradre_products(Plan& p, double* const r[12], double* rx, double* ry, double* rz) {
  // products only on the local physical planes ksta..pkend (pseudospec_hd.f90:255)
  const int nzp = p.pkend - p.ksta + 1;
  if (nzp <= 0) return 0;
  const size_t n = (size_t)p.nx * p.ny * nzp;
  R12 in;
  for (int i = 0; i < 12; ++i) in.r[i] = r[i];
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  const double tmp = 1.0 / (N * N);
  SX_EW_LAUNCH(p, k_gradre_products, n, n, in, rx, ry, rz, tmp);
  return 0;
}
int op_cross_products(Plan& p, const double* a1, const double* a2, const double* a3, const double* b1,
                      const double* b2, const double* b3, double* rx, double* ry, double* rz) {
  const int nzp = p.pkend - p.ksta + 1;
  if (nzp <= 0) return 0;
  const size_t n = (size_t)p.nx * p.ny * nzp;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  const double tmp = 1.0 / (N * N);
  SX_EW_LAUNCH(p, k_cross_products, n, n, a1, a2, a3, b1, b2, b3, rx, ry, rz, tmp);
  return 0;
}
int op_proj_inhomogeneous(Plan& p, cplx* a, cplx* b, cplx* c, cplx* d, cplx* C1, int bctarget) {
  const Dims dm = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky, *kz = p.d_kz;
  const int has_mean = p.ista == 1;
  const double inv_nz = 1.0 / (double)p.nz;
  SX_EW_LAUNCH(p, k_proj_inh, dm.n, dm, a, b, c, d, C1, kx, ky, kz, bctarget, has_mean, inv_nz);
  return 0;
}
int op_laplace_z(Plan& p, const cplx* C1, const cplx* C2in, cplx* C2, cplx* C3, int bctarget,
                 int bczsta, int bczend) {
  const int ks = bczsta + bctarget, ke = bczend + bctarget;
  const bool ok = (ks == 0 && ke == 0) || (ks == 1 && ke == 1) || (ks == 2 && ke == 2) || (ks == 0 && ke == 2);
  SX_REQUIRE(ok, "Unsupported BC combination in call to laplace_z. Aborting...");
  SX_REQUIRE((bczsta == 0 || bczsta == 2) && (bczend == 0 || bczend == 2),
             "Unsupported BC kind in call to sol_project. Aborting...");
  const Dims dm = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky, *zc = p.d_z;
  const int top = p.nz - p.Cz - 1, has_mean = p.ista == 1;
  const double Lz = p.Lz;
  SX_EW_LAUNCH(p, k_laplace_z, dm.n, dm, C1, C2in, C2, C3, kx, ky, zc, top, bctarget, bczsta, bczend,
               has_mean, Lz);
  return 0;
}
int op_pr_combine(Plan& p, cplx* d, const cplx* C1, const cplx* C2, int bctarget) {
  const size_t n = p.csize();
  const double inv_nz = 1.0 / (double)p.nz;
  SX_EW_LAUNCH(p, k_pr_combine, n, n, d, C1, C2, bctarget, inv_nz);
  return 0;
}
int op_apply_hom(Plan& p, cplx* a, cplx* b, cplx* c, const cplx* C2, const cplx* C3) {
  const Dims dm = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky;
  SX_EW_LAUNCH(p, k_apply_hom, dm.n, dm, a, b, c, C2, C3, kx, ky);
  return 0;
}
int op_noslip(Plan& p, cplx* vx, cplx* vy, const cplx* pr, int o, double vbx0, double vby0, double vbx1,
              double vby1) {
  const Dims dm = dims_of(p);
  const double *kx = p.d_kx, *ky = p.d_ky;
  double tmp = 1.0 / (double)o;
  if (o != p.ord) tmp = (double)(o + 1) * tmp;
  const int top = p.nz - p.Cz - 1, has_mean = p.ista == 1;
  const double sc = (double)p.nx * (double)p.ny;
  const size_t n = 2 * (size_t)p.ny * p.nxl;
  SX_EW_LAUNCH(p, k_noslip, n, dm, vx, vy, pr, kx, ky, top, tmp, has_mean, sc * vbx0, sc * vby0,
               sc * vbx1, sc * vby1);
  return 0;
}

int op_reduce_phys(Plan& p, const cplx* a, const cplx* b, int mode, int row, double scale, double* result) {
  const Dims dm = dims_of(p);
  const int blocks = p.red_blocks, nph = p.nphys(), first = p.ista == 1;
  double* partial = p.d_red;
  cudaStream_t st = p.stream;
  auto kfn = k_reduce_phys;
  if (stage_mark(p, ST_REDUCE)) return 1;
  SX_LAUNCH(kfn, dim3(blocks), dim3(256), 256 * sizeof(double), st, dm, a, b, mode, nph, row, first, scale,
            partial);
  p.launches++;
  SX_KERNEL_CHECK();
  SX_CUDA_CHECK(cudaMemcpyAsync(p.h_red, p.d_red, blocks * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  double s = 0.0;
  for (int i = 0; i < blocks; ++i) s += p.h_red[i];
  if (allreduce_sum(p, &s, 1)) return 1;  // the reference: MPI_REDUCE to rank 0 (pseudospec_hd.f90:631)
  *result = s;
  return 0;
}

}  // namespace sx
