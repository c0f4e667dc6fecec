// Boussinesq and vector-potential MHD: the operators the HD path does not need (advect, vector,
// s_imposebc, a_imposebc_and_project with neumann_reconstruct), the plan-owned device state of
// the two solvers and their Runge-Kutta substeps (include/bouss/bouss_rkstep{1,2}.f90,
// include/mhd/mhd_rkstep{1,2}.f90).  impl=1 composes a substep from the per-operator kernels in
// the reference's order; impl=0 is the fused slab-parallel path (sx_fused.cu).
#include <fstream>

#include "../../include/specter_b200.h"
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
  const size_t cap = 148u * 16u;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

#define SX_EW_LAUNCH(p, kernel, n, ...)                                        \
  do {                                                                         \
    auto kfn = kernel;                                                         \
    cudaStream_t st_ = (p).stream;                                             \
    if (stage_mark((p), ST_EW)) return 1;                                      \
    SX_LAUNCH(kfn, dim3(ew_grid(n)), dim3(256), 0, st_, __VA_ARGS__);          \
    (p).launches++;                                                            \
    SX_KERNEL_CHECK();                                                         \
  } while (0)

// ---- kernels ---------------------------------------------------------------------------
// advect products (pseudospec_phd.f90:66-107): out = (a1 b1 + a2 b2 + a3 b3) / N^2
__global__ void k_dot3_products(size_t n, const double* __restrict__ a1, const double* __restrict__ a2,
                                const double* __restrict__ a3, const double* __restrict__ b1,
                                const double* __restrict__ b2, const double* __restrict__ b3,
                                double* __restrict__ out, double tmp) {
  SX_GRID_STRIDE(idx, n) {
    double s = a1[idx] * b1[idx];
    s += a2[idx] * b2[idx];
    s += a3[idx] * b3[idx];
    out[idx] = s * tmp;
  }
}

// wall rows z=0 and z=Lz of up to three mixed-domain fields set to zero
// (s_constant_z sboundary.f90:122-165; int_conducting_z bboundary.f90:192-236)
__global__ void k_zero_walls(Dims d, cplx* __restrict__ a, cplx* __restrict__ b, cplx* __restrict__ c, int top) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(t, 2 * npen) {
    const size_t idx = (t % npen) * d.nz + (t / npen ? top : 0);
    const cplx z = cmake(0.0, 0.0);
    a[idx] = z;
    if (b) b[idx] = z;
    if (c) c[idx] = z;
  }
}

// conducting_z (bboundary.f90:239-290) and insulating_z (:294-344): wall rows <- 0, then, per wall,
// neumann_reconstruct (fcgram_mod.f90:456-499) with the 2nd-order weights for a,b and the 1st-order weights for c
// (conducting, kind 0), or robin_reconstruct (fcgram_mod.f90:612-635) with the 1st-order weights and the
// coefficient khom = sqrt(kx^2+ky^2) for all three components (vacuum, kind 1)
struct NeuW { double w1[10], w2[10]; };
__global__ void k_magnetic_walls(Dims d, cplx* __restrict__ a, cplx* __restrict__ b, cplx* __restrict__ c,
                                 int top, int dd, NeuW nw, int kind_sta, int kind_end,
                                 const double* __restrict__ kx, const double* __restrict__ ky) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(t, 2 * npen) {
    const bool upper = t / npen;
    const size_t pen = t % npen;
    const size_t base = pen * d.nz;
    const int kind = upper ? kind_end : kind_sta;
    double inv = 1.0;
    if (kind == 1) {
      const double x = kx[pen / d.ny], y = ky[pen % d.ny];
      inv = 1.0 / (sqrt(x * x + y * y) * nw.w1[dd - 1] + 1.0);
    }
    cplx* f[3] = {a, b, c};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double* w = (q < 2 && kind == 0) ? nw.w2 : nw.w1;
      // the prescribed wall datum (wall row) is zero: neu(d)*0 + sum_k neu(k) f(.)
      double sx_ = 0.0, sy_ = 0.0;
      for (int k = 1; k < dd; ++k) {
        const cplx v = upper ? f[q][base + top - dd + k] : f[q][base + dd - k];
        sx_ += w[k - 1] * v.x;
        sy_ += w[k - 1] * v.y;
      }
      f[q][base + (upper ? top : 0)] = kind == 1 ? cmake(sx_ * inv, sy_ * inv) : cmake(sx_, sy_);
    }
  }
}

// wall rows of one or both walls of up to three fields set to zero (int_conducting_z only acts on conducting walls)
__global__ void k_zero_wall_sel(Dims d, cplx* __restrict__ a, cplx* __restrict__ b, int top, int do_sta, int do_end) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(t, 2 * npen) {
    const bool upper = t / npen;
    if (upper ? !do_end : !do_sta) continue;
    const size_t idx = (t % npen) * d.nz + (upper ? top : 0);
    a[idx] = cmake(0.0, 0.0);
    b[idx] = cmake(0.0, 0.0);
  }
}

__global__ void k_sub(size_t n, cplx* __restrict__ a, const cplx* __restrict__ b) {
  SX_GRID_STRIDE(idx, n) a[idx] = csub(a[idx], b[idx]);
}

// bouss_rkstep2.f90:9-18: C6 -= xmom*th ; C8 -= xtemp*vz
__global__ void k_bouss_couple(size_t n, cplx* __restrict__ c6, cplx* __restrict__ c8, const cplx* __restrict__ th,
                               const cplx* __restrict__ vz, double xmom, double xtemp) {
  SX_GRID_STRIDE(idx, n) {
    c6[idx] = caxpy(-xmom, th[idx], c6[idx]);
    c8[idx] = caxpy(-xtemp, vz[idx], c8[idx]);
  }
}

// rotbouss_rkstep2.f90:14-18: the Coriolis (+ buoyancy) terms added to the nonlinear term before the filter,
// as three stand-alone fields:  cx = 2(oy vz - oz vy), cy = 2(oz vx - ox vz), cz = 2(ox vy - oy vx) - xmom th
__global__ void k_rot_couple(size_t n, const cplx* __restrict__ vx, const cplx* __restrict__ vy,
                             const cplx* __restrict__ vz, const cplx* __restrict__ th, double ox, double oy, double oz,
                             double xmom, cplx* __restrict__ cx, cplx* __restrict__ cy, cplx* __restrict__ cz) {
  SX_GRID_STRIDE(idx, n) {
    const cplx X = vx[idx], Y = vy[idx], Z = vz[idx], T = th[idx];
    cx[idx] = cmake(2 * (oy * Z.x - oz * Y.x), 2 * (oy * Z.y - oz * Y.y));
    cy[idx] = cmake(2 * (oz * X.x - ox * Z.x), 2 * (oz * X.y - ox * Z.y));
    cz[idx] = cmake(2 * (ox * Y.x - oy * X.x) - xmom * T.x, 2 * (ox * Y.y - oy * X.y) - xmom * T.y);
  }
}

// mhd_rkstep2.f90:69-74: a = a0 + dt*(-mu*a + emf + m)*rmp   (a holds J on entry)
__global__ void k_rk_axpy_a(size_t n, cplx* __restrict__ a, const cplx* __restrict__ a0, const cplx* __restrict__ emf,
                            const cplx* __restrict__ m, double dt, double mu, double rmp) {
  SX_GRID_STRIDE(idx, n) {
    const cplx J = a[idx], B = a0[idx], E = emf[idx], M = m[idx];
    a[idx] = cmake(B.x + dt * (-mu * J.x + E.x + M.x) * rmp, B.y + dt * (-mu * J.y + E.y + M.y) * rmp);
  }
}

// MHD substep, all three curls in one pass (mhd_rkstep2.f90:6-20 and the prodre call at :29): omega = curl v,
// B = curl A with the uniform field injected at the mean mode (B(1,1,1) = b0 N), J = curl B written over A.  The nine
// curlk calls it replaces move 27 full-field reads / writes; this pass reads 6 and writes 9.  Same expressions as
// k_curlk (sx_kernels_ops.cu): curl_1 = i ky c - i kz b, curl_2 = i kz a - i kx c, curl_3 = i kx b - i ky a.
__device__ __forceinline__ void curl3(double x, double y, double z, cplx a, cplx b, cplx c, cplx& o1, cplx& o2, cplx& o3) {
  o1 = csub(cmake(-y * c.y, y * c.x), cmake(-z * b.y, z * b.x));
  o2 = csub(cmake(-z * a.y, z * a.x), cmake(-x * c.y, x * c.x));
  o3 = csub(cmake(-x * b.y, x * b.x), cmake(-y * a.y, y * a.x));
}
__global__ void k_mhd_curls(Dims d, const cplx* __restrict__ vx, const cplx* __restrict__ vy, const cplx* __restrict__ vz,
                            cplx* __restrict__ ax, cplx* __restrict__ ay, cplx* __restrict__ az, cplx* __restrict__ wx,
                            cplx* __restrict__ wy, cplx* __restrict__ wz, cplx* __restrict__ bx, cplx* __restrict__ by,
                            cplx* __restrict__ bz, const double* __restrict__ kx, const double* __restrict__ ky,
                            const double* __restrict__ kz, int mean, double m0, double m1, double m2) {
  SX_GRID_STRIDE(idx, d.n) {
    const int k = (int)(idx % d.nz);
    const size_t ji = idx / d.nz;
    const int j = (int)(ji % d.ny), i = (int)(ji / d.ny);
    const double x = __ldg(&kx[i]), y = __ldg(&ky[j]), z = __ldg(&kz[k]);
    cplx o1, o2, o3;
    curl3(x, y, z, vx[idx], vy[idx], vz[idx], o1, o2, o3);
    wx[idx] = o1; wy[idx] = o2; wz[idx] = o3;
    curl3(x, y, z, ax[idx], ay[idx], az[idx], o1, o2, o3);
    if (mean && idx == 0) { o1 = cmake(m0, 0.0); o2 = cmake(m1, 0.0); o3 = cmake(m2, 0.0); }
    bx[idx] = o1; by[idx] = o2; bz[idx] = o3;
    cplx j1, j2, j3;
    curl3(x, y, z, o1, o2, o3, j1, j2, j3);
    ax[idx] = j1; ay[idx] = j2; az[idx] = j3;
  }
}

__global__ void k_set_elem(cplx* __restrict__ a, size_t idx, double re, double im) {
  if (blockIdx.x == 0 && threadIdx.x == 0) a[idx] = cmake(re, im);
}

// the theta `hack' (bouss_rkstep2.f90:57-59) in the mixed domain: a 2-D c2r followed by r2c is the
// identity except on the kx = 0 and kx = nx/2 planes, where FFTW's c2r drops the imaginary part of the
// y-transformed line, i.e. a(ky) <- (a(ky) + conj(a(-ky)))/2.  One thread per (z, ky <= ny/2) of a plane.
__global__ void k_hermitian_plane(cplx* __restrict__ plane, int nz, int ny, int nrows) {
  const size_t n = (size_t)(ny / 2 + 1) * nrows;
  SX_GRID_STRIDE(t, n) {
    const int z = (int)(t % nrows), j = (int)(t / nrows);
    const int jm = (ny - j) % ny;
    const cplx A = plane[(size_t)j * nz + z], B = plane[(size_t)jm * nz + z];
    plane[(size_t)j * nz + z] = cmake(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
    plane[(size_t)jm * nz + z] = cmake(0.5 * (A.x + B.x), 0.5 * (B.y - A.y));
  }
}

// ---- small launch helpers ------------------------------------------------------------------
static int op_sub(Plan& p, cplx* a, const cplx* b) {
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_sub, n, n, a, b);
  return 0;
}
int mhd_curls(Plan& p, const cplx* vx, const cplx* vy, const cplx* vz, cplx* ax, cplx* ay, cplx* az, cplx* const* W,
              cplx* const* B, const double* b0) {
  const Dims d = dims_of(p);
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  const double *kx = p.d_kx, *ky = p.d_ky, *kz = p.d_kz;
  const int mean = p.ista == 1 ? 1 : 0;
  const double m0 = (b0 ? b0[0] : 0.0) * N, m1 = (b0 ? b0[1] : 0.0) * N, m2 = (b0 ? b0[2] : 0.0) * N;
  cplx *w0 = W[0], *w1 = W[1], *w2 = W[2], *c0 = B[0], *c1 = B[1], *c2 = B[2];
  SX_EW_LAUNCH(p, k_mhd_curls, d.n, d, vx, vy, vz, ax, ay, az, w0, w1, w2, c0, c1, c2, kx, ky, kz, mean, m0, m1, m2);
  return 0;
}
int op_set_elem(Plan& p, cplx* a, size_t idx, double re, double im) {
  SX_EW_LAUNCH(p, k_set_elem, 1, a, idx, re, im);
  return 0;
}
static int op_zero_walls(Plan& p, cplx* a, cplx* b, cplx* c) {
  const Dims d = dims_of(p);
  const size_t n = 2 * (size_t)p.ny * p.nxl;
  SX_EW_LAUNCH(p, k_zero_walls, n, d, a, b, c, p.nphys() - 1);
  return 0;
}

// load_neumann_tables (fcgram_mod.f90:261-365): neu = Q(d,:) . Qn^T, last entry x (dz/dxp)^ord
int load_neumann(Plan& p) {
  if (!p.h_neu.empty()) return 0;
  SX_REQUIRE(!p.tdir.empty(), "Neumann reconstruction needs the FC-Gram table directory (tdir)");
  const int d = p.oz;
  auto rd = [&](const std::string& path, size_t count, std::vector<double>& out) -> int {
    std::ifstream f(path, std::ios::binary);
    SX_REQUIRE(f.good(), "Could not find table " + path);
    out.resize(count);
    f.read(reinterpret_cast<char*>(out.data()), (std::streamsize)(count * sizeof(double)));
    SX_REQUIRE((size_t)f.gcount() == count * sizeof(double), "FC-Gram table too short: " + path);
    return 0;
  };
  std::vector<double> Q;
  if (rd(p.tdir + "/Q" + std::to_string(d) + ".dat", (size_t)d * d, Q)) return 1;
  for (int ord = 1; ord <= 2; ++ord) {
    std::vector<double> raw;
    if (rd(p.tdir + "/Q" + std::to_string(ord) + "n" + std::to_string(d) + ".dat", (size_t)d * d + 1, raw)) return 1;
    const double dxp = raw[0];
    std::vector<double> neu(d);
    for (int k = 0; k < d; ++k) {
      double s = 0.0;
      for (int j = 0; j < d; ++j) s += Q[(size_t)j * d + (d - 1)] * raw[1 + (size_t)j * d + k];  // Q(d,j) Qn(k,j)
      neu[k] = s;
    }
    neu[d - 1] *= ord == 1 ? p.dz / dxp : (p.dz / dxp) * (p.dz / dxp);
    (ord == 1 ? p.h_neu : p.h_neu2) = neu;
  }
  return 0;
}

// ---- pseudo: advect / vector ------------------------------------------------------------------
// advect (pseudospec_phd.f90:23-113): e = FFT3[ A . grad(d) ] / N^2
int advect(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* d, cplx* e) {
  double* r[7];
  for (int i = 0; i < 7; ++i) if (plan_rwork(p, i, &r[i])) return 1;
  cplx* t;
  if (plan_cwork(p, 1, &t)) return 1;
  const cplx* comp[3] = {a, b, c};
  for (int dir = 1; dir <= 3; ++dir) {
    if (fft3d_c2r(p, comp[dir - 1], r[dir - 1])) return 1;
    if (op_derivk(p, d, t, dir) || fft3d_c2r(p, t, r[2 + dir])) return 1;
  }
  const int nzp = p.pkend - p.ksta + 1;
  if (nzp > 0) {
    const size_t n = (size_t)p.nx * p.ny * nzp;
    const double N = (double)p.nx * (double)p.ny * (double)p.nz;
    SX_EW_LAUNCH(p, k_dot3_products, n, n, r[0], r[1], r[2], r[3], r[4], r[5], r[6], 1.0 / (N * N));
  }
  return fft3d_r2c(p, r[6], e);
}

// vector (pseudospec_mhd.f90:22-105): (x,y,z) = FFT3[ A x B ] / N^2
int vector(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* d, const cplx* e, const cplx* f,
           cplx* x, cplx* y, cplx* z) {
  double* r[9];
  for (int i = 0; i < 9; ++i) if (plan_rwork(p, i, &r[i])) return 1;
  const cplx* in[6] = {a, b, c, d, e, f};
  for (int i = 0; i < 6; ++i) if (fft3d_c2r(p, in[i], r[i])) return 1;
  if (op_cross_products(p, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8])) return 1;
  return fft3d_r2c(p, r[6], x) || fft3d_r2c(p, r[7], y) || fft3d_r2c(p, r[8], z);
}

// ---- boundary: s_imposebc / a_imposebc_and_project -----------------------------------------------
// s_imposebc (sboundary.f90:67-119) with `constant' walls
int s_imposebc(Plan& p, cplx* th) {
  SX_REQUIRE(p.Cz > 0, "scalar wall BCs need a non-periodic z direction (Cz > 0)");
  if (fft1d_z_bwd(p, th, th, 1.0 / (double)p.nz)) return 1;   // goto_domain_w_boundaries
  if (op_zero_walls(p, th, nullptr, nullptr)) return 1;        // s_constant_z
  return fft1d_z_fwd(p, th);                                   // goto_3d_fourier
}

// the theta `hack' (bouss_rkstep2.f90:57-59): th <- FFT3(IFFT3(th)/N), all pencil-local but for the
// Hermitian pairing on the two self-conjugate kx planes
int theta_roundtrip(Plan& p, cplx* th, cplx* out) {
  if (fft1d_z_bwd(p, th, th, 1.0 / (double)p.nz)) return 1;
  const size_t plane = (size_t)p.ny * p.nz;
  const size_t n = (size_t)(p.ny / 2 + 1) * p.nz;
  if (p.ista == 1) SX_EW_LAUNCH(p, k_hermitian_plane, n, th, p.nz, p.ny, p.nz);
  if (p.iend == p.nxh) SX_EW_LAUNCH(p, k_hermitian_plane, n, th + (size_t)(p.nxl - 1) * plane, p.nz, p.ny, p.nz);
  return launch_zfft(p, th, out, (long)p.ny * p.nxl, -1, true, 1.0, 1.0);
}

// a_imposebc_and_project (bboundary.f90:100-189); wall kinds from sx_setup_bc(plan, "b", ...): 0 conducting
// (default), 1 vacuum.  The combinations laplace_z refuses (vacuum bottom under a conducting top) fail as there.
int a_imposebc_and_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph) {
  SX_REQUIRE(p.Cz > 0, "vector-potential wall BCs need a non-periodic z direction (Cz > 0)");
  const int ks = p.b_bczsta, ke = p.b_bczend;
  SX_REQUIRE((ks == 0 || ks == 1) && (ke == 0 || ke == 1), "Unsupported boundary conditions in Z direction. Aborting...");
  SX_REQUIRE(!(ks == 1 && ke == 0), "Unsupported BC combination in call to laplace_z. Aborting...");
  if (load_neumann(p)) return 1;
  const double inv_nz = 1.0 / (double)p.nz;
  const Dims d = dims_of(p);
  const size_t n = 2 * (size_t)p.ny * p.nxl;
  const int top = p.nphys() - 1;
  if (p.ista == 1 && op_set_elem(p, az, 0, 0.0, 0.0)) return 1;           // bboundary.f90:147-149
  if (ks == 0 || ke == 0) {
    if (fft1d_z_bwd(p, ax, ax, inv_nz) || fft1d_z_bwd(p, ay, ay, inv_nz)) return 1;
    SX_EW_LAUNCH(p, k_zero_wall_sel, n, d, ax, ay, top, ks == 0, ke == 0);  // int_conducting_z
    if (fft1d_z_fwd(p, ax) || fft1d_z_fwd(p, ay)) return 1;
  }
  if (sol_project(p, ax, ay, az, ph, 0, 2 * ks, 2 * ke)) return 1;
  if (fft1d_z_bwd(p, ax, ax, inv_nz) || fft1d_z_bwd(p, ay, ay, inv_nz) || fft1d_z_bwd(p, az, az, inv_nz)) return 1;
  NeuW nw;
  for (int k = 0; k < 10; ++k) {
    nw.w1[k] = k < p.oz ? p.h_neu[k] : 0.0;
    nw.w2[k] = k < p.oz ? p.h_neu2[k] : 0.0;
  }
  const double *kx = p.d_kx, *ky = p.d_ky;
  SX_EW_LAUNCH(p, k_magnetic_walls, n, d, ax, ay, az, top, p.oz, nw, ks, ke, kx, ky);  // conducting_z / insulating_z
  return fft1d_z_fwd(p, ax) || fft1d_z_fwd(p, ay) || fft1d_z_fwd(p, az);
}

// variance (pseudospec_phd.f90:116-196), kin = 1: <a^2>, kin = 0: <|k^2 a|^2>
int variance(Plan& p, const cplx* a, int kin, double* out) {
  SX_REQUIRE(kin == 0 || kin == 1, "variance: kin must be 0 or 1");
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  const double tmp = 1.0 / (N * N) / (double)(p.nz - p.Cz);
  cplx *w, *t;
  if (plan_cwork(p, 2, &w)) return 1;
  if (kin == 0) {
    if (plan_cwork(p, 3, &t) || op_laplak(p, a, t)) return 1;
    a = t;
  }
  if (fft1d_z_bwd(p, a, w, 1.0)) return 1;
  return op_reduce_phys(p, w, nullptr, 0, -1, tmp, out);
}

// ---- solver states ---------------------------------------------------------------------------------
// BOUSS: 0..2 v, 3 pr, 4..6 f, 7..9 C1..C3, 10 th, 11 fs, 12 C7, 13 scratch (fused path)
// MHD:   0..2 v, 3 pr, 4..6 f, 7..9 C1..C3, 10..12 a, 13 ph, 14..16 m, 17..19 C9..C11
struct SolverState {
  std::vector<cplx*> f;
};

static int state_get(Plan& p, SolverState** slot, int nfields, SolverState** out) {
  if (!*slot) {
    SolverState* s = new SolverState();
    *slot = s;
    s->f.assign(nfields, nullptr);
    for (int i = 0; i < nfields; ++i) {
      SX_CUDA_CHECK(cudaMalloc((void**)&s->f[i], p.csize() * sizeof(cplx)));
      SX_CUDA_CHECK(cudaMemsetAsync(s->f[i], 0, p.csize() * sizeof(cplx), p.stream));
    }
  }
  *out = *slot;
  return 0;
}
static void state_free(SolverState** slot) {
  if (*slot) {
    for (cplx* q : (*slot)->f) if (q) cudaFree(q);
    delete *slot;
    *slot = nullptr;
  }
}
int solver_states_free(Plan& p) {
  state_free(&p.bouss);
  state_free(&p.mhd);
  state_free(&p.mhdbouss);
  return 0;
}
// MHDBOUSS: the MHD slots, then 20 th, 21 fs, 22 C7
constexpr int kBoussFields = 14, kMhdFields = 20, kMhdBoussFields = 23;

static int put_fields(Plan& p, SolverState& s, const double* const* h, const int* which, int n) {
  const size_t bytes = p.csize() * sizeof(cplx);
  for (int i = 0; i < n; ++i)
    if (h[i]) SX_CUDA_CHECK(cudaMemcpyAsync(s.f[which[i]], h[i], bytes, cudaMemcpyHostToDevice, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  return 0;
}
static int get_fields(Plan& p, SolverState& s, double* const* h, const int* which, int n) {
  const size_t bytes = p.csize() * sizeof(cplx);
  for (int i = 0; i < n; ++i)
    if (h[i]) SX_CUDA_CHECK(cudaMemcpyAsync(h[i], s.f[which[i]], bytes, cudaMemcpyDeviceToHost, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  return 0;
}
static int copy_field(Plan& p, cplx* dst, const cplx* src) {
  SX_CUDA_CHECK(cudaMemcpyAsync(dst, src, p.csize() * sizeof(cplx), cudaMemcpyDeviceToDevice, p.stream));
  return 0;
}

// ---- substeps composed from stand-alone operators ---------------------------------------------------
// bouss_rkstep2.f90:3-59
static int bouss_rkstep2_modular(Plan& p, SolverState& s, int o, double dt, double nu, double kappa, double xmom,
                                 double xtemp, const double* zs, const double* ze) {
  cplx *c4, *c5, *c6, *c8;
  if (plan_cwork(p, 6, &c4) || plan_cwork(p, 7, &c5) || plan_cwork(p, 8, &c6) || plan_cwork(p, 9, &c8)) return 1;
  const double rmp = 1.0 / (double)o;
  cplx** f = s.f.data();
  if (gradre(p, f[0], f[1], f[2], c4, c5, c6)) return 1;
  if (advect(p, f[0], f[1], f[2], f[10], c8)) return 1;
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_bouss_couple, n, n, c6, c8, f[10], f[2], xmom, xtemp);
  if (op_fc_filter(p, c4) || op_fc_filter(p, c5) || op_fc_filter(p, c6) || op_fc_filter(p, c8)) return 1;
  cplx* nl[3] = {c4, c5, c6};
  for (int q = 0; q < 3; ++q) {
    if (op_laplak(p, f[q], f[q])) return 1;
    if (op_rk_axpy(p, f[q], f[7 + q], nl[q], f[4 + q], dt, nu, rmp)) return 1;
  }
  if (op_laplak(p, f[10], f[10]) || op_rk_axpy(p, f[10], f[12], c8, f[11], dt, kappa, rmp)) return 1;
  if (v_imposebc_and_project(p, f[0], f[1], f[2], f[3], o, zs, ze)) return 1;
  if (s_imposebc(p, f[10]) || op_fc_filter(p, f[10])) return 1;
  return theta_roundtrip(p, f[10], f[10]);
}

// mhd_rkstep2.f90:3-84
static int mhd_rkstep2_modular(Plan& p, SolverState& s, int o, double dt, double nu, double mu, const double* b0) {
  cplx *c4, *c5, *c6, *c12, *c13, *c14, *c15, *c16, *c17;
  if (plan_cwork(p, 6, &c4) || plan_cwork(p, 7, &c5) || plan_cwork(p, 8, &c6)) return 1;
  if (plan_cwork(p, 9, &c12) || plan_cwork(p, 10, &c13) || plan_cwork(p, 11, &c14)) return 1;
  if (plan_cwork(p, 12, &c15) || plan_cwork(p, 13, &c16) || plan_cwork(p, 14, &c17)) return 1;
  const double rmp = 1.0 / (double)o;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  cplx** f = s.f.data();
  cplx *ax = f[10], *ay = f[11], *az = f[12];
  if (op_curlk(p, ay, az, c12, 1) || op_curlk(p, ax, az, c13, 2) || op_curlk(p, ax, ay, c14, 3)) return 1;
  if (p.ista == 1) {
    if (op_set_elem(p, c12, 0, (b0 ? b0[0] : 0.0) * N, 0.0) || op_set_elem(p, c13, 0, (b0 ? b0[1] : 0.0) * N, 0.0) ||
        op_set_elem(p, c14, 0, (b0 ? b0[2] : 0.0) * N, 0.0)) return 1;
  }
  if (op_curlk(p, c13, c14, ax, 1) || op_curlk(p, c12, c14, ay, 2) || op_curlk(p, c12, c13, az, 3)) return 1;
  if (prodre(p, f[0], f[1], f[2], c4, c5, c6)) return 1;
  if (vector(p, ax, ay, az, c12, c13, c14, c15, c16, c17)) return 1;
  if (op_sub(p, c4, c15) || op_sub(p, c5, c16) || op_sub(p, c6, c17)) return 1;
  if (op_fc_filter(p, c4) || op_fc_filter(p, c5) || op_fc_filter(p, c6)) return 1;
  if (vector(p, f[0], f[1], f[2], c12, c13, c14, c15, c16, c17)) return 1;
  if (op_fc_filter(p, c15) || op_fc_filter(p, c16) || op_fc_filter(p, c17)) return 1;
  cplx* nl[3] = {c4, c5, c6};
  cplx* emf[3] = {c15, c16, c17};
  const size_t n = p.csize();
  for (int q = 0; q < 3; ++q) {
    if (op_laplak(p, f[q], f[q])) return 1;
    if (op_rk_axpy(p, f[q], f[7 + q], nl[q], f[4 + q], dt, nu, rmp)) return 1;
    SX_EW_LAUNCH(p, k_rk_axpy_a, n, n, f[10 + q], f[17 + q], emf[q], f[14 + q], dt, mu, rmp);
  }
  if (v_imposebc_and_project(p, f[0], f[1], f[2], f[3], o, nullptr, nullptr)) return 1;
  return a_imposebc_and_project(p, ax, ay, az, f[13]);
}

// rotbouss_rkstep2.f90:3-56 (ends with s_imposebc: no theta filter / round trip)
int rot_couple(Plan& p, cplx* const* f, const double* om, double xmom, cplx* cx, cplx* cy, cplx* cz) {
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_rot_couple, n, n, f[0], f[1], f[2], f[10], om[0], om[1], om[2], xmom, cx, cy, cz);
  return 0;
}
static int rotbouss_rkstep2_modular(Plan& p, SolverState& s, int o, double dt, double nu, double kappa, double xmom,
                                    double xtemp, const double* om, const double* zs, const double* ze) {
  cplx *c4, *c5, *c6, *c8, *r[3];
  if (plan_cwork(p, 6, &c4) || plan_cwork(p, 7, &c5) || plan_cwork(p, 8, &c6) || plan_cwork(p, 9, &c8)) return 1;
  for (int q = 0; q < 3; ++q) if (plan_cwork(p, 10 + q, &r[q])) return 1;
  const double rmp = 1.0 / (double)o;
  cplx** f = s.f.data();
  if (gradre(p, f[0], f[1], f[2], c4, c5, c6)) return 1;
  if (advect(p, f[0], f[1], f[2], f[10], c8)) return 1;
  const size_t n = p.csize();
  if (rot_couple(p, f, om, xmom, r[0], r[1], r[2])) return 1;
  cplx* nl[3] = {c4, c5, c6};
  for (int q = 0; q < 3; ++q) if (op_add(p, nl[q], r[q])) return 1;
  SX_EW_LAUNCH(p, k_bouss_couple, n, n, c6, c8, f[10], f[2], 0.0, xtemp);   // heat current only
  if (op_fc_filter(p, c4) || op_fc_filter(p, c5) || op_fc_filter(p, c6) || op_fc_filter(p, c8)) return 1;
  for (int q = 0; q < 3; ++q) {
    if (op_laplak(p, f[q], f[q])) return 1;
    if (op_rk_axpy(p, f[q], f[7 + q], nl[q], f[4 + q], dt, nu, rmp)) return 1;
  }
  if (op_laplak(p, f[10], f[10]) || op_rk_axpy(p, f[10], f[12], c8, f[11], dt, kappa, rmp)) return 1;
  if (v_imposebc_and_project(p, f[0], f[1], f[2], f[3], o, zs, ze)) return 1;
  return s_imposebc(p, f[10]);
}

// mhdbouss_rkstep2.f90:3-106, composed from the stand-alone operators in the reference's order.
// State: the MHD slots (0..19), 20 th, 21 fs, 22 C7.
static int mhdbouss_rkstep2_modular(Plan& p, SolverState& s, int o, double dt, double nu, double mu, double kappa,
                                    double xmom, double xtemp, const double* b0) {
  cplx *c4, *c5, *c6, *c8, *c12, *c13, *c14, *c15, *c16, *c17;
  if (plan_cwork(p, 6, &c4) || plan_cwork(p, 7, &c5) || plan_cwork(p, 8, &c6)) return 1;
  if (plan_cwork(p, 9, &c12) || plan_cwork(p, 10, &c13) || plan_cwork(p, 11, &c14)) return 1;
  if (plan_cwork(p, 12, &c15) || plan_cwork(p, 13, &c16) || plan_cwork(p, 14, &c17) || plan_cwork(p, 15, &c8)) return 1;
  const double rmp = 1.0 / (double)o;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  cplx** f = s.f.data();
  cplx *ax = f[10], *ay = f[11], *az = f[12], *th = f[20];
  if (op_curlk(p, ay, az, c12, 1) || op_curlk(p, ax, az, c13, 2) || op_curlk(p, ax, ay, c14, 3)) return 1;
  if (p.ista == 1) {
    if (op_set_elem(p, c12, 0, (b0 ? b0[0] : 0.0) * N, 0.0) || op_set_elem(p, c13, 0, (b0 ? b0[1] : 0.0) * N, 0.0) ||
        op_set_elem(p, c14, 0, (b0 ? b0[2] : 0.0) * N, 0.0)) return 1;
  }
  if (op_curlk(p, c13, c14, ax, 1) || op_curlk(p, c12, c14, ay, 2) || op_curlk(p, c12, c13, az, 3)) return 1;
  if (prodre(p, f[0], f[1], f[2], c4, c5, c6)) return 1;
  if (advect(p, f[0], f[1], f[2], th, c8)) return 1;
  if (vector(p, ax, ay, az, c12, c13, c14, c15, c16, c17)) return 1;
  if (op_sub(p, c4, c15) || op_sub(p, c5, c16) || op_sub(p, c6, c17)) return 1;
  const size_t n = p.csize();
  SX_EW_LAUNCH(p, k_bouss_couple, n, n, c6, c8, th, f[2], xmom, xtemp);
  if (op_fc_filter(p, c4) || op_fc_filter(p, c5) || op_fc_filter(p, c6) || op_fc_filter(p, c8)) return 1;
  if (vector(p, f[0], f[1], f[2], c12, c13, c14, c15, c16, c17)) return 1;
  if (op_fc_filter(p, c15) || op_fc_filter(p, c16) || op_fc_filter(p, c17)) return 1;
  cplx* nl[3] = {c4, c5, c6};
  cplx* emf[3] = {c15, c16, c17};
  for (int q = 0; q < 3; ++q) {
    if (op_laplak(p, f[q], f[q])) return 1;
    if (op_rk_axpy(p, f[q], f[7 + q], nl[q], f[4 + q], dt, nu, rmp)) return 1;
    SX_EW_LAUNCH(p, k_rk_axpy_a, n, n, f[10 + q], f[17 + q], emf[q], f[14 + q], dt, mu, rmp);
  }
  if (op_laplak(p, th, th) || op_rk_axpy(p, th, f[22], c8, f[21], dt, kappa, rmp)) return 1;
  if (v_imposebc_and_project(p, f[0], f[1], f[2], f[3], o, nullptr, nullptr)) return 1;
  if (a_imposebc_and_project(p, ax, ay, az, f[13])) return 1;
  if (s_imposebc(p, th)) return 1;
  return theta_roundtrip(p, th, th);
}

int bouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                        const double* zs, const double* ze);
int mhdbouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double mu, double kappa, double xmom,
                           double xtemp, const double* b0);
int rotbouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double kappa, double xmom,
                           double xtemp, const double* om, const double* zs, const double* ze);
int mhd_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double mu, const double* b0);

}  // namespace sx

using namespace sx;
#define SX_PLAN(pl) \
  if (!(pl)) { sx::set_error("[ERROR] null plan"); return 1; } \
  sx::Plan& p = (pl)->p
static inline cplx* C(double* a) { return reinterpret_cast<cplx*>(a); }
static inline const cplx* C(const double* a) { return reinterpret_cast<const cplx*>(a); }

extern "C" {

int sx_advect(sx_plan* plan, const double* a, const double* b, const double* c, const double* d, double* e) {
  SX_PLAN(plan);
  return advect(p, C(a), C(b), C(c), C(d), C(e));
}
int sx_vector(sx_plan* plan, const double* a, const double* b, const double* c, const double* d, const double* e,
              const double* f, double* x, double* y, double* z) {
  SX_PLAN(plan);
  return vector(p, C(a), C(b), C(c), C(d), C(e), C(f), C(x), C(y), C(z));
}
int sx_variance(sx_plan* plan, const double* a, int kin, double* out) { SX_PLAN(plan); return variance(p, C(a), kin, out); }
int sx_s_imposebc(sx_plan* plan, double* th) { SX_PLAN(plan); return s_imposebc(p, C(th)); }
int sx_a_imposebc_and_project(sx_plan* plan, double* ax, double* ay, double* az, double* ph) {
  SX_PLAN(plan);
  return a_imposebc_and_project(p, C(ax), C(ay), C(az), C(ph));
}

// ---- BOUSS ---------------------------------------------------------------------------------------
int sx_bouss_put_state(sx_plan* plan, const double* vx, const double* vy, const double* vz, const double* pr,
                       const double* th, const double* fx, const double* fy, const double* fz, const double* fs) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  const double* h[9] = {vx, vy, vz, pr, th, fx, fy, fz, fs};
  const int which[9] = {0, 1, 2, 3, 10, 4, 5, 6, 11};
  return put_fields(p, *s, h, which, 9);
}
int sx_bouss_get_state(sx_plan* plan, double* vx, double* vy, double* vz, double* pr, double* th) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  double* h[5] = {vx, vy, vz, pr, th};
  const int which[5] = {0, 1, 2, 3, 10};
  return get_fields(p, *s, h, which, 5);
}
int sx_bouss_state_ptr(sx_plan* plan, int which, double** dptr) {
  SX_PLAN(plan);
  SX_REQUIRE(which >= 0 && which < kBoussFields && dptr, "sx_bouss_state_ptr: which must be 0..13");
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  *dptr = reinterpret_cast<double*>(s->f[which]);
  return 0;
}
/* bouss_rkstep1.f90:4-7 */
int sx_bouss_rkstep1(sx_plan* plan) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  for (int q = 0; q < 3; ++q) if (copy_field(p, s->f[7 + q], s->f[q])) return 1;
  return copy_field(p, s->f[12], s->f[10]);
}
int sx_bouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                     const double v_zsta[2], const double v_zend[2], int impl) {
  SX_PLAN(plan);
  SX_REQUIRE(o >= 1 && o <= p.ord, "sx_bouss_rkstep2: substep index o must be in 1..ord");
  SX_REQUIRE(p.Cz > 0, "wall BCs need a non-periodic z direction (Cz > 0)");
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  if (impl == 1) return bouss_rkstep2_modular(p, *s, o, dt, nu, kappa, xmom, xtemp, v_zsta, v_zend);
  SX_REQUIRE(impl == 0, "sx_bouss_rkstep2: impl must be 0 (fused) or 1 (per-operator)");
  return bouss_rkstep2_fused(p, s->f.data(), o, dt, nu, kappa, xmom, xtemp, v_zsta, v_zend);
}

/* rotbouss_rkstep2.f90:3-56 on the BOUSS state (rotbouss_rkstep1.f90 = bouss_rkstep1.f90) */
int sx_rotbouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                        const double omega[3], const double v_zsta[2], const double v_zend[2], int impl) {
  SX_PLAN(plan);
  SX_REQUIRE(o >= 1 && o <= p.ord, "sx_rotbouss_rkstep2: substep index o must be in 1..ord");
  SX_REQUIRE(p.Cz > 0, "wall BCs need a non-periodic z direction (Cz > 0)");
  SX_REQUIRE(omega != nullptr, "sx_rotbouss_rkstep2: omega[3] is required");
  SolverState* s;
  if (state_get(p, &p.bouss, kBoussFields, &s)) return 1;
  if (impl == 1) return rotbouss_rkstep2_modular(p, *s, o, dt, nu, kappa, xmom, xtemp, omega, v_zsta, v_zend);
  SX_REQUIRE(impl == 0, "sx_rotbouss_rkstep2: impl must be 0 (fused) or 1 (per-operator)");
  return rotbouss_rkstep2_fused(p, s->f.data(), o, dt, nu, kappa, xmom, xtemp, omega, v_zsta, v_zend);
}

// ---- MHD -----------------------------------------------------------------------------------------
int sx_mhd_put_state(sx_plan* plan, const double* vx, const double* vy, const double* vz, const double* pr,
                     const double* ax, const double* ay, const double* az, const double* fx, const double* fy,
                     const double* fz, const double* mx, const double* my, const double* mz) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhd, kMhdFields, &s)) return 1;
  const double* h[13] = {vx, vy, vz, pr, ax, ay, az, fx, fy, fz, mx, my, mz};
  const int which[13] = {0, 1, 2, 3, 10, 11, 12, 4, 5, 6, 14, 15, 16};
  return put_fields(p, *s, h, which, 13);
}
int sx_mhd_get_state(sx_plan* plan, double* vx, double* vy, double* vz, double* pr, double* ax, double* ay,
                     double* az, double* ph) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhd, kMhdFields, &s)) return 1;
  double* h[8] = {vx, vy, vz, pr, ax, ay, az, ph};
  const int which[8] = {0, 1, 2, 3, 10, 11, 12, 13};
  return get_fields(p, *s, h, which, 8);
}
int sx_mhd_state_ptr(sx_plan* plan, int which, double** dptr) {
  SX_PLAN(plan);
  SX_REQUIRE(which >= 0 && which < kMhdFields && dptr, "sx_mhd_state_ptr: which must be 0..19");
  SolverState* s;
  if (state_get(p, &p.mhd, kMhdFields, &s)) return 1;
  *dptr = reinterpret_cast<double*>(s->f[which]);
  return 0;
}
/* mhd_rkstep1.f90:4-9 */
int sx_mhd_rkstep1(sx_plan* plan) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhd, kMhdFields, &s)) return 1;
  for (int q = 0; q < 3; ++q)
    if (copy_field(p, s->f[7 + q], s->f[q]) || copy_field(p, s->f[17 + q], s->f[10 + q])) return 1;
  return 0;
}
int sx_mhd_rkstep2(sx_plan* plan, int o, double dt, double nu, double mu, const double b0[3], int impl) {
  SX_PLAN(plan);
  SX_REQUIRE(o >= 1 && o <= p.ord, "sx_mhd_rkstep2: substep index o must be in 1..ord");
  SX_REQUIRE(p.Cz > 0, "wall BCs need a non-periodic z direction (Cz > 0)");
  SolverState* s;
  if (state_get(p, &p.mhd, kMhdFields, &s)) return 1;
  if (impl == 1) return mhd_rkstep2_modular(p, *s, o, dt, nu, mu, b0);
  SX_REQUIRE(impl == 0, "sx_mhd_rkstep2: impl must be 0 (fused) or 1 (per-operator)");
  return mhd_rkstep2_fused(p, s->f.data(), o, dt, nu, mu, b0);
}

// ---- MHDBOUSS ------------------------------------------------------------------------------------
int sx_mhdbouss_put_state(sx_plan* plan, const double* vx, const double* vy, const double* vz, const double* pr,
                          const double* ax, const double* ay, const double* az, const double* th, const double* fx,
                          const double* fy, const double* fz, const double* mx, const double* my, const double* mz,
                          const double* fs) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhdbouss, kMhdBoussFields, &s)) return 1;
  const double* h[15] = {vx, vy, vz, pr, ax, ay, az, th, fx, fy, fz, mx, my, mz, fs};
  const int which[15] = {0, 1, 2, 3, 10, 11, 12, 20, 4, 5, 6, 14, 15, 16, 21};
  return put_fields(p, *s, h, which, 15);
}
int sx_mhdbouss_get_state(sx_plan* plan, double* vx, double* vy, double* vz, double* pr, double* ax, double* ay,
                          double* az, double* ph, double* th) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhdbouss, kMhdBoussFields, &s)) return 1;
  double* h[9] = {vx, vy, vz, pr, ax, ay, az, ph, th};
  const int which[9] = {0, 1, 2, 3, 10, 11, 12, 13, 20};
  return get_fields(p, *s, h, which, 9);
}
int sx_mhdbouss_state_ptr(sx_plan* plan, int which, double** dptr) {
  SX_PLAN(plan);
  SX_REQUIRE(which >= 0 && which < kMhdBoussFields && dptr, "sx_mhdbouss_state_ptr: which must be 0..22");
  SolverState* s;
  if (state_get(p, &p.mhdbouss, kMhdBoussFields, &s)) return 1;
  *dptr = reinterpret_cast<double*>(s->f[which]);
  return 0;
}
/* mhdbouss_rkstep1.f90:6-12 */
int sx_mhdbouss_rkstep1(sx_plan* plan) {
  SX_PLAN(plan);
  SolverState* s;
  if (state_get(p, &p.mhdbouss, kMhdBoussFields, &s)) return 1;
  for (int q = 0; q < 3; ++q)
    if (copy_field(p, s->f[7 + q], s->f[q]) || copy_field(p, s->f[17 + q], s->f[10 + q])) return 1;
  return copy_field(p, s->f[22], s->f[20]);
}
/* mhdbouss_rkstep2.f90:3-106 */
int sx_mhdbouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double mu, double kappa, double xmom,
                        double xtemp, const double b0[3], int impl) {
  SX_PLAN(plan);
  SX_REQUIRE(o >= 1 && o <= p.ord, "sx_mhdbouss_rkstep2: substep index o must be in 1..ord");
  SX_REQUIRE(p.Cz > 0, "wall BCs need a non-periodic z direction (Cz > 0)");
  SolverState* s;
  if (state_get(p, &p.mhdbouss, kMhdBoussFields, &s)) return 1;
  if (impl == 1) return mhdbouss_rkstep2_modular(p, *s, o, dt, nu, mu, kappa, xmom, xtemp, b0);
  SX_REQUIRE(impl == 0, "sx_mhdbouss_rkstep2: impl must be 0 (fused) or 1 (per-operator)");
  return mhdbouss_rkstep2_fused(p, s->f.data(), o, dt, nu, mu, kappa, xmom, xtemp, b0);
}

}  // extern "C"
