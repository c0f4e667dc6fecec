// Wall boundary-condition kinds (setup_bc), the stand-alone wall reconstructions of the fcgram module
// (neumann_reconstruct, robin_reconstruct) and the diagnostics of the solvers' *_global.f90 blocks that the HD
// path does not need: helicity, product, pscheck, maxabs, mhdcheck, robcheck, bdiagnostic, sdiagnostic.
// Everything here runs every `cstep` steps or at set-up, never inside a substep.
#include <algorithm>
#include <cctype>
#include <cstring>

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

// ---- kernels ---------------------------------------------------------------------------------------
struct RecW { double w[10]; };
// neumann_reconstruct (fcgram_mod.f90:456-499) / robin_reconstruct (:612-635), z branches, one thread per
// (ky,kx) pencil: wall <- [ neu(d) g + sum_k neu(k) f(neighbour k) ] / (a neu(d) + 1), g = the wall row on entry.
// coef == nullptr and robin: a = khom; !robin: a = 0.
__global__ void k_wall_reconstruct(Dims d, cplx* __restrict__ f, int upper, int top, int dd, RecW rw, int robin,
                                   const double* __restrict__ coef, const double* __restrict__ kx,
                                   const double* __restrict__ ky) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(pen, npen) {
    const size_t base = pen * d.nz;
    const size_t wall = base + (upper ? top : 0);
    const cplx g = f[wall];
    double sx_ = rw.w[dd - 1] * g.x, sy_ = rw.w[dd - 1] * g.y;
    for (int k = 1; k < dd; ++k) {
      const cplx v = upper ? f[base + top - dd + k] : f[base + dd - k];
      sx_ += rw.w[k - 1] * v.x;
      sy_ += rw.w[k - 1] * v.y;
    }
    if (robin) {
      double a;
      if (coef) a = coef[pen];
      else { const double x = kx[pen / d.ny], y = ky[pen % d.ny]; a = sqrt(x * x + y * y); }
      const double den = a * rw.w[dd - 1] + 1.0;
      sx_ /= den;
      sy_ /= den;
    }
    f[wall] = cmake(sx_, sy_);
  }
}

// robcheck (bboundary.f90:475-483): the wall rows of C2 (= IFFT_z of i kz a) are replaced by the residual of the
// vacuum condition, -C2 + khom C1 at z=0 and C2 + khom C1 at z=Lz
__global__ void k_robin_residual(Dims d, const cplx* __restrict__ C1, cplx* __restrict__ C2, int top,
                                 const double* __restrict__ kx, const double* __restrict__ ky) {
  const size_t npen = (size_t)d.ny * d.nxl;
  SX_GRID_STRIDE(t, 2 * npen) {
    const bool upper = t / npen;
    const size_t pen = t % npen;
    const double x = kx[pen / d.ny], y = ky[pen % d.ny];
    const double kh = sqrt(x * x + y * y);
    const size_t idx = pen * d.nz + (upper ? top : 0);
    const cplx a = C1[idx], da = C2[idx];
    const double sg = upper ? 1.0 : -1.0;
    C2[idx] = cmake(sg * da.x + kh * a.x, sg * da.y + kh * a.y);
  }
}

// maxabs (pseudospec_hd.f90:1063-1071): block maxima of r1^2 + r2^2 + r3^2 over the first n points
__global__ void k_max_norm3(size_t n, const double* __restrict__ r1, const double* __restrict__ r2,
                            const double* __restrict__ r3, double* __restrict__ partial) {
  SX_DYN_SMEM(double, sh);
  double m = 0.0;
  SX_GRID_STRIDE(idx, n) {
    const double a = r1[idx], b = r2[idx], c = r3[idx];
    m = fmax(m, a * a + b * b + c * c);
  }
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---- reconstructions ----------------------------------------------------------------------------------
static int wall_reconstruct(Plan& p, cplx* f, int boun, int order, bool robin, const double* coef) {
  SX_REQUIRE(p.Cz > 0, "wall reconstruction needs a non-periodic z direction (Cz > 0)");
  if (robin) SX_REQUIRE(boun == 5 || boun == 6, "Robin reconstruction not performed. Wrong boundary specified. Aborting...");
  else SX_REQUIRE(boun == 5 || boun == 6, "Neumann reconstruction not performed. Wrong boundary specified. Aborting...");
  SX_REQUIRE(order == 1 || order == 2, "neumann_reconstruct: order must be 1 or 2");
  if (load_neumann(p)) return 1;
  RecW rw;
  const std::vector<double>& w = order == 1 ? p.h_neu : p.h_neu2;
  for (int k = 0; k < 10; ++k) rw.w[k] = k < p.oz ? w[k] : 0.0;
  const Dims d = dims_of(p);
  const size_t npen = (size_t)p.ny * p.nxl;
  const double *kx = p.d_kx, *ky = p.d_ky;
  SX_EW_LAUNCH(p, k_wall_reconstruct, npen, d, f, boun == 6 ? 1 : 0, p.nphys() - 1, p.oz, rw, robin ? 1 : 0, coef, kx, ky);
  return 0;
}

// ---- diagnostics --------------------------------------------------------------------------------------
static double norm_tmp(const Plan& p) {
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  return 1.0 / (N * N) / (double)(p.nz - p.Cz);
}

// product (pseudospec_phd.f90:199-272)
int product(Plan& p, const cplx* a, const cplx* b, double* out) {
  cplx *w1, *w2;
  if (plan_cwork(p, 2, &w1) || plan_cwork(p, 3, &w2)) return 1;
  if (fft1d_z_bwd(p, a, w1, 1.0) || fft1d_z_bwd(p, b, w2, 1.0)) return 1;
  return op_reduce_phys(p, w1, w2, 1, -1, norm_tmp(p), out);
}

// helicity (pseudospec_hd.f90:638-775): sum over components of Re( IFFT_z(A_q) conj IFFT_z((curl A)_q) )
int helicity(Plan& p, const cplx* a, const cplx* b, const cplx* c, double* out) {
  cplx *w1, *w2;
  if (plan_cwork(p, 2, &w1) || plan_cwork(p, 3, &w2)) return 1;
  const cplx* A[3] = {a, b, c};
  double s = 0.0, t = 0.0;
  for (int q = 0; q < 3; ++q) {
    const int i1 = q == 0 ? 1 : 0, i2 = q == 2 ? 1 : 2;  // (b,c), (a,c), (a,b)
    if (op_curlk(p, A[i1], A[i2], w2, q + 1)) return 1;
    if (fft1d_z_bwd(p, A[q], w1, 1.0) || fft1d_z_bwd(p, w2, w2, 1.0)) return 1;
    if (op_reduce_phys(p, w1, w2, 1, -1, norm_tmp(p), &t)) return 1;
    s += t;
  }
  *out = s;
  return 0;
}

// maxabs (pseudospec_hd.f90:1008-1079)
int maxabs(Plan& p, const cplx* a, const cplx* b, const cplx* c, int kin, double* out) {
  SX_REQUIRE(kin >= 0 && kin <= 2, "maxabs: kin must be 0, 1 or 2");
  cplx* w[3];
  double* r[3];
  for (int q = 0; q < 3; ++q) if (plan_cwork(p, 15 + q, &w[q]) || plan_rwork(p, q, &r[q])) return 1;
  const cplx* in[3] = {a, b, c};
  if (kin == 0) {
    if (op_curlk(p, b, c, w[0], 1) || op_curlk(p, a, c, w[1], 2) || op_curlk(p, a, b, w[2], 3)) return 1;
    for (int q = 0; q < 3; ++q) in[q] = w[q];
  } else if (kin == 1) {
    for (int q = 0; q < 3; ++q) { if (op_laplak(p, in[q], w[q])) return 1; in[q] = w[q]; }
  }
  for (int q = 0; q < 3; ++q) if (fft3d_c2r(p, in[q], r[q])) return 1;
  const int nzp = p.pkend - p.ksta + 1;
  double m = 0.0;
  if (nzp > 0) {
    const size_t n = (size_t)p.nx * p.ny * nzp;
    const int blocks = p.red_blocks;
    double* partial = p.d_red;
    cudaStream_t st = p.stream;
    auto kfn = k_max_norm3;
    if (stage_mark(p, ST_REDUCE)) return 1;
    SX_LAUNCH(kfn, dim3(blocks), dim3(256), 256 * sizeof(double), st, n, r[0], r[1], r[2], partial);
    p.launches++;
    SX_KERNEL_CHECK();
    SX_CUDA_CHECK(cudaMemcpyAsync(p.h_red, p.d_red, blocks * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
    for (int i = 0; i < blocks; ++i) m = std::max(m, p.h_red[i]);
  }
  // MPI_REDUCE(MPI_MAX) (:1076) through the sum all-reduce: every rank fills its own slot of a zero vector
  std::vector<double> slots(p.nprocs, 0.0);
  slots[p.myrank] = m;
  if (p.nprocs > 1 && allreduce_sum(p, slots.data(), p.nprocs)) return 1;
  m = *std::max_element(slots.begin(), slots.end());
  *out = sqrt(m) / ((double)p.nx * (double)p.ny * (double)p.nz);
  return 0;
}

// robcheck (bboundary.f90:434-602): out = tangential z=0, tangential z=Lz, normal z=0, normal z=Lz
int robcheck(Plan& p, const cplx* a, const cplx* b, const cplx* c, double out[4]) {
  SX_REQUIRE(p.Cz > 0, "robcheck needs a non-periodic z direction (Cz > 0)");
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  const double tmp = 1.0 / (N * N);
  const Dims d = dims_of(p);
  const int top = p.nphys() - 1;
  const size_t n = 2 * (size_t)p.ny * p.nxl;
  const double *kx = p.d_kx, *ky = p.d_ky;
  cplx *C1, *C2;
  if (plan_cwork(p, 19, &C1) || plan_cwork(p, 20, &C2)) return 1;
  const cplx* f[3] = {a, b, c};
  double r[3][2];
  for (int q = 0; q < 3; ++q) {
    if (op_derivk(p, f[q], C2, 3)) return 1;
    if (fft1d_z_bwd(p, f[q], C1, 1.0) || fft1d_z_bwd(p, C2, C2, 1.0)) return 1;
    SX_EW_LAUNCH(p, k_robin_residual, n, d, C1, C2, top, kx, ky);
    if (op_reduce_phys(p, C2, nullptr, 0, 0, tmp, &r[q][0])) return 1;
    if (op_reduce_phys(p, C2, nullptr, 0, top, tmp, &r[q][1])) return 1;
  }
  out[0] = r[0][0] + r[1][0];
  out[1] = r[0][1] + r[1][1];
  out[2] = r[2][0];
  out[3] = r[2][1];
  return 0;
}

// bdiagnostic (bboundary.f90:348-430): which bit 0 = conducting[6] filled, bit 1 = vacuum[6] filled
int bdiagnostic(Plan& p, const cplx* a, const cplx* b, const cplx* c, double conducting[6], double vacuum[6],
                int* which) {
  cplx *c1, *c2, *c3, *c4;
  if (plan_cwork(p, 15, &c1) || plan_cwork(p, 16, &c2) || plan_cwork(p, 17, &c3) || plan_cwork(p, 18, &c4)) return 1;
  if (op_curlk(p, b, c, c1, 1) || op_curlk(p, a, c, c2, 2) || op_curlk(p, a, b, c3, 3)) return 1;
  double tm1, tm2;
  if (divergence(p, a, b, c, &tm1) || divergence(p, c1, c2, c3, &tm2)) return 1;
  int w = 0;
  if (p.b_bczsta == 0 || p.b_bczend == 0) {
    double tmp, tmq, tmr, tms;
    if (bouncheck_z(p, &tmr, &tms, c3, nullptr)) return 1;
    if (op_curlk(p, c2, c3, c4, 1) || op_curlk(p, c1, c3, c2, 2)) return 1;   // the current (:397-399)
    if (bouncheck_z(p, &tmp, &tmq, c4, c2)) return 1;
    const double v[6] = {tm1, tm2, tmp, tmq, tmr, tms};
    memcpy(conducting, v, sizeof(v));
    w |= 1;
  }
  if (p.b_bczsta == 1 || p.b_bczend == 1) {
    vacuum[0] = tm1;
    vacuum[1] = tm2;
    if (robcheck(p, a, b, c, vacuum + 2)) return 1;
    w |= 2;
  }
  *which = w;
  return 0;
}

// mhdcheck (pseudospec_mhd.f90:109-212): out = eng, ens, cur, engk, engm, helk, helm, crh, asq
int mhdcheck(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* ma, const cplx* mb, const cplx* mc,
             int hel, int crs, double out[9]) {
  double engk, engm, ens, cur;
  if (energy(p, a, b, c, 1, &engk) || energy(p, a, b, c, 0, &ens)) return 1;
  if (energy(p, ma, mb, mc, 0, &engm) || energy(p, ma, mb, mc, 2, &cur)) return 1;
  out[0] = engk + engm; out[1] = ens; out[2] = cur; out[3] = engk; out[4] = engm;
  out[5] = out[6] = out[7] = out[8] = 0.0;
  if (hel == 1 && (helicity(p, a, b, c, &out[5]) || helicity(p, ma, mb, mc, &out[6]))) return 1;
  if (crs == 1) {
    cplx *c1, *c2, *c3;
    if (plan_cwork(p, 15, &c1) || plan_cwork(p, 16, &c2) || plan_cwork(p, 17, &c3)) return 1;
    if (energy(p, ma, mb, mc, 1, &out[8])) return 1;
    if (op_derivk(p, ma, c1, 1) || op_derivk(p, mb, c2, 2) || op_derivk(p, mc, c3, 3)) return 1;
    if (cross(p, a, b, c, c1, c2, c3, 1, &out[7])) return 1;
  }
  return 0;
}

// ---- setup_bc -----------------------------------------------------------------------------------------
static std::string preprocess(const char* s) {   // trims and lower-cases, like `preprocess' of the reference
  std::string t = s ? s : "";
  size_t b = 0, e = t.size();
  while (b < e && isspace((unsigned char)t[b])) ++b;
  while (e > b && isspace((unsigned char)t[e - 1])) --e;
  t = t.substr(b, e - b);
  for (char& ch : t) ch = (char)tolower((unsigned char)ch);
  return t;
}

// v_parsebc (vboundary.f90:14-37), s_parsebc (sboundary.f90:14-37), b_parsebc (bboundary.f90:15-40)
static int parsebc(char field, const std::string& str, int* bc) {
  if (str == "periodic") { *bc = -1; return 0; }
  if (field == 'v' && str == "noslip") { *bc = 0; return 0; }
  if (field == 's' && str == "constant") { *bc = 0; return 0; }
  if (field == 'b' && str == "conducting") { *bc = 0; return 0; }
  if (field == 'b' && str == "vacuum") { *bc = 1; return 0; }
  SX_REQUIRE(false, "Unknown boundary condition type " + str + " Aborting...");
}

int setup_bc(Plan& p, const char* field, const char* const bckind[6]) {
  SX_REQUIRE(field && bckind, "setup_bc: null argument");
  const std::string f = preprocess(field);
  SX_REQUIRE(f == "v" || f == "s" || f == "b", "setup_bc: field must be 'v', 's' or 'b'");
  int bc[6];
  for (int i = 0; i < 6; ++i) if (parsebc(f[0], preprocess(bckind[i]), &bc[i])) return 1;
  SX_REQUIRE(bc[0] == -1 && bc[1] == -1 && bc[2] == -1 && bc[3] == -1,
             "Non-periodic boundary conditions in the X and Y directions are not supported. Aborting...");
  if (p.Cz > 0)
    SX_REQUIRE(bc[4] >= 0 && bc[5] >= 0, "a non-periodic z direction (Cz > 0) needs wall boundary conditions at z=0 and z=Lz");
  else
    SX_REQUIRE(bc[4] == -1 && bc[5] == -1, "a periodic z direction (Cz = 0) takes 'periodic' at z=0 and z=Lz");
  if (f == "v") { p.v_bczsta = bc[4]; p.v_bczend = bc[5]; }
  else if (f == "s") { p.s_bczsta = bc[4]; p.s_bczend = bc[5]; }
  else {
    p.b_bczsta = bc[4]; p.b_bczend = bc[5];
    if (p.Cz > 0 && load_neumann(p)) return 1;   // b_setup loads the Neumann tables (bboundary.f90:83-95)
  }
  return 0;
}

}  // namespace sx

using namespace sx;
#define SX_PLAN(pl) \
  if (!(pl)) { sx::set_error("[ERROR] null plan"); return 1; } \
  sx::Plan& p = (pl)->p
static inline cplx* C(double* a) { return reinterpret_cast<cplx*>(a); }
static inline const cplx* C(const double* a) { return reinterpret_cast<const cplx*>(a); }

extern "C" {

int sx_setup_bc(sx_plan* plan, const char* field, const char* const bckind[6]) { SX_PLAN(plan); return setup_bc(p, field, bckind); }
int sx_neumann_reconstruct(sx_plan* plan, double* f, int boun, int order) {
  SX_PLAN(plan); return wall_reconstruct(p, C(f), boun, order, false, nullptr);
}
int sx_robin_reconstruct(sx_plan* plan, double* f, int boun, const double* a) {
  SX_PLAN(plan); return wall_reconstruct(p, C(f), boun, 1, true, a);
}
int sx_helicity(sx_plan* plan, const double* a, const double* b, const double* c, double* out) {
  SX_PLAN(plan); return helicity(p, C(a), C(b), C(c), out);
}
int sx_product(sx_plan* plan, const double* a, const double* b, double* out) { SX_PLAN(plan); return product(p, C(a), C(b), out); }
int sx_pscheck(sx_plan* plan, const double* a, const double* b, double out[3]) {
  SX_PLAN(plan);
  return variance(p, C(a), 1, &out[0]) || variance(p, C(a), 0, &out[1]) || product(p, C(a), C(b), &out[2]);
}
int sx_maxabs(sx_plan* plan, const double* a, const double* b, const double* c, int kin, double* out) {
  SX_PLAN(plan); return maxabs(p, C(a), C(b), C(c), kin, out);
}
int sx_mhdcheck(sx_plan* plan, const double* a, const double* b, const double* c, const double* ma, const double* mb,
                const double* mc, int hel, int crs, double out[9]) {
  SX_PLAN(plan); return mhdcheck(p, C(a), C(b), C(c), C(ma), C(mb), C(mc), hel, crs, out);
}
int sx_robcheck(sx_plan* plan, const double* a, const double* b, const double* c, double out[4]) {
  SX_PLAN(plan); return robcheck(p, C(a), C(b), C(c), out);
}
int sx_bdiagnostic(sx_plan* plan, const double* a, const double* b, const double* c, double conducting[6],
                   double vacuum[6], int* which) {
  SX_PLAN(plan); return bdiagnostic(p, C(a), C(b), C(c), conducting, vacuum, which);
}
int sx_sdiagnostic(sx_plan* plan, const double* a, double out[2]) {
  SX_PLAN(plan); return bouncheck_z(p, &out[0], &out[1], C(a), nullptr);
}

}  // extern "C"
