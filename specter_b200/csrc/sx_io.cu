// The data formats either side of the path: the reference's binary field files (mpiio/binary_io.f90) and the
// output / restart blocks of the HD driver that produce and consume them (specter.fpp:1005-1053, 886-912).
//
// File format (binary_io.f90:15-86, 165-223): `<dir>/<name>.<nmb>.out`, raw native-endian reals, Fortran order,
// global extent (nx-Cx, ny-Cy, nz-Cz) -- the physical box, continuation planes are never written -- each rank
// owning the z planes ksta..min(kend, nz-Cz) of the MPI-IO subarray view.  Here every rank pwrite()s / pread()s its
// planes at the byte offset of its first plane, which is the same file image.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/specter_b200.h"
#include "sx_plan.h"

namespace sx {

static std::string io_path(const char* dir, const char* fname, const char* nmb) {
  return std::string(dir) + "/" + fname + "." + nmb + ".out";   // binary_io.f90:202-204
}

// planes of this rank that exist in the file
static void io_extent(const Plan& p, size_t* first_plane, size_t* nplanes) {
  const int kend = p.kend < p.nphys() ? p.kend : p.nphys();
  *first_plane = (size_t)(p.ksta - 1);
  *nplanes = kend >= p.ksta ? (size_t)(kend - p.ksta + 1) : 0;
}

int io_write(Plan& p, const double* real_dev, const char* dir, const char* fname, const char* nmb) {
  size_t k0, nk;
  io_extent(p, &k0, &nk);
  const size_t plane = (size_t)p.nx * p.ny;
  std::vector<double> host(plane * nk);
  if (nk) {
    SX_CUDA_CHECK(cudaMemcpyAsync(host.data(), real_dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  }
  const std::string path = io_path(dir, fname, nmb);
  const int fd = open(path.c_str(), O_CREAT | O_WRONLY, 0644);
  SX_REQUIRE(fd >= 0, "io_write: cannot open file for writing: " + path);
  size_t done = 0;
  const size_t bytes = host.size() * sizeof(double);
  const off_t off = (off_t)(k0 * plane * sizeof(double));
  while (done < bytes) {
    const ssize_t w = pwrite(fd, (const char*)host.data() + done, bytes - done, off + (off_t)done);
    if (w <= 0) { close(fd); SX_REQUIRE(false, "io_write: short write to " + path); }
    done += (size_t)w;
  }
  close(fd);
  return 0;
}

int io_read(Plan& p, double* real_dev, const char* dir, const char* fname, const char* nmb) {
  size_t k0, nk;
  io_extent(p, &k0, &nk);
  const size_t plane = (size_t)p.nx * p.ny;
  const std::string path = io_path(dir, fname, nmb);
  const int fd = open(path.c_str(), O_RDONLY);
  SX_REQUIRE(fd >= 0, "io_read: cannot open file for reading: " + path);   // binary_io.f90:129-133
  std::vector<double> host(plane * nk);
  size_t done = 0;
  const size_t bytes = host.size() * sizeof(double);
  const off_t off = (off_t)(k0 * plane * sizeof(double));
  while (done < bytes) {
    const ssize_t r = pread(fd, (char*)host.data() + done, bytes - done, off + (off_t)done);
    if (r <= 0) { close(fd); SX_REQUIRE(false, "io_read: file too short: " + path); }
    done += (size_t)r;
  }
  close(fd);
  // planes above the physical box are not in the file: zero them like a freshly allocated slab
  SX_CUDA_CHECK(cudaMemsetAsync(real_dev, 0, p.rsize() * sizeof(double), p.stream));
  if (nk) SX_CUDA_CHECK(cudaMemcpyAsync(real_dev, host.data(), bytes, cudaMemcpyHostToDevice, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  return 0;
}

// specter.fpp:1005-1053: C = v/N -> (optional vorticity) -> c2r -> io_write; p = p'/(nx ny dt) -> c2r_xy -> io_write
int hd_output(Plan& p, cplx* const* st, const char* odir, const char* ext, double dt, int outs) {
  cplx *c1, *c2, *c3, *c4;
  double* r1;
  if (plan_cwork(p, 2, &c1) || plan_cwork(p, 3, &c2) || plan_cwork(p, 4, &c3) || plan_cwork(p, 5, &c4)) return 1;
  if (plan_rwork(p, 0, &r1)) return 1;
  const double rmp = 1.0 / ((double)p.nx * (double)p.ny * (double)p.nz);
  if (op_scale_copy(p, st[0], c1, rmp) || op_scale_copy(p, st[1], c2, rmp) || op_scale_copy(p, st[2], c3, rmp)) return 1;
  if (outs >= 1) {
    const char* wn[3] = {"wx", "wy", "wz"};
    for (int d = 1; d <= 3; ++d) {
      const cplx* a = d == 1 ? c2 : c1;
      const cplx* b = d == 3 ? c2 : c3;
      if (op_curlk(p, a, b, c4, d) || fft3d_c2r(p, c4, r1) || io_write(p, r1, odir, wn[d - 1], ext)) return 1;
    }
  }
  const char* vn[3] = {"vx", "vy", "vz"};
  cplx* c[3] = {c1, c2, c3};
  for (int q = 0; q < 3; ++q)
    if (fft3d_c2r(p, c[q], r1) || io_write(p, r1, odir, vn[q], ext)) return 1;
  // pressure: p' -> p (specter.fpp:1040-1053)
  if (op_scale_copy(p, st[3], c1, 1.0 / ((double)p.nx * (double)p.ny * dt))) return 1;
  if (fft2d_xy_c2r(p, c1, r1, p.nz) || io_write(p, r1, odir, "pr", ext)) return 1;
  return 0;
}

// specter.fpp:886-912: io_read -> r2c for v; pr -> r2c_xy, physical rows x dt (back to p')
int hd_restart(Plan& p, cplx* const* st, const char* idir, const char* ext, double dt) {
  double* r1;
  if (plan_rwork(p, 0, &r1)) return 1;
  const char* vn[3] = {"vx", "vy", "vz"};
  for (int q = 0; q < 3; ++q)
    if (io_read(p, r1, idir, vn[q], ext) || fft3d_r2c(p, r1, st[q])) return 1;
  if (io_read(p, r1, idir, "pr", ext) || fft2d_xy_r2c(p, r1, st[3], p.nz)) return 1;
  return op_scale_phys(p, st[3], dt);
}

// the SCALAR_ block of the BIN output (specter.fpp:1055-1069): th/N -> c2r -> io_write
int scalar_output(Plan& p, const cplx* th, const char* odir, const char* ext) {
  cplx* c1;
  double* r1;
  if (plan_cwork(p, 2, &c1) || plan_rwork(p, 0, &r1)) return 1;
  const double rmp = 1.0 / ((double)p.nx * (double)p.ny * (double)p.nz);
  return op_scale_copy(p, th, c1, rmp) || fft3d_c2r(p, c1, r1) || io_write(p, r1, odir, "th", ext);
}

// the MAGFIELD_ block (specter.fpp:1071-1128): a/N; outs >= 1: b = curl a; outs == 2: j = laplak(a) as written
// there (the sign convention of the reference); a; ph = ph'/(nx ny dt) -> c2r_xy
int magnetic_output(Plan& p, cplx* const a[3], const cplx* ph, const char* odir, const char* ext, double dt, int outs) {
  cplx *c1, *c2, *c3, *c4;
  double* r1;
  if (plan_cwork(p, 2, &c1) || plan_cwork(p, 3, &c2) || plan_cwork(p, 4, &c3) || plan_cwork(p, 5, &c4)) return 1;
  if (plan_rwork(p, 0, &r1)) return 1;
  const double rmp = 1.0 / ((double)p.nx * (double)p.ny * (double)p.nz);
  if (op_scale_copy(p, a[0], c1, rmp) || op_scale_copy(p, a[1], c2, rmp) || op_scale_copy(p, a[2], c3, rmp)) return 1;
  cplx* c[3] = {c1, c2, c3};
  if (outs >= 1) {
    const char* bn[3] = {"bx", "by", "bz"};
    for (int d = 1; d <= 3; ++d) {
      const cplx* u = d == 1 ? c2 : c1;
      const cplx* v = d == 3 ? c2 : c3;
      if (op_curlk(p, u, v, c4, d) || fft3d_c2r(p, c4, r1) || io_write(p, r1, odir, bn[d - 1], ext)) return 1;
    }
  }
  if (outs == 2) {
    const char* jn[3] = {"jx", "jy", "jz"};
    for (int q = 0; q < 3; ++q)
      if (op_laplak(p, c[q], c4) || fft3d_c2r(p, c4, r1) || io_write(p, r1, odir, jn[q], ext)) return 1;
  }
  const char* an[3] = {"ax", "ay", "az"};
  for (int q = 0; q < 3; ++q)
    if (fft3d_c2r(p, c[q], r1) || io_write(p, r1, odir, an[q], ext)) return 1;
  if (op_scale_copy(p, ph, c1, 1.0 / ((double)p.nx * (double)p.ny * dt))) return 1;
  return fft2d_xy_c2r(p, c1, r1, p.nz) || io_write(p, r1, odir, "ph", ext);
}

// specter.fpp:935-938 and :941-957 (dyna = 0): th; a and ph (physical rows x dt, back to ph')
int scalar_restart(Plan& p, cplx* th, const char* idir, const char* ext) {
  double* r1;
  if (plan_rwork(p, 0, &r1)) return 1;
  return io_read(p, r1, idir, "th", ext) || fft3d_r2c(p, r1, th);
}
int magnetic_restart(Plan& p, cplx* const a[3], cplx* ph, const char* idir, const char* ext, double dt) {
  double* r1;
  if (plan_rwork(p, 0, &r1)) return 1;
  const char* an[3] = {"ax", "ay", "az"};
  for (int q = 0; q < 3; ++q)
    if (io_read(p, r1, idir, an[q], ext) || fft3d_r2c(p, r1, a[q])) return 1;
  if (io_read(p, r1, idir, "ph", ext) || fft2d_xy_r2c(p, r1, ph, p.nz)) return 1;
  return op_scale_phys(p, ph, dt);
}

// benchmark.txt (specter.fpp:1182-1228, the non-CUDA column set): appended by rank 0, header on creation.
// TFFT / TTRA / TCOM / TCONT / TNEU / TROB / TTOT come from the stage timers (seconds per step): every transform
// kernel carries its transposition, continuation and wall reconstructions fused in, so TTRA, TCONT, TNEU and TROB
// are part of TFFT and reported as zero; TCOM is the device time of the slab exchanges.
int benchmark_write(Plan& p, const char* path, int nsteps, int nth, double tcpu, double tomp, double twtime) {
  SX_REQUIRE(path && nsteps > 0, "sx_benchmark_write: bad arguments");
  if (stage_flush(p)) return 1;
  double fft = 0.0, com = 0.0, tot = 0.0;
  for (int i = 0; i < ST_COUNT; ++i) {
    const double s = p.timer.ms[i] * 1e-3;
    tot += s;
    if (i == ST_EXCHANGE) com += s;
    else if (i != ST_EW && i != ST_REDUCE && i != ST_OTHER) fft += s;
  }
  if (p.myrank != 0) return 0;
  struct stat sb;
  const bool exists = stat(path, &sb) == 0;
  FILE* f = fopen(path, "a");
  SX_REQUIRE(f != nullptr, std::string("sx_benchmark_write: cannot open ") + path);
  if (!exists) fprintf(f, " # nx ny nz nsteps nprocs nth TCPU TOMP TWTIME TFFT TTRA TCOMTCONT TNEU TROB TTOT\n");
  const double n = (double)nsteps;
  fprintf(f, " %11d %11d %11d %11d %11d %11d %24.16E %24.16E %24.16E %24.16E %24.16E %24.16E %24.16E %24.16E %24.16E %24.16E\n",
          p.nx, p.ny, p.nz, nsteps, p.nprocs, nth, tcpu / n, tomp / n, twtime / n, fft / n, 0.0, com / n, 0.0, 0.0, 0.0,
          tot / n);
  fclose(f);
  return 0;
}

}  // namespace sx

using namespace sx;
#define SX_PLAN(pl) \
  if (!(pl)) { sx::set_error("[ERROR] null plan"); return 1; } \
  sx::Plan& p = (pl)->p

extern "C" {

int sx_io_write(sx_plan* plan, const double* real_dev, const char* dir, const char* fname, const char* nmb) {
  SX_PLAN(plan);
  SX_REQUIRE(real_dev && dir && fname && nmb, "sx_io_write: null argument");
  return io_write(p, real_dev, dir, fname, nmb);
}

int sx_io_read(sx_plan* plan, double* real_dev, const char* dir, const char* fname, const char* nmb) {
  SX_PLAN(plan);
  SX_REQUIRE(real_dev && dir && fname && nmb, "sx_io_read: null argument");
  return io_read(p, real_dev, dir, fname, nmb);
}

int sx_hd_output(sx_plan* plan, const char* odir, const char* ext, double dt, int outs) {
  SX_PLAN(plan);
  SX_REQUIRE(odir && ext && dt > 0.0, "sx_hd_output: bad arguments");
  double* d[4];
  for (int i = 0; i < 4; ++i)
    if (sx_hd_state_ptr(plan, i, &d[i])) return 1;
  cplx* st[4] = {(cplx*)d[0], (cplx*)d[1], (cplx*)d[2], (cplx*)d[3]};
  return hd_output(p, st, odir, ext, dt, outs);
}

int sx_hd_restart(sx_plan* plan, const char* idir, const char* ext, double dt) {
  SX_PLAN(plan);
  SX_REQUIRE(idir && ext && dt > 0.0, "sx_hd_restart: bad arguments");
  double* d[4];
  for (int i = 0; i < 4; ++i)
    if (sx_hd_state_ptr(plan, i, &d[i])) return 1;
  cplx* st[4] = {(cplx*)d[0], (cplx*)d[1], (cplx*)d[2], (cplx*)d[3]};
  return hd_restart(p, st, idir, ext, dt);
}

/* The BIN output block and the restart branch for any solver of the reference (specter.fpp:1005-1128, 886-957):
 * solver = "HD" | "BOUSS" | "ROTBOUSS" | "MHD" | "MHDBOUSS", acting on that solver's plan-owned state. */
static int solver_fields(sx_plan* plan, const char* solver, cplx* v[4], cplx** th, cplx* a[3], cplx** ph) {
  sx::Plan& p = plan->p;
  (void)p;
  const std::string s = solver ? solver : "";
  *th = *ph = nullptr;
  a[0] = a[1] = a[2] = nullptr;
  double* d = nullptr;
  auto get = [&](int (*fn)(sx_plan*, int, double**), int which, cplx** out) -> int {
    if (fn(plan, which, &d)) return 1;
    *out = (cplx*)d;
    return 0;
  };
  int (*fn)(sx_plan*, int, double**) = nullptr;
  int ith = -1, ia = -1;
  if (s == "HD") fn = sx_hd_state_ptr;
  else if (s == "BOUSS" || s == "ROTBOUSS") { fn = sx_bouss_state_ptr; ith = 10; }
  else if (s == "MHD") { fn = sx_mhd_state_ptr; ia = 10; }
  else if (s == "MHDBOUSS") { fn = sx_mhdbouss_state_ptr; ia = 10; ith = 20; }
  else SX_REQUIRE(false, "unknown solver '" + s + "' (HD, BOUSS, ROTBOUSS, MHD, MHDBOUSS)");
  for (int i = 0; i < 4; ++i) if (get(fn, i, &v[i])) return 1;
  if (ith >= 0 && get(fn, ith, th)) return 1;
  if (ia >= 0) {
    for (int q = 0; q < 3; ++q) if (get(fn, ia + q, &a[q])) return 1;
    if (get(fn, ia + 3, ph)) return 1;
  }
  return 0;
}

int sx_output(sx_plan* plan, const char* solver, const char* odir, const char* ext, double dt, int outs) {
  SX_PLAN(plan);
  SX_REQUIRE(odir && ext && dt > 0.0, "sx_output: bad arguments");
  cplx *v[4], *th, *a[3], *ph;
  if (solver_fields(plan, solver, v, &th, a, &ph)) return 1;
  if (hd_output(p, v, odir, ext, dt, outs)) return 1;
  if (th && scalar_output(p, th, odir, ext)) return 1;
  if (ph && magnetic_output(p, a, ph, odir, ext, dt, outs)) return 1;
  return 0;
}

int sx_restart(sx_plan* plan, const char* solver, const char* idir, const char* ext, double dt) {
  SX_PLAN(plan);
  SX_REQUIRE(idir && ext && dt > 0.0, "sx_restart: bad arguments");
  cplx *v[4], *th, *a[3], *ph;
  if (solver_fields(plan, solver, v, &th, a, &ph)) return 1;
  if (hd_restart(p, v, idir, ext, dt)) return 1;
  if (th && scalar_restart(p, th, idir, ext)) return 1;
  if (ph && magnetic_restart(p, a, ph, idir, ext, dt)) return 1;
  return 0;
}

/* ---- the global-quantity text files (include/<solver>/<solver>_global.f90) ------------------------------------------
 * One `1P Ew.d' field as the Fortran run-time prints it: d.dddE+ee right-justified; three-digit exponents drop the E. */
static std::string fortran_e(double x, int w, int d) {
  char buf[64];
  std::string s;
  if (x != x) s = "NaN";
  else if (x > 1.7976931348623157e308) s = "Infinity";
  else if (x < -1.7976931348623157e308) s = "-Infinity";
  else {
    snprintf(buf, sizeof buf, "%.*E", d, x);
    s = buf;
    const size_t e = s.find('E');
    const int ex = atoi(s.c_str() + e + 1);
    char eb[16];
    if (ex > -100 && ex < 100) snprintf(eb, sizeof eb, "E%+03d", ex);
    else snprintf(eb, sizeof eb, "%+04d", ex);
    s = s.substr(0, e) + eb;
  }
  if ((int)s.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)w - s.size(), ' ') + s;
}
struct Field { double x; int w, d; };
static int append_row(const std::string& path, const std::vector<Field>& f) {
  FILE* fp = fopen(path.c_str(), "a");
  SX_REQUIRE(fp != nullptr, "cannot open " + path + " for appending");
  std::string line;
  for (const Field& q : f) line += fortran_e(q.x, q.w, q.d);
  fprintf(fp, "%s\n", line.c_str());
  fclose(fp);
  return 0;
}

int sx_global(sx_plan* plan, const char* solver, const char* odir, int t, double dt) {
  SX_PLAN(plan);
  SX_REQUIRE(odir != nullptr, "sx_global: bad arguments");
  cplx *v[4], *th, *a[3], *ph;
  if (solver_fields(plan, solver, v, &th, a, &ph)) return 1;
  const std::string s = solver, dir = std::string(odir) + "/";
  const bool root = p.myrank == 0;   // the reductions are collective; rank 0 writes (IF (myrank.eq.0))
  const double tl = (double)(t - 1) * dt;
  int (*fn)(sx_plan*, int, double**) = s == "HD" ? sx_hd_state_ptr : (s == "MHD" ? sx_mhd_state_ptr
                                       : (s == "MHDBOUSS" ? sx_mhdbouss_state_ptr : sx_bouss_state_ptr));
  double *vx = (double*)v[0], *vy = (double*)v[1], *vz = (double*)v[2];
  if (!ph) {   // hdcheck(vx,vy,vz,fx,fy,fz,t,dt,1,0): pseudospec_hd.f90:943-1005
    double *f[3], eng, ens, pot, khe;
    for (int q = 0; q < 3; ++q) if (fn(plan, 4 + q, &f[q])) return 1;
    if (sx_hdcheck(plan, vx, vy, vz, f[0], f[1], f[2], &eng, &ens, &pot) || sx_helicity(plan, vx, vy, vz, &khe)) return 1;
    if (root && (append_row(dir + "balance.txt", {{tl, 13, 6}, {eng, 23, 16}, {ens, 23, 16}, {pot, 24, 16}}) ||
                 append_row(dir + "helicity.txt", {{tl, 13, 6}, {khe, 24, 16}}))) return 1;
  } else {     // mhdcheck(vx,vy,vz,ax,ay,az,t,dt,1,1): pseudospec_mhd.f90:109-212
    double o[9];
    if (sx_mhdcheck(plan, vx, vy, vz, (double*)a[0], (double*)a[1], (double*)a[2], 1, 1, o)) return 1;
    if (root && (append_row(dir + "balance.txt", {{tl, 13, 6}, {o[0], 23, 16}, {o[1], 23, 16}, {o[2], 23, 16}}) ||
                 append_row(dir + "energy.txt", {{tl, 13, 6}, {o[3], 23, 16}, {o[4], 23, 16}}) ||
                 append_row(dir + "helicity.txt", {{tl, 13, 6}, {o[5], 24, 16}, {o[6], 24, 16}}) ||
                 append_row(dir + "cross.txt", {{tl, 13, 6}, {o[7], 23, 16}, {o[8], 24, 16}}))) return 1;
  }
  if (th) {    // pscheck(th,fs,t,dt): pseudospec_phd.f90:275-321
    double *fs, o[3];
    if (fn(plan, s == "MHDBOUSS" ? 21 : 11, &fs) || sx_pscheck(plan, (double*)th, fs, o)) return 1;
    if (root && append_row(dir + "scalar.txt", {{tl, 13, 6}, {o[0], 22, 14}, {o[1], 22, 14}, {o[2], 23, 14}})) return 1;
  }
  {            // vdiagnostic: vboundary.f90:214-269
    double o[5];
    if (sx_vdiagnostic(plan, vx, vy, vz, o)) return 1;
    if (root && append_row(dir + "noslip_diagnostic.txt",
                           {{tl, 13, 6}, {o[0], 13, 6}, {o[1], 13, 6}, {o[2], 13, 6}, {o[3], 13, 6}, {o[4], 13, 6}})) return 1;
  }
  if (ph) {    // bdiagnostic: bboundary.f90:348-430
    double c[6], vac[6];
    int which = 0;
    if (sx_bdiagnostic(plan, (double*)a[0], (double*)a[1], (double*)a[2], c, vac, &which)) return 1;
    for (int k = 0; k < 2; ++k) {
      if (!(which & (1 << k)) || !root) continue;
      const double* o = k == 0 ? c : vac;
      if (append_row(dir + (k == 0 ? "conducting_diagnostic.txt" : "vacuum_diagnostic.txt"),
                     {{tl, 13, 6}, {o[0], 13, 6}, {o[1], 13, 6}, {o[2], 13, 6}, {o[3], 13, 6}, {o[4], 13, 6}, {o[5], 13, 6}})) return 1;
    }
  }
  if (th) {    // sdiagnostic: sboundary.f90:168-210
    double o[2];
    if (sx_sdiagnostic(plan, (double*)th, o)) return 1;
    if (root && append_row(dir + "scalar_constant_diagnostic.txt", {{tl, 13, 6}, {o[0], 13, 6}, {o[1], 13, 6}})) return 1;
  }
  return 0;
}

int sx_benchmark_write(sx_plan* plan, const char* path, int nsteps, int nth, double tcpu, double tomp, double twtime) {
  SX_PLAN(plan);
  return benchmark_write(p, path, nsteps, nth, tcpu, tomp, twtime);
}

}  // extern "C"
