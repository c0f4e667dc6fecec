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

}  // extern "C"
