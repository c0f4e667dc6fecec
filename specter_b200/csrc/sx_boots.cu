// BOOTS regridder (tools/boots.fpp): prolongation of a field file of the old grid (nxt, nyt, nzt physical rows) to the
// new grid (nx, ny, nzp = nz-Cz physical rows) by zero padding in Fourier space, with the FC-Gram continuation making
// the non-periodic z direction periodic on both grids (Czt / Czn continuation points so that the two periods coincide,
// boots.fpp:181-182).
//
// The x and y directions are powers of two on both grids and use the transform kernels of the solver.  The z lengths
// m = nzt+Czt and M = nzp+Czn are whatever the gcd gives (567 = 3^4 7 in the shipped boots.inp), so the z part --
// continuation, forward transform of length m, the padding of boots.fpp:275-300 and the backward transform of length M
// evaluated on the nzp rows that are written -- is applied as ONE dense nzp x nzt operator
//     T = fact E_M S F_m K          (K continuation, F_m / E_M the two DFTs, S the padding, fact = 1/(nxt nyt m))
// that is the same for every (ky,kx) pencil: T is built on the device (two elementwise fills and one product) and the
// field goes through it in a tiled FP64 product whose epilogue scatters the pencils to their zero-padded ky position.
// An initialisation-time tool on one GPU: O(nzp nzt) per pencil instead of O(n log n), 9e5 complex FMAs per pencil at
// the shipped sizes, ~1e11 flops per 256x128x487 file.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <numeric>
#include <string>
#include <vector>

#include "../../include/specter_b200.h"
#include "sx_plan.h"

namespace sx {

// ---- operator pieces ---------------------------------------------------------------------------------
// E[z' + c nzp] = fact e^{+2 pi i rows[c] z' / M}: the backward transform of length M restricted to the rows written
__global__ void k_boots_fill_e(cplx* __restrict__ E, const int* __restrict__ rows, int nzp, int R, int M, double fact) {
  const size_t n = (size_t)nzp * R;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t / nzp), z = (int)(t - (size_t)c * nzp);
    const long long q = ((long long)rows[c] * z) % M;
    double s, co;
    sincospi(2.0 * (double)q / (double)M, &s, &co);
    E[t] = cmake(fact * co, fact * s);
  }
}

// G[c + z R] = (F_m K)[srcs[c], z]: forward transform of length m of the continued pencil, as a matrix on the nzt
// physical rows.  K (fftp.fpp:757-772) adds to the first / last d columns the C continuation rows weighted by dir.
__global__ void k_boots_fill_g(cplx* __restrict__ G, const int* __restrict__ srcs, int R, int nzt, int m, int C, int d,
                               const double* __restrict__ dir) {
  const size_t n = (size_t)R * nzt;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(t / R), c = (int)(t - (size_t)z * R);
    const long long k = srcs[c];
    double s, co;
    sincospi(-2.0 * (double)((k * z) % m) / (double)m, &s, &co);
    double re = co, im = s;
    const int jl = z - (nzt - d);   // column of dir(ii, .) for the last d rows
    const int jf = d - 1 - z;       // column of dir(C-ii+1, .) for the first d rows
    if (jl >= 0 || jf >= 0) {
      for (int ii = 0; ii < C; ++ii) {
        double w = 0.0;
        if (jl >= 0) w += dir[(size_t)ii * d + jl];
        if (jf >= 0) w += dir[(size_t)(C - 1 - ii) * d + jf];
        sincospi(-2.0 * (double)((k * (nzt + ii)) % m) / (double)m, &s, &co);
        re += w * co;
        im += w * s;
      }
    }
    G[t] = cmake(re, im);
  }
}

// ---- tiled FP64 complex product, column-major:  C(:, dst(n)) = A (Mr x Kd) . B(:, n) ---------------------
// 64 x 64 tile per 256-thread CTA, K chunks of 16 through shared memory, 4 x 4 outputs per thread with the rows
// dealt round-robin so that shared-memory reads and the column stores are 16 consecutive complex values.
// nyt > 0: column n = (jy, kx) of the old half-spectrum goes to its zero-padded position(s) on the new one
// (boots.fpp:275-300: jy <= nyt/2 keeps its index, jy >= nyt/2-1 moves up by ny-nyt; both when the loops overlap).
constexpr int kBM = 64, kBN = 64, kBK = 16;
__global__ void __launch_bounds__(256) k_zgemm(const cplx* __restrict__ A, int lda, const cplx* __restrict__ B, int ldb,
                                                cplx* __restrict__ Cm, int ldc, int Mr, long long N, int Kd, int nyt, int ny) {
  SX_DYN_SMEM(cplx, smem);
  cplx* As = smem;                 // [kBK][kBM]
  cplx* Bs = smem + kBK * kBM;     // [kBK][kBN]
  const int tid = threadIdx.x, ti = tid % 16, tn = tid / 16;
  const int i0 = blockIdx.x * kBM;
  const long long n0 = (long long)blockIdx.y * kBN;
  cplx acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = cmake(0.0, 0.0);
  for (int k0 = 0; k0 < Kd; k0 += kBK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      {
        const int i = idx % kBM, k = idx / kBM;
        As[k * kBM + i] = (i0 + i < Mr && k0 + k < Kd) ? A[(size_t)(k0 + k) * lda + i0 + i] : cmake(0.0, 0.0);
      }
      {
        const int k = idx % kBK, n = idx / kBK;
        Bs[k * kBN + n] = (n0 + n < N && k0 + k < Kd) ? B[(size_t)(n0 + n) * ldb + k0 + k] : cmake(0.0, 0.0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      cplx a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[k * kBM + ti + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[k * kBN + tn + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[r][c].x = fma(-a[r].y, b[c].y, fma(a[r].x, b[c].x, acc[r][c].x));   // 4 DFMA per complex product
          acc[r][c].y = fma(a[r].y, b[c].x, fma(a[r].x, b[c].y, acc[r][c].y));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const long long n = n0 + tn + 16 * c;
    if (n >= N) continue;
    long long d0 = n, d1 = -1;
    if (nyt > 0) {
      const long long kx = n / nyt;
      const int jy = (int)(n - kx * nyt);
      d0 = jy <= nyt / 2 ? kx * ny + jy : -1;
      d1 = jy >= nyt / 2 - 1 ? kx * ny + jy + (ny - nyt) : -1;
      if (d1 == d0) d1 = -1;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ti + 16 * r;
      if (i >= Mr) continue;
      if (d0 >= 0) Cm[(size_t)d0 * ldc + i] = acc[r][c];
      if (d1 >= 0) Cm[(size_t)d1 * ldc + i] = acc[r][c];
    }
  }
}

static int launch_zgemm(Plan& p, const cplx* A, int lda, const cplx* B, int ldb, cplx* Cm, int ldc, int Mr, long long N,
                        int Kd, int nyt, int ny) {
  const size_t smem = (size_t)kBK * (kBM + kBN) * sizeof(cplx);
  const long long gy = (N + kBN - 1) / kBN;
  SX_REQUIRE(gy <= 65535, "boots: too many pencils for one product launch");
  dim3 grid((unsigned)((Mr + kBM - 1) / kBM), (unsigned)gy);
  if (grid.x == 0 || grid.y == 0) return 0;
  auto kfn = k_zgemm;
  cudaStream_t st = p.stream;
  SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SX_LAUNCH(kfn, grid, dim3(256), smem, st, A, lda, B, ldb, Cm, ldc, Mr, N, Kd, nyt, ny);
  p.launches++;
  SX_KERNEL_CHECK();
  return 0;
}

static int gcd_(int a, int b) {   // boots.fpp:388-400
  int t1 = a, g = b;
  for (;;) {
    const int t2 = t1 % g;
    if (t2 == 0) return g;
    t1 = g;
    g = t2;
  }
}

static int boots_points(int nzt, int nzp, int* Czt, int* Czn) {
  SX_REQUIRE(nzt >= 2 && nzp >= nzt, "MAIN: prolongation specification incorrect; input nzt must be less than Nz");
  const int g = gcd_(nzt - 1, nzp - 1);
  *Czt = (nzt - 1) / g - 1;
  *Czn = (nzp - 1) / g - 1;
  return 0;
}

struct DevBuf {   // device allocations of one regridding, released on every exit path
  std::vector<void*> ptrs;
  ~DevBuf() { for (void* q : ptrs) if (q) cudaFree(q); }
  template <class T> int get(T** out, size_t n) {
    void* q = nullptr;
    SX_CUDA_CHECK(cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
    ptrs.push_back(q);
    *out = (T*)q;
    return 0;
  }
};

struct PlanPair {
  sx_plan *told = nullptr, *tnew = nullptr;
  ~PlanPair() { if (told) sx_plan_destroy(told); if (tnew) sx_plan_destroy(tnew); }
};

static unsigned long long g_boots_launches = 0;

// one field: in_host (nxt, nyt, nzt) -> out_host (nx, ny, nzp), both x fastest
static int boots_regrid(int device, int nxt, int nyt, int nzt, int ozt, const char* tdir, int nx, int ny, int nzp,
                        const double* in_host, double* out_host) {
  SX_REQUIRE(nxt >= 1 && nxt <= nx, "MAIN: prolongation specification incorrect; input nxt must be less than Nx");   // boots.fpp:146-160
  SX_REQUIRE(nyt >= 1 && nyt <= ny, "MAIN: prolongation specification incorrect; input nyt must be less than Ny");
  SX_REQUIRE(nzt >= 1 && nzt <= nzp, "MAIN: prolongation specification incorrect; input nzt must be less than Nz");
  SX_REQUIRE(fft_size_supported(nxt, false) && fft_size_supported(nyt, false) && fft_size_supported(nx, false) &&
                 fft_size_supported(ny, false),
             "boots: nxt, nyt, nx, ny must be powers of two in [16,2048]");
  int Czt, Czn;
  if (boots_points(nzt, nzp, &Czt, &Czn)) return 1;
  // fcgram_create_plan of the old grid (fcgram_mod.f90:128-141)
  SX_REQUIRE((Czt == 0 && ozt == 0) || (Czt > 0 && ozt > 0),
             "Mismatch in continuation or matching points in z direction. Aborting...");
  SX_REQUIRE(Czt == 0 || (ozt <= 10 && 2 * ozt <= nzt), "boots: matching points do not fit the old grid");
  const int m = nzt + Czt, M = nzp + Czn, nxth = nxt / 2 + 1, nxh = nx / 2 + 1;

  PlanPair pl;
  sx_config cfg{};
  cfg.nz = 16; cfg.ord = 1; cfg.Lx = cfg.Ly = cfg.Lz = 1.0; cfg.nprocs = 1; cfg.device = device;
  cfg.nx = nxt; cfg.ny = nyt;
  if (sx_plan_create(&cfg, &pl.told)) return 1;
  cfg.nx = nx; cfg.ny = ny;
  if (sx_plan_create(&cfg, &pl.tnew)) return 1;
  Plan& po = pl.told->p;
  Plan& pn = pl.tnew->p;

  // padding S of the z direction (boots.fpp:279-284, 1-based as written; the second loop overwrites the first)
  std::vector<int> src(M, -1);
  for (int k = 1; k <= m / 2 + 1; ++k) src[k - 1] = k - 1;
  for (int k = M - m / 2; k <= M; ++k) src[k - 1] = k - M + m - 1;
  std::vector<int> rows, srcs;
  for (int k = 0; k < M; ++k)
    if (src[k] >= 0) { rows.push_back(k); srcs.push_back(src[k]); }
  const int R = (int)rows.size();

  Plan tab;   // dir = A Q^T of the old grid
  tab.Cz = Czt; tab.oz = ozt;
  if (Czt > 0) {
    SX_REQUIRE(tdir != nullptr, "tdir is required for the continuation of the old grid");
    if (load_dirichlet(tab, tdir)) return 1;
  } else {
    tab.h_dir.assign(1, 0.0);
  }

  DevBuf mem;
  int *d_rows, *d_srcs;
  double *d_dir, *d_in, *d_out;
  cplx *d_E, *d_G, *d_T, *d_A, *d_B;
  const size_t npen_old = (size_t)nyt * nxth;
  if (mem.get(&d_rows, R) || mem.get(&d_srcs, R) || mem.get(&d_dir, tab.h_dir.size())) return 1;
  if (mem.get(&d_E, (size_t)nzp * R) || mem.get(&d_G, (size_t)R * nzt) || mem.get(&d_T, (size_t)nzp * nzt)) return 1;
  if (mem.get(&d_in, (size_t)nxt * nyt * nzt) || mem.get(&d_A, npen_old * nzt)) return 1;
  if (mem.get(&d_B, (size_t)nzp * ny * nxh) || mem.get(&d_out, (size_t)nx * ny * nzp)) return 1;
  cudaStream_t so = po.stream, sn = pn.stream;
  SX_CUDA_CHECK(cudaMemcpyAsync(d_rows, rows.data(), R * sizeof(int), cudaMemcpyHostToDevice, so));
  SX_CUDA_CHECK(cudaMemcpyAsync(d_srcs, srcs.data(), R * sizeof(int), cudaMemcpyHostToDevice, so));
  SX_CUDA_CHECK(cudaMemcpyAsync(d_dir, tab.h_dir.data(), tab.h_dir.size() * sizeof(double), cudaMemcpyHostToDevice, so));
  SX_CUDA_CHECK(cudaMemcpyAsync(d_in, in_host, (size_t)nxt * nyt * nzt * sizeof(double), cudaMemcpyHostToDevice, so));

  // T = (fact E_M S) (F_m K)
  const double fact = 1.0 / ((double)nxt * (double)nyt * (double)m);   // boots.fpp:273-274
  {
    const size_t ne = (size_t)nzp * R, ng = (size_t)R * nzt;
    const unsigned ge = (unsigned)((ne + 255) / 256 < 148u * 16u ? (ne + 255) / 256 : 148u * 16u);
    const unsigned gg = (unsigned)((ng + 255) / 256 < 148u * 16u ? (ng + 255) / 256 : 148u * 16u);
    auto kfe = k_boots_fill_e;
    auto kfg = k_boots_fill_g;
    SX_LAUNCH(kfe, dim3(ge ? ge : 1), dim3(256), 0, so, d_E, d_rows, nzp, R, M, fact);
    SX_KERNEL_CHECK();
    SX_LAUNCH(kfg, dim3(gg ? gg : 1), dim3(256), 0, so, d_G, d_srcs, R, nzt, m, Czt, ozt, d_dir);
    SX_KERNEL_CHECK();
    po.launches += 2;
  }
  if (launch_zgemm(po, d_E, nzp, d_G, R, d_T, nzp, nzp, nzt, R, 0, 0)) return 1;

  // old grid: x r2c and y forward transform of the nzt physical planes (fftp.fpp:428-524) -> (nzt, nyt, nxth)
  if (launch_x_r2c(po, d_in, d_A, nzt, nzt, 1.0)) return 1;
  if (launch_yfft(po, d_A, d_A, nzt, nxth, nzt, -1, 1.0)) return 1;
  // z operator + zero padding in ky, kx -> (nzp, ny, nxh)
  SX_CUDA_CHECK(cudaMemsetAsync(d_B, 0, (size_t)nzp * ny * nxh * sizeof(cplx), so));
  if (launch_zgemm(po, d_T, nzp, d_A, nzt, d_B, nzp, nzp, (long long)npen_old, nzt, nyt, ny)) return 1;
  SX_CUDA_CHECK(cudaStreamSynchronize(so));
  // new grid: y backward transform and x c2r of the nzp planes that are written (fftp.fpp:824-919)
  if (launch_yfft(pn, d_B, d_B, nzp, nxh, nzp, +1, 1.0)) return 1;
  if (launch_x_c2r(pn, d_B, d_out, nzp, nzp, 1.0)) return 1;
  SX_CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, (size_t)nx * ny * nzp * sizeof(double), cudaMemcpyDeviceToHost, sn));
  SX_CUDA_CHECK(cudaStreamSynchronize(sn));
  g_boots_launches += po.launches + pn.launches;
  return 0;
}

static int read_all(const std::string& path, double* dst, size_t bytes) {
  const int fd = open(path.c_str(), O_RDONLY);
  SX_REQUIRE(fd >= 0, "io_read: cannot open file for reading: " + path);   // binary_io.f90:139-143
  size_t done = 0;
  while (done < bytes) {
    const ssize_t r = pread(fd, (char*)dst + done, bytes - done, (off_t)done);
    if (r <= 0) { close(fd); SX_REQUIRE(false, "io_read: file too short: " + path); }
    done += (size_t)r;
  }
  close(fd);
  return 0;
}

static int write_all(const std::string& path, const double* src, size_t bytes) {
  const int fd = open(path.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644);
  SX_REQUIRE(fd >= 0, "io_write: cannot open file for writing: " + path);
  size_t done = 0;
  while (done < bytes) {
    const ssize_t w = pwrite(fd, (const char*)src + done, bytes - done, (off_t)done);
    if (w <= 0) { close(fd); SX_REQUIRE(false, "io_write: short write to " + path); }
    done += (size_t)w;
  }
  close(fd);
  return 0;
}

static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) --b;
  return s.substr(a, b - a);
}

}  // namespace sx

using namespace sx;

extern "C" {

int sx_boots_points(int nzt, int nzp, int* Czt, int* Czn) {
  SX_REQUIRE(Czt && Czn, "null argument");
  return boots_points(nzt, nzp, Czt, Czn);
}

int sx_boots_regrid(int device, int nxt, int nyt, int nzt, int ozt, const char* tdir, int nx, int ny, int nzp,
                    const double* in_host, double* out_host) {
  SX_REQUIRE(in_host && out_host, "null argument");
  return boots_regrid(device, nxt, nyt, nzt, ozt, tdir, nx, ny, nzp, in_host, out_host);
}

int sx_boots_files(int device, const char* idir, const char* odir, const char* tdir, const char* fnlist, int nxt, int nyt,
                   int nzt, int ozt, int nx, int ny, int nzp) {
  SX_REQUIRE(idir && odir && fnlist, "null argument");
  SX_REQUIRE(nxt >= 1 && nyt >= 1 && nzt >= 1 && nx >= nxt && ny >= nyt && nzp >= nzt,
             "MAIN: prolongation specification incorrect");
  char suff[32];
  snprintf(suff, sizeof suff, "_P%05d-%05d-%05d", nx, ny, nzp);   // boots.fpp:174
  std::vector<double> vin((size_t)nxt * nyt * nzt), vout((size_t)nx * ny * nzp);
  const std::string list(fnlist);
  size_t ib = 0;
  while (ib < list.size()) {   // boots.fpp:228-240: names separated by ';'
    size_t ie = list.find(';', ib);
    if (ie == std::string::npos) ie = list.size();
    const std::string fname = trim(list.substr(ib, ie - ib));
    ib = ie + 1;
    if (fname.empty()) continue;
    if (read_all(std::string(idir) + "/" + fname, vin.data(), vin.size() * sizeof(double))) return 1;
    if (boots_regrid(device, nxt, nyt, nzt, ozt, tdir, nx, ny, nzp, vin.data(), vout.data())) return 1;
    if (write_all(std::string(odir) + "/" + fname + suff, vout.data(), vout.size() * sizeof(double))) return 1;
  }
  return 0;
}

unsigned long long sx_boots_launch_count(void) { return g_boots_launches; }

}  // extern "C"
