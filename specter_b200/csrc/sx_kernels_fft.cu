// Batched FP64 FFT stages: z pencils (with fused FC-Gram continuation), y lines and
// x r2c / c2r lines, all operating on z-fastest arrays so no explicit transpose
// kernel exists (the reference's blocked transposes, fftp/fftp.fpp:505-519,866-880,
// become strided tile loads with >=128 B contiguous segments).
#include "sx_fft.cuh"
#include "sx_plan.h"

namespace sx {

// ---------------------------------------------------------------------------------
// z pencils.  Layout (nz, ny, nxl), z contiguous.  Replaces fftp1d_real_to_complex_z
// (fftp.fpp:720-786, continuation :757-772 fused into the load) and
// fftp1d_complex_to_real_z (:1060-1094).
// ---------------------------------------------------------------------------------
struct ZArgs {
  const cplx* in;
  cplx* out;
  long npencils;
  int C, d;
  const double* dir;  // [C][d] row-major
  double scale_phys, scale_cont;
};

constexpr int kMaxD = 10;

// FC-Gram continuation of one pencil held in registers: bnd[0..d-1] = f(1..d),
// bnd[d..2d-1] = f(n-C-d+1..n-C) (1-based reference indices).
template <int N>
__device__ __forceinline__ void fc_continue_regs(cplx (&v)[8], int j, const cplx* bnd, int C, int d,
                                                 const double* __restrict__ dir) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e >= N - C) {
      const int ii = e - (N - C);
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
__device__ __forceinline__ void stash_boundary(const cplx (&v)[8], int j, cplx* bnd, int C, int d) {
  constexpr int T = N / 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    if (e < d) bnd[e] = v[k];
    if (e >= N - C - d && e < N - C) bnd[d + (e - (N - C - d))] = v[k];
  }
}

template <int N, int DIR, bool CONT>
__global__ void __launch_bounds__(N / 8 >= 256 ? N / 8 : 256)
zfft_kernel(ZArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8;
  const int ppb = blockDim.x / T;
  const int pl = threadIdx.x / T, j = threadIdx.x % T;
  const long pencil = (long)blockIdx.x * ppb + pl;
  const bool active = pencil < a.npencils;
  const SIdxElem si{pl * sidx_elem_stride<N>()};
  cplx v[8];
  const cplx* src = a.in + pencil * N;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = j + k * T;
    v[k] = (active && (!CONT || e < N - a.C)) ? src[e] : cmake(0.0, 0.0);
  }
  if (CONT) {
    cplx* bnd = smem + ppb * sidx_elem_stride<N>() + pl * 2 * kMaxD;
    stash_boundary<N>(v, j, bnd, a.C, a.d);
    __syncthreads();
    fc_continue_regs<N>(v, j, bnd, a.C, a.d, a.dir);
  }
  fft_regs<N, DIR>(v, j, smem, si, tw);
  if (active) {
    cplx* dst = a.out + pencil * N;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = j + k * T;
      dst[e] = cscale(v[k], e < N - a.C ? a.scale_phys : a.scale_cont);
    }
  }
}

template <int N> static int zfft_dispatch(Plan& p, const ZArgs& a, int dir, bool cont) {
  constexpr int T = N / 8;
  const int ppb = T >= 256 ? 1 : 256 / T;
  const int threads = ppb * T;
  const size_t smem = ((size_t)ppb * sidx_elem_stride<N>() + (size_t)ppb * 2 * kMaxD) * sizeof(cplx);
  const unsigned grid = (unsigned)((a.npencils + ppb - 1) / ppb);
  if (grid == 0) return 0;
  const cplx* tw = p.tw_z;
  cudaStream_t st = p.stream;
  if (stage_mark(p, ST_ZFFT)) return 1;
#define SX_Z_CASE(D, CT)                                                                          \
  {                                                                                               \
    auto kfn = zfft_kernel<N, D, CT>;                                                             \
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    SX_LAUNCH(kfn, dim3(grid), dim3(threads), smem, st, a, tw);                        \
  }
  if (dir < 0 && cont) SX_Z_CASE(-1, true)
  else if (dir < 0) SX_Z_CASE(-1, false)
  else SX_Z_CASE(1, false)
#undef SX_Z_CASE
  p.launches++;
  SX_KERNEL_CHECK();
  return 0;
}

bool fft_size_supported(int n, bool zdir) {
  if (n < 16 || (n & (n - 1))) return false;
  return n <= (zdir ? 4096 : 2048);
}

int launch_zfft(Plan& p, const cplx* in, cplx* out, long npencils, int dir, bool cont,
                double scale_phys, double scale_cont) {
  SX_REQUIRE(!cont || (p.Cz > 0 && p.oz > 0 && p.oz <= kMaxD), "continuation needs 0 < d <= 10");
  ZArgs a{in, out, npencils, p.Cz, p.oz, p.d_dir, scale_phys, scale_cont};
  switch (p.nz) {
    case 16: return zfft_dispatch<16>(p, a, dir, cont);
    case 32: return zfft_dispatch<32>(p, a, dir, cont);
    case 64: return zfft_dispatch<64>(p, a, dir, cont);
    case 128: return zfft_dispatch<128>(p, a, dir, cont);
    case 256: return zfft_dispatch<256>(p, a, dir, cont);
    case 512: return zfft_dispatch<512>(p, a, dir, cont);
    case 1024: return zfft_dispatch<1024>(p, a, dir, cont);
    case 2048: return zfft_dispatch<2048>(p, a, dir, cont);
    case 4096: return zfft_dispatch<4096>(p, a, dir, cont);
  }
  SX_REQUIRE(false, "unsupported nz (power of two in [16,4096] required)");
}

// ---------------------------------------------------------------------------------
// y lines on a (nzc, ny, nxc) z-fastest array: a CTA transforms NP adjacent z's of
// one kx at once (NP*16 B contiguous per row).  Replaces the y half of the FFTW 2-D
// plans (fftp.fpp:99-103).
// ---------------------------------------------------------------------------------
struct YArgs {
  const cplx* in;
  cplx* out;
  int nzc, nxc, nz_active;
  double scale;
};

template <int N> struct TileNP {  // z's (or z-pairs) per CTA for the y and x stages
  static constexpr int value = N <= 64 ? 32 : (N == 128 ? 16 : (N <= 1024 ? 8 : 4));
};

template <int N, int DIR, int NP>
__global__ void __launch_bounds__(NP * (N / 8)) yfft_kernel(YArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8;
  const int p = threadIdx.x % NP, j = threadIdx.x / NP;
  const int z = blockIdx.x * NP + p;
  const bool active = z < a.nz_active;
  const size_t base = (size_t)blockIdx.y * N * a.nzc + z;
  cplx v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    v[k] = active ? a.in[base + (size_t)(j + k * T) * a.nzc] : cmake(0.0, 0.0);
  fft_regs<N, DIR>(v, j, smem, SIdxPencil{p, NP}, tw);
  if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a.out[base + (size_t)(j + k * T) * a.nzc] = cscale(v[k], a.scale);
  }
}

template <int N> static int yfft_dispatch(Plan& p, const YArgs& a, int dir) {
  constexpr int NP = TileNP<N>::value;
  const int threads = NP * (N / 8);
  const size_t smem = (size_t)NP * N * sizeof(cplx);
  dim3 grid(cdiv(a.nz_active, NP), a.nxc);
  if (grid.x == 0 || grid.y == 0) return 0;
  const cplx* tw = p.tw_y;
  cudaStream_t st = p.stream;
  if (stage_mark(p, ST_YFFT)) return 1;
  if (dir < 0) {
    auto kfn = yfft_kernel<N, -1, NP>;
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SX_LAUNCH(kfn, grid, dim3(threads), smem, st, a, tw);
  } else {
    auto kfn = yfft_kernel<N, 1, NP>;
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SX_LAUNCH(kfn, grid, dim3(threads), smem, st, a, tw);
  }
  p.launches++;
  SX_KERNEL_CHECK();
  return 0;
}

int launch_yfft(Plan& p, const cplx* in, cplx* out, int nzc, int nxc, int nz_active, int dir,
                double scale) {
  YArgs a{in, out, nzc, nxc, nz_active, scale};
  switch (p.ny) {
    case 16: return yfft_dispatch<16>(p, a, dir);
    case 32: return yfft_dispatch<32>(p, a, dir);
    case 64: return yfft_dispatch<64>(p, a, dir);
    case 128: return yfft_dispatch<128>(p, a, dir);
    case 256: return yfft_dispatch<256>(p, a, dir);
    case 512: return yfft_dispatch<512>(p, a, dir);
    case 1024: return yfft_dispatch<1024>(p, a, dir);
    case 2048: return yfft_dispatch<2048>(p, a, dir);
  }
  SX_REQUIRE(false, "unsupported ny (power of two in [16,2048] required)");
}

// ---------------------------------------------------------------------------------
// x lines.  Two real lines (adjacent z) ride one complex FFT of length nx:
//   c2r: Z(k) = A(k) + i B(k) on the Hermitian-completed spectrum, z = IFFT(Z) = a + i b
//   r2c: Z = FFT(a + i b), A(k) = (Z(k)+conj Z(N-k))/2, B(k) = (Z(k)-conj Z(N-k))/(2i)
// FFTW's c2r ignores Im of the kx=0 and kx=nx/2 entries; so do we (explicitly zeroed).
// Spectral side: (nzc, ny, nxh) z-fastest.  Real side: (nx, ny, nzc) x-fastest.
// Replaces the x half of the FFTW 2-D r2c/c2r plans (fftp.fpp:99-103, 469, 913).
// ---------------------------------------------------------------------------------
struct XArgs {
  const cplx* spec_in;
  cplx* spec_out;
  const double* real_in;
  double* real_out;
  int nzc, ny, nz_active;
  double scale;
};

template <int N, int NP>
__global__ void __launch_bounds__(NP * (N / 8)) xfft_c2r_kernel(XArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8;
  const int p = threadIdx.x % NP, t = threadIdx.x / NP;
  const int jy = blockIdx.y;
  const int z0 = blockIdx.x * 2 * NP;
  const int zA = z0 + 2 * p, zB = zA + 1;
  cplx v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = t + k * T;
    const int kx = e <= N / 2 ? e : N - e;
    const size_t idx = ((size_t)kx * a.ny + jy) * a.nzc;
    cplx A = zA < a.nz_active ? a.spec_in[idx + zA] : cmake(0.0, 0.0);
    cplx B = zB < a.nz_active ? a.spec_in[idx + zB] : cmake(0.0, 0.0);
    if (kx == 0 || kx == N / 2) { A.y = 0.0; B.y = 0.0; }
    if (e > N / 2) { A.y = -A.y; B.y = -B.y; }
    v[k] = cmake(A.x - B.y, A.y + B.x);
  }
  fft_regs<N, 1>(v, t, smem, SIdxPencil{p, NP}, tw);
  // transpose through shared memory so the real rows are written x-contiguous
  double* sr = reinterpret_cast<double*>(smem);
  constexpr int RS = N + 1;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = t + k * T;
    sr[(2 * p) * RS + x] = v[k].x * a.scale;
    sr[(2 * p + 1) * RS + x] = v[k].y * a.scale;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * NP * N; idx += blockDim.x) {
    const int row = idx / N, x = idx % N;
    const int z = z0 + row;
    if (z < a.nz_active) a.real_out[((size_t)z * a.ny + jy) * N + x] = sr[row * RS + x];
  }
}

template <int N, int NP>
__global__ void __launch_bounds__(NP * (N / 8)) xfft_r2c_kernel(XArgs a, const cplx* __restrict__ tw) {
  SX_DYN_SMEM(cplx, smem);
  constexpr int T = N / 8;
  const int p = threadIdx.x % NP, t = threadIdx.x / NP;
  const int jy = blockIdx.y;
  const int z0 = blockIdx.x * 2 * NP;
  const int zA = z0 + 2 * p, zB = zA + 1;
  double* sr = reinterpret_cast<double*>(smem);
  constexpr int RS = N + 1;
  for (int idx = threadIdx.x; idx < 2 * NP * N; idx += blockDim.x) {
    const int row = idx / N, x = idx % N;
    const int z = z0 + row;
    sr[row * RS + x] = z < a.nz_active ? a.real_in[((size_t)z * a.ny + jy) * N + x] : 0.0;
  }
  __syncthreads();
  cplx v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = t + k * T;
    v[k] = cmake(sr[(2 * p) * RS + x], sr[(2 * p + 1) * RS + x]);
  }
  const SIdxPencil si{p, NP};
  fft_regs<N, -1>(v, t, smem, si, tw);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) smem[si(t + k * T)] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int kk = t + k * T;
    if (kk <= N / 2) {
      const cplx Zk = v[k];
      const cplx Zn = smem[si((N - kk) & (N - 1))];
      const cplx A = cmake(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
      const cplx B = cmake(0.5 * (Zk.y + Zn.y), -0.5 * (Zk.x - Zn.x));
      const size_t idx = ((size_t)kk * a.ny + jy) * a.nzc;
      if (zA < a.nz_active) a.spec_out[idx + zA] = cscale(A, a.scale);
      if (zB < a.nz_active) a.spec_out[idx + zB] = cscale(B, a.scale);
    }
  }
}

template <int N> static int xfft_dispatch(Plan& p, const XArgs& a, bool c2r) {
  constexpr int NP = TileNP<N>::value;
  const int threads = NP * (N / 8);
  const size_t smem = (size_t)NP * N * sizeof(cplx) + 2 * NP * sizeof(double) + 16;
  dim3 grid(cdiv(a.nz_active, 2 * NP), a.ny);
  if (grid.x == 0) return 0;
  const cplx* tw = p.tw_x;
  cudaStream_t st = p.stream;
  if (stage_mark(p, ST_XFFT)) return 1;
  if (c2r) {
    auto kfn = xfft_c2r_kernel<N, NP>;
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SX_LAUNCH(kfn, grid, dim3(threads), smem, st, a, tw);
  } else {
    auto kfn = xfft_r2c_kernel<N, NP>;
    SX_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SX_LAUNCH(kfn, grid, dim3(threads), smem, st, a, tw);
  }
  p.launches++;
  SX_KERNEL_CHECK();
  return 0;
}

static int xfft_switch(Plan& p, const XArgs& a, bool c2r) {
  switch (p.nx) {
    case 16: return xfft_dispatch<16>(p, a, c2r);
    case 32: return xfft_dispatch<32>(p, a, c2r);
    case 64: return xfft_dispatch<64>(p, a, c2r);
    case 128: return xfft_dispatch<128>(p, a, c2r);
    case 256: return xfft_dispatch<256>(p, a, c2r);
    case 512: return xfft_dispatch<512>(p, a, c2r);
    case 1024: return xfft_dispatch<1024>(p, a, c2r);
    case 2048: return xfft_dispatch<2048>(p, a, c2r);
  }
  SX_REQUIRE(false, "unsupported nx (power of two in [16,2048] required)");
}

int launch_x_c2r(Plan& p, const cplx* spec, double* real, int nzc, int nz_active, double scale) {
  XArgs a{spec, nullptr, nullptr, real, nzc, p.ny, nz_active, scale};
  return xfft_switch(p, a, true);
}
int launch_x_r2c(Plan& p, const double* real, cplx* spec, int nzc, int nz_active, double scale) {
  XArgs a{nullptr, spec, real, nullptr, nzc, p.ny, nz_active, scale};
  return xfft_switch(p, a, false);
}

// ---------------------------------------------------------------------------------
// twiddles (host, long double)
// ---------------------------------------------------------------------------------
template <int N> static void tw_fill(cplx* out) {
  typedef Fft1D<N> F;
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int p = 1; p < F::npass; ++p) {
    const int r = F::radix(p), Ns = F::ns(p), M = Ns * r;
    for (int m = 1; m < r; ++m)
      for (int q = 0; q < Ns; ++q) {
        const int num = (m * q) % M;
        const long double ang = -two_pi * (long double)num / (long double)M;
        out[F::twoff(p) + (m - 1) * Ns + q] = cmake((double)cosl(ang), (double)sinl(ang));
      }
  }
}

int twiddle_count(int N) {
  switch (N) {
    case 16: return Fft1D<16>::twsize;
    case 32: return Fft1D<32>::twsize;
    case 64: return Fft1D<64>::twsize;
    case 128: return Fft1D<128>::twsize;
    case 256: return Fft1D<256>::twsize;
    case 512: return Fft1D<512>::twsize;
    case 1024: return Fft1D<1024>::twsize;
    case 2048: return Fft1D<2048>::twsize;
    case 4096: return Fft1D<4096>::twsize;
  }
  return -1;
}

void build_twiddles(int N, cplx* out, int* count) {
  *count = twiddle_count(N);
  switch (N) {
    case 16: tw_fill<16>(out); break;
    case 32: tw_fill<32>(out); break;
    case 64: tw_fill<64>(out); break;
    case 128: tw_fill<128>(out); break;
    case 256: tw_fill<256>(out); break;
    case 512: tw_fill<512>(out); break;
    case 1024: tw_fill<1024>(out); break;
    case 2048: tw_fill<2048>(out); break;
    case 4096: tw_fill<4096>(out); break;
  }
}

}  // namespace sx
