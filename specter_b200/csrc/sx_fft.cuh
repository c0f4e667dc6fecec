// Register-resident FP64 Stockham FFT engine (power-of-two N = 8*T, radices 8/4/2).
//
// Every thread owns 8 complex values.  On entry thread j of a pencil holds elements
// {j + k*T, k=0..7}; on exit it holds the transform at the SAME element indices, so
// first-pass loads and last-pass stores go straight to/from global memory (coalesced)
// and elementwise work composes with the FFT without re-distribution.  Passes
// exchange data through shared memory with conflict-free indexing (SIdx*).
// Replaces the FFTW executes of the reference (fftp/fftp.fpp:469,610,780,913,1089).
#pragma once
#include "sx_common.cuh"

namespace sx {

// ---- decomposition: as many radix-8 passes as possible, remainder (2|4) last ----
template <int N> struct Fft1D {
  static_assert(N >= 16 && (N & (N - 1)) == 0, "N must be a power of two >= 16");
  static constexpr int T = N / 8;
  static constexpr int log2n() { int l = 0, n = N; while (n > 1) { n >>= 1; ++l; } return l; }
  static constexpr int n8 = log2n() / 3;
  static constexpr int rem = 1 << (log2n() % 3);  // 1, 2 or 4
  static constexpr int npass = n8 + (rem > 1 ? 1 : 0);
  static constexpr int radix(int p) { return p < n8 ? 8 : rem; }
  static constexpr int ns(int p) { int s = 1; for (int q = 0; q < p; ++q) s *= radix(q); return s; }
  // twiddle table: for pass p>=1, entries [(m-1)*Ns + q] = exp(-2 pi i m q /(Ns*r)), m=1..r-1
  static constexpr int twoff(int p) { int o = 0; for (int q = 1; q < p; ++q) o += (radix(q) - 1) * ns(q); return o; }
  static constexpr int twsize = twoff(npass);
};

// shared-memory indexers --------------------------------------------------------
// element-fastest with one pad slot every 8 elements (16 B elements => a quarter-warp
// of consecutive or stride-8 element accesses is bank-conflict free)
struct SIdxElem {
  int base;
  __device__ __forceinline__ int operator()(int e) const { return base + e + (e >> 3); }
};
template <int N> constexpr int sidx_elem_stride() { return N + N / 8; }
// pencil-fastest: NP pencils interleaved (conflict free when NP % 8 == 0)
struct SIdxPencil {
  int p, np;
  __device__ __forceinline__ int operator()(int e) const { return e * np + p; }
};

// ---- small in-register DFTs (natural order in/out) ---------------------------------
template <int DIR> __device__ __forceinline__ cplx mul_pm_i(cplx a) {  // a * (DIR*i)
  return DIR > 0 ? cmuli(a) : cmulmi(a);
}
template <int DIR> __device__ __forceinline__ void dft2(cplx& a, cplx& b) {
  cplx t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
template <int DIR> __device__ __forceinline__ void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
  cplx t0 = cadd(x0, x2), t1 = csub(x0, x2);
  cplx t2 = cadd(x1, x3), t3 = mul_pm_i<DIR>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}
template <int DIR> __device__ __forceinline__ void dft8(cplx (&x)[8]) {
  const double h = 0.70710678118654752440;
  cplx a0 = cadd(x[0], x[4]), b0 = csub(x[0], x[4]);
  cplx a1 = cadd(x[1], x[5]), b1 = csub(x[1], x[5]);
  cplx a2 = cadd(x[2], x[6]), b2 = csub(x[2], x[6]);
  cplx a3 = cadd(x[3], x[7]), b3 = csub(x[3], x[7]);
  // b_n *= w8^n, w8 = exp(DIR*i*pi/4)
  if (DIR < 0) {
    b1 = cmake((b1.x + b1.y) * h, (b1.y - b1.x) * h);
    b2 = cmulmi(b2);
    b3 = cmake((b3.y - b3.x) * h, -(b3.x + b3.y) * h);
  } else {
    b1 = cmake((b1.x - b1.y) * h, (b1.x + b1.y) * h);
    b2 = cmuli(b2);
    b3 = cmake(-(b3.x + b3.y) * h, (b3.x - b3.y) * h);
  }
  dft4<DIR>(a0, a1, a2, a3);
  dft4<DIR>(b0, b1, b2, b3);
  x[0] = a0; x[2] = a1; x[4] = a2; x[6] = a3;
  x[1] = b0; x[3] = b1; x[5] = b2; x[7] = b3;
}

template <int R, int DIR> __device__ __forceinline__ void butterflies(cplx (&v)[8]) {
  // slot of (butterfly b, leg m) is b + m*(8/R)
  if (R == 8) {
    dft8<DIR>(v);
  } else if (R == 4) {
    dft4<DIR>(v[0], v[2], v[4], v[6]);
    dft4<DIR>(v[1], v[3], v[5], v[7]);
  } else {
    dft2<DIR>(v[0], v[4]);
    dft2<DIR>(v[1], v[5]);
    dft2<DIR>(v[2], v[6]);
    dft2<DIR>(v[3], v[7]);
  }
}

template <int N, int DIR, int P, class SI>
__device__ __forceinline__ void fft_pass(cplx (&v)[8], const int j, cplx* s, const SI& si,
                                         const cplx* __restrict__ tw) {
  typedef Fft1D<N> F;
  constexpr int r = F::radix(P), nb = 8 / r, Ns = F::ns(P), T = F::T;
  constexpr bool first = (P == 0), last = (P == F::npass - 1);
  constexpr int toff = F::twoff(P);
  if (!first) {
#pragma unroll
    for (int b = 0; b < nb; ++b) {
#pragma unroll
      for (int m = 0; m < r; ++m) v[b + m * nb] = s[si(j + b * T + m * (N / r))];
    }
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int q = (j + b * T) & (Ns - 1);
#pragma unroll
      for (int m = 1; m < r; ++m) {
        cplx w = __ldg(&tw[toff + (m - 1) * Ns + q]);
        if (DIR > 0) w.y = -w.y;
        v[b + m * nb] = cmul(v[b + m * nb], w);
      }
    }
  }
  butterflies<r, DIR>(v);
  if (!last) {
    __syncthreads();  // WAR: everybody has finished reading this buffer
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int jb = j + b * T;
      const int j0 = (jb / Ns) * (Ns * r) + (jb & (Ns - 1));
#pragma unroll
      for (int m = 0; m < r; ++m) s[si(j0 + m * Ns)] = v[b + m * nb];
    }
    __syncthreads();
  }
}

template <int N, int DIR, int P, class SI>
__device__ __forceinline__ void fft_run(cplx (&v)[8], int j, cplx* s, const SI& si,
                                        const cplx* __restrict__ tw) {
  fft_pass<N, DIR, P, SI>(v, j, s, si, tw);
  if constexpr (P + 1 < Fft1D<N>::npass) fft_run<N, DIR, P + 1, SI>(v, j, s, si, tw);
}

// v[k] <-> element j + k*T on entry and exit.  `s` is this CTA's exchange buffer; every
// thread of the CTA must call this together (it contains __syncthreads()).
template <int N, int DIR, class SI>
__device__ __forceinline__ void fft_regs(cplx (&v)[8], int j, cplx* s, const SI& si,
                                         const cplx* __restrict__ tw) {
#ifdef SX_NOFFT  // timing experiment only (tools/gpu_nofft.sh): data movement without the transforms
  (void)v; (void)j; (void)s; (void)si; (void)tw;
#else
  fft_run<N, DIR, 0, SI>(v, j, s, si, tw);
#endif
}

// host: twiddle table for Fft1D<N> (forward sign), long-double accurate
void build_twiddles(int N, cplx* out, int* count);
int twiddle_count(int N);

}  // namespace sx
