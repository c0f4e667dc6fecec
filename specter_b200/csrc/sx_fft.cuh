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
  // index distance of two elements n apart: n a multiple of 8, or both inside one group of 8
  __device__ __forceinline__ constexpr int stride(int n) const { return n + (n >> 3); }
};
template <int N> constexpr int sidx_elem_stride() { return N + N / 8; }
// pencil-fastest: NP pencils interleaved (conflict free when NP % 8 == 0)
// With np == 4 a quarter-warp covers two butterflies whose first-pass scatter addresses differ by a multiple
// of 128 B (2-way conflict); flipping the low element bit with bit 3 separates them and leaves the gathers and
// the second scatter conflict free -- but costs address arithmetic in the 80-register tile kernels and measured
// 3-15 % SLOWER on B200 (profiles/r1i_session4.md), so it is off unless built with -DSX_SWIZZLE.
struct SIdxPencil {
  int p, np;
  __device__ __forceinline__ int operator()(int e) const {
#ifdef SX_SWIZZLE
    if (np == 4) e ^= (e >> 3) & 1;
#endif
    return e * np + p;
  }
  __device__ __forceinline__ constexpr int stride(int n) const { return n * np; }
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

// ---- twiddles ------------------------------------------------------------------------------
// The twiddle of (pass P, butterfly b, leg m) is w1^m with w1 = exp(-2 pi i q /(Ns*r)), and q only
// depends on the thread (q = (j + b*T) mod Ns).  TwRegs keeps the w1 of every pass in registers
// (loaded once per kernel, N=512: two complex values) and the passes generate the powers with
// 21 flops instead of loading r-1 table entries: the table loads were ~17 % of the L1/shared
// pipe traffic of every kernel (profiles/r1h_knobs.md), the FP64 pipe has the headroom.
template <int N> struct TwSlots {
  typedef Fft1D<N> F;
  static constexpr int nb(int p) { return 8 / F::radix(p); }
  static constexpr int off(int p) { int o = 0; for (int q = 1; q < p; ++q) o += nb(q); return o; }
  static constexpr int count = off(F::npass) > 0 ? off(F::npass) : 1;
};
template <int N> struct TwRegs {
  cplx w[TwSlots<N>::count];
  const cplx* table;
  template <int P> __device__ __forceinline__ void load_pass(const cplx* __restrict__ tw, int j) {
    typedef Fft1D<N> F;
    if constexpr (P < F::npass) {
      constexpr int r = F::radix(P), nb = 8 / r, Ns = F::ns(P);
#pragma unroll
      for (int b = 0; b < nb; ++b) w[TwSlots<N>::off(P) + b] = __ldg(&tw[F::twoff(P) + ((j + b * F::T) & (Ns - 1))]);
      load_pass<P + 1>(tw, j);
    }
  }
  __device__ __forceinline__ void load(const cplx* __restrict__ tw, int j) {
    table = tw;
    load_pass<1>(tw, j);
  }
};
__device__ __forceinline__ cplx csqr(cplx a) { return cmake(fma(a.x, a.x, -(a.y * a.y)), (a.x + a.x) * a.y); }

// Hook: code run by every thread right before / after the first barrier of a transform (the point where the
// inputs of all threads have been consumed and the exchange buffer is about to be overwritten); the bulk-copy
// kernels use it to wait for an asynchronous store that still reads the buffer and to start the next load.
struct NoHook {
  __device__ __forceinline__ void before_sync() const {}
  __device__ __forceinline__ void after_sync() const {}
};

template <int N, int DIR, int P, class SI, class Hook>
__device__ __forceinline__ void fft_pass(cplx (&v)[8], const int j, cplx* s, const SI& si,
                                         const cplx* __restrict__ tw, const TwRegs<N>* twr, const Hook& hook) {
  typedef Fft1D<N> F;
  constexpr int r = F::radix(P), nb = 8 / r, Ns = F::ns(P), T = F::T;
  constexpr bool first = (P == 0), last = (P == F::npass - 1);
  constexpr int toff = F::twoff(P);
  // For T a multiple of 8 every element offset inside a pass is a multiple of 8 (or stays inside one group of 8),
  // so the shared-memory index is ONE per-thread base plus compile-time strides: the accesses become
  // [base + immediate] instead of one address register per butterfly leg.
#ifdef SX_SWIZZLE
  constexpr bool linear = false;
#else
  constexpr bool linear = (T % 8 == 0);
#endif
  if (!first) {
    if constexpr (linear) {
      const int g0 = si(j);
#pragma unroll
      for (int b = 0; b < nb; ++b) {
#pragma unroll
        for (int m = 0; m < r; ++m) v[b + m * nb] = s[g0 + si.stride(b * T + m * (N / r))];
      }
    } else {
#pragma unroll
      for (int b = 0; b < nb; ++b) {
#pragma unroll
        for (int m = 0; m < r; ++m) v[b + m * nb] = s[si(j + b * T + m * (N / r))];
      }
    }
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      if (twr != nullptr) {
        cplx w1 = twr->w[TwSlots<N>::off(P) + b];
        if (DIR > 0) w1.y = -w1.y;
        v[b + nb] = cmul(v[b + nb], w1);
        if constexpr (r >= 4) {
          const cplx w2 = csqr(w1), w3 = cmul(w2, w1);
          v[b + 2 * nb] = cmul(v[b + 2 * nb], w2);
          v[b + 3 * nb] = cmul(v[b + 3 * nb], w3);
          if constexpr (r == 8) {
            const cplx w4 = csqr(w2), w5 = cmul(w4, w1), w6 = csqr(w3), w7 = cmul(w4, w3);
            v[b + 4 * nb] = cmul(v[b + 4 * nb], w4);
            v[b + 5 * nb] = cmul(v[b + 5 * nb], w5);
            v[b + 6 * nb] = cmul(v[b + 6 * nb], w6);
            v[b + 7 * nb] = cmul(v[b + 7 * nb], w7);
          }
        }
      } else {
        const int q = (j + b * T) & (Ns - 1);
#pragma unroll
        for (int m = 1; m < r; ++m) {
          cplx w = __ldg(&tw[toff + (m - 1) * Ns + q]);
          if (DIR > 0) w.y = -w.y;
          v[b + m * nb] = cmul(v[b + m * nb], w);
        }
      }
    }
  }
  butterflies<r, DIR>(v);
  if (!last) {
    if (first) hook.before_sync();
    __syncthreads();  // WAR: everybody has finished reading this buffer
    if (first) hook.after_sync();
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int jb = j + b * T;
      const int j0 = (jb / Ns) * (Ns * r) + (jb & (Ns - 1));
      if constexpr (linear) {
        const int s0 = si(j0);
#pragma unroll
        for (int m = 0; m < r; ++m) s[s0 + si.stride(m * Ns)] = v[b + m * nb];
      } else {
#pragma unroll
        for (int m = 0; m < r; ++m) s[si(j0 + m * Ns)] = v[b + m * nb];
      }
    }
    __syncthreads();
  }
}

template <int N, int DIR, int P, class SI, class Hook>
__device__ __forceinline__ void fft_run(cplx (&v)[8], int j, cplx* s, const SI& si,
                                        const cplx* __restrict__ tw, const TwRegs<N>* twr, const Hook& hook) {
  fft_pass<N, DIR, P, SI, Hook>(v, j, s, si, tw, twr, hook);
  if constexpr (P + 1 < Fft1D<N>::npass) fft_run<N, DIR, P + 1, SI, Hook>(v, j, s, si, tw, twr, hook);
}

// v[k] <-> element j + k*T on entry and exit.  `s` is this CTA's exchange buffer; every
// thread of the CTA must call this together (it contains __syncthreads()).
template <int N, int DIR, class SI>
__device__ __forceinline__ void fft_regs(cplx (&v)[8], int j, cplx* s, const SI& si,
                                         const cplx* __restrict__ tw) {
#ifdef SX_NOFFT  // timing experiment only (tools/gpu_nofft.sh): data movement without the transforms
  (void)v; (void)j; (void)s; (void)si; (void)tw;
#else
  fft_run<N, DIR, 0, SI, NoHook>(v, j, s, si, tw, nullptr, NoHook());
#endif
}
// same, twiddles from registers (SX_TW_TABLE: build-time switch back to the table, for A/B timing)
template <int N, int DIR, class SI, class Hook = NoHook>
__device__ __forceinline__ void fft_regs(cplx (&v)[8], int j, cplx* s, const SI& si, const TwRegs<N>& twr,
                                         const Hook& hook = Hook()) {
#ifdef SX_NOFFT
  (void)v; (void)j; (void)s; (void)si; (void)twr;
  hook.before_sync();
  __syncthreads();
  hook.after_sync();
#elif defined(SX_TW_TABLE)
  fft_run<N, DIR, 0, SI, Hook>(v, j, s, si, twr.table, nullptr, hook);
#else
  fft_run<N, DIR, 0, SI, Hook>(v, j, s, si, twr.table, &twr, hook);
#endif
}

// ---- two transforms at once ------------------------------------------------------------------------------------
// Two independent pencils per thread (v, w: same length, possibly different directions) share every barrier and the
// twiddle powers of each pass: twice the independent work between two barriers and half the barriers per transform,
// which is what the transform-bound z-stage kernel needs at 8 warps per SM.  s0 / s1 are two exchange buffers.
template <int D, bool CONJ> __device__ __forceinline__ cplx tw_dir(cplx w) {   // table twiddles are the forward (-1) ones
  (void)D;
  return CONJ ? cmake(w.x, -w.y) : w;
}
template <int N, int D0, int D1, int P, class SI>
__device__ __forceinline__ void fft_pass2(cplx (&v)[8], cplx (&w)[8], const int j, cplx* s0, cplx* s1, const SI& si,
                                          const TwRegs<N>& twr) {
  typedef Fft1D<N> F;
  constexpr int r = F::radix(P), nb = 8 / r, Ns = F::ns(P), T = F::T;
  constexpr bool first = (P == 0), last = (P == F::npass - 1);
#ifdef SX_SWIZZLE
  constexpr bool linear = false;
#else
  constexpr bool linear = (T % 8 == 0);
#endif
  if (!first) {
    if constexpr (linear) {
      const int g0 = si(j);
#pragma unroll
      for (int b = 0; b < nb; ++b) {
#pragma unroll
        for (int m = 0; m < r; ++m) {
          v[b + m * nb] = s0[g0 + si.stride(b * T + m * (N / r))];
          w[b + m * nb] = s1[g0 + si.stride(b * T + m * (N / r))];
        }
      }
    } else {
#pragma unroll
      for (int b = 0; b < nb; ++b) {
#pragma unroll
        for (int m = 0; m < r; ++m) {
          v[b + m * nb] = s0[si(j + b * T + m * (N / r))];
          w[b + m * nb] = s1[si(j + b * T + m * (N / r))];
        }
      }
    }
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      // powers of the forward root; a backward transform multiplies by their conjugates
      const cplx w1 = twr.w[TwSlots<N>::off(P) + b];
      cplx pw[8];
      pw[1] = w1;
      if constexpr (r >= 4) {
        pw[2] = csqr(w1);
        pw[3] = cmul(pw[2], w1);
        if constexpr (r == 8) {
          pw[4] = csqr(pw[2]);
          pw[5] = cmul(pw[4], w1);
          pw[6] = csqr(pw[3]);
          pw[7] = cmul(pw[4], pw[3]);
        }
      }
#pragma unroll
      for (int m = 1; m < r; ++m) {
        v[b + m * nb] = cmul(v[b + m * nb], tw_dir<D0, (D0 > 0)>(pw[m]));
        w[b + m * nb] = cmul(w[b + m * nb], tw_dir<D1, (D1 > 0)>(pw[m]));
      }
    }
  }
  butterflies<r, D0>(v);
  butterflies<r, D1>(w);
  if (!last) {
    __syncthreads();  // WAR: everybody has finished reading both buffers
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int jb = j + b * T;
      const int j0 = (jb / Ns) * (Ns * r) + (jb & (Ns - 1));
      if constexpr (linear) {
        const int q0 = si(j0);
#pragma unroll
        for (int m = 0; m < r; ++m) {
          s0[q0 + si.stride(m * Ns)] = v[b + m * nb];
          s1[q0 + si.stride(m * Ns)] = w[b + m * nb];
        }
      } else {
#pragma unroll
        for (int m = 0; m < r; ++m) {
          s0[si(j0 + m * Ns)] = v[b + m * nb];
          s1[si(j0 + m * Ns)] = w[b + m * nb];
        }
      }
    }
    __syncthreads();
  }
}
template <int N, int D0, int D1, int P, class SI>
__device__ __forceinline__ void fft_run2(cplx (&v)[8], cplx (&w)[8], int j, cplx* s0, cplx* s1, const SI& si,
                                         const TwRegs<N>& twr) {
  fft_pass2<N, D0, D1, P, SI>(v, w, j, s0, s1, si, twr);
  if constexpr (P + 1 < Fft1D<N>::npass) fft_run2<N, D0, D1, P + 1, SI>(v, w, j, s0, s1, si, twr);
}
// v, w <-> elements j + k*T on entry and exit; every thread of the CTA calls this together
template <int N, int D0, int D1, class SI>
__device__ __forceinline__ void fft_regs2(cplx (&v)[8], cplx (&w)[8], int j, cplx* s0, cplx* s1, const SI& si,
                                          const TwRegs<N>& twr) {
#ifdef SX_NOFFT
  (void)v; (void)w; (void)j; (void)s0; (void)s1; (void)si; (void)twr;
  __syncthreads();
#else
  fft_run2<N, D0, D1, 0, SI>(v, w, j, s0, s1, si, twr);
#endif
}

// host: twiddle table for Fft1D<N> (forward sign), long-double accurate
void build_twiddles(int N, cplx* out, int* count);
int twiddle_count(int N);

}  // namespace sx
