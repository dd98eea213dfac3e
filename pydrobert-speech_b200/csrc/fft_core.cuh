// fft_core.cuh -- in-register FFT building blocks shared by the device kernels and by the CPU
// emulation harness (csrc/emu_fft.cpp, compiled with g++ for the no-GPU test-suite).
//
// A real frame of N samples is transformed as a complex FFT of NC = N/2 points over
// z[n] = x[2n] + i x[2n+1], followed by the usual even/odd split.  The NC-point FFT is a
// two-stage Cooley-Tukey decomposition spread over G lanes (a sub-group of a warp):
//
//   n = G*n1 + n2   (lane n2 holds the R1 = NC/G inputs n1 = 0..R1-1)
//   stage 1 : Y[n2][k1] = sum_n1 z[G*n1+n2] W_R1^(n1 k1)            (R1-point DFT in registers)
//   twiddle : Y[n2][k1] *= W_NC^(n2 k1)
//   exchange: lane l' collects Y[.][k1] for k1 = l' + G*j           (through shared memory)
//   stage 2 : Z[k1 + R1*k2] = sum_n2 Y[n2][k1] W_G^(n2 k2)           (R1/G G-point DFTs)
//
// after which lane l' owns Z[k] for all k = l' (mod G), stored at register index m = (k-l')/G.
// The partner NC-k needed by the split lives in lane (G-l') mod G at index R1-1-m (lane 0: at
// index (R1-m) mod R1 of lane 0 itself).
#pragma once

// Arithmetic: on the device a complex value is ONE 64-bit register pair and every complex
// add / sub / multiply-add is a packed FP32x2 instruction (PTX add/sub/mul/fma.rn.f32x2 -> SASS
// FADD2 / FMUL2 / FFMA2, new on sm_100).  The FMA pipe retires the same number of lanes per clock
// as with scalar FFMA, but each packed instruction takes one issue slot instead of two, and the
// half swaps / sign flips a complex multiply needs are folded by ptxas into operand modifiers
// (R4.F32x2.LO_HI.NP ...), so a twiddle multiply is 2 instructions and a butterfly 2 -- the fft
// phase is issue-bound, so this is where the time goes (measured: tools/ubench/f32x2.cu).
// Compiled by g++ (csrc/emu_fft.cpp) the same templates run on plain float pairs.
#include <cstdint>

#ifdef __CUDACC__
#define PDS_HD __device__ __forceinline__
#else
#define PDS_HD inline
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
#endif

namespace pds {

// cos(2 pi m / 32) and sin(2 pi m / 32), m = 0..31
constexpr double kCos32[32] = {
    1.0,
    0.98078528040323044913,
    0.92387953251128675613,
    0.83146961230254523708,
    0.70710678118654752440,
    0.55557023301960222474,
    0.38268343236508977173,
    0.19509032201612826785,
    0.0,
    -0.19509032201612826785,
    -0.38268343236508977173,
    -0.55557023301960222474,
    -0.70710678118654752440,
    -0.83146961230254523708,
    -0.92387953251128675613,
    -0.98078528040323044913,
    -1.0,
    -0.98078528040323044913,
    -0.92387953251128675613,
    -0.83146961230254523708,
    -0.70710678118654752440,
    -0.55557023301960222474,
    -0.38268343236508977173,
    -0.19509032201612826785,
    0.0,
    0.19509032201612826785,
    0.38268343236508977173,
    0.55557023301960222474,
    0.70710678118654752440,
    0.83146961230254523708,
    0.92387953251128675613,
    0.98078528040323044913};

constexpr double kSin32[32] = {
    0.0,
    0.19509032201612826785,
    0.38268343236508977173,
    0.55557023301960222474,
    0.70710678118654752440,
    0.83146961230254523708,
    0.92387953251128675613,
    0.98078528040323044913,
    1.0,
    0.98078528040323044913,
    0.92387953251128675613,
    0.83146961230254523708,
    0.70710678118654752440,
    0.55557023301960222474,
    0.38268343236508977173,
    0.19509032201612826785,
    0.0,
    -0.19509032201612826785,
    -0.38268343236508977173,
    -0.55557023301960222474,
    -0.70710678118654752440,
    -0.83146961230254523708,
    -0.92387953251128675613,
    -0.98078528040323044913,
    -1.0,
    -0.98078528040323044913,
    -0.92387953251128675613,
    -0.83146961230254523708,
    -0.70710678118654752440,
    -0.55557023301960222474,
    -0.38268343236508977173,
    -0.19509032201612826785};

// ---- the complex value type -----------------------------------------------------------------
#ifdef __CUDACC__
struct cplx {
  unsigned long long v;  // {re (low half), im (high half)}: bit-compatible with float2 in memory
};
PDS_HD cplx cmake(float re, float im) {
  cplx r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im));
  return r;
}
PDS_HD float cre(cplx a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return lo;
}
PDS_HD float cim(cplx a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return hi;
}
PDS_HD cplx cadd(cplx a, cplx b) {
  cplx r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PDS_HD cplx csub(cplx a, cplx b) {
  cplx r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PDS_HD cplx cmul2(cplx a, cplx b) {  // element-wise: (a.re b.re, a.im b.im)
  cplx r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PDS_HD cplx cfma2(cplx a, cplx b, cplx c) {  // element-wise a * b + c
  cplx r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
#else
struct cplx {
  float x, y;
};
PDS_HD cplx cmake(float re, float im) {
  cplx r;
  r.x = re;
  r.y = im;
  return r;
}
PDS_HD float cre(cplx a) { return a.x; }
PDS_HD float cim(cplx a) { return a.y; }
PDS_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
PDS_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
PDS_HD cplx cmul2(cplx a, cplx b) { return cmake(a.x * b.x, a.y * b.y); }
PDS_HD cplx cfma2(cplx a, cplx b, cplx c) { return cmake(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif

PDS_HD cplx cswap(cplx a) { return cmake(cim(a), cre(a)); }       // folded into an operand modifier
PDS_HD float2 c2f(cplx a) { return make_float2(cre(a), cim(a)); }
PDS_HD cplx f2c(float2 a) { return cmake(a.x, a.y); }
// a * (wr + i wi) = (a.re wr - a.im wi, a.im wr + a.re wi): FMUL2 + FFMA2
PDS_HD cplx cmul(cplx a, float wr, float wi) {
  return cfma2(a, cmake(wr, wr), cmul2(cswap(a), cmake(-wi, wi)));
}
PDS_HD cplx cmul(cplx a, float2 w) { return cmul(a, w.x, w.y); }
// e + (-i) o and e - (-i) o, the radix-4 style butterfly with the trivial twiddle -i
PDS_HD cplx cadd_negi(cplx e, cplx o) { return cfma2(cswap(o), cmake(1.f, -1.f), e); }
PDS_HD cplx csub_negi(cplx e, cplx o) { return cfma2(cswap(o), cmake(-1.f, 1.f), e); }
// a + conj(b), a - conj(b)
PDS_HD cplx cadd_conj(cplx a, cplx b) { return cfma2(b, cmake(1.f, -1.f), a); }
PDS_HD cplx csub_conj(cplx a, cplx b) { return cfma2(b, cmake(-1.f, 1.f), a); }
PDS_HD float cnorm(cplx a) {  // |a|^2
  const cplx sq = cmul2(a, a);
  return cre(sq) + cim(sq);
}

// One radix-2 butterfly with twiddle W_R^K on the odd input: x_lo = e + W o, x_hi = e - W o
template <int R, int K>
PDS_HD void butterfly(cplx e, cplx o, cplx& x_lo, cplx& x_hi) {
  static_assert(R <= 32 && 32 % R == 0, "radix limited to 32");
  constexpr int idx = (K * (32 / R)) % 32;
  if constexpr (idx == 0) {
    x_lo = cadd(e, o);
    x_hi = csub(e, o);
  } else if constexpr (idx == 8) {  // W = -i
    x_lo = cadd_negi(e, o);
    x_hi = csub_negi(e, o);
  } else {
    constexpr float wr = (float)kCos32[idx];
    constexpr float wi = (float)(-kSin32[idx]);
    const cplx t = cmul(o, wr, wi);
    x_lo = cadd(e, t);
    x_hi = csub(e, t);
  }
}

template <int R, int K>
struct Butterflies {
  static PDS_HD void run(cplx (&x)[R], const cplx (&e)[R / 2], const cplx (&o)[R / 2]) {
    butterfly<R, K>(e[K], o[K], x[K], x[K + R / 2]);
    if constexpr (K + 1 < R / 2) Butterflies<R, K + 1>::run(x, e, o);
  }
};

// Compile-time bookkeeping for inputs known to be exact zeros (the zero padding of a frame that
// is shorter than the DFT): bit i of ZM set = input i is zero.
template <int R>
constexpr unsigned zmask_half(unsigned zm, int odd) {
  unsigned out = 0;
  for (int i = 0; i < R / 2; ++i)
    if (zm & (1u << (2 * i + odd))) out |= 1u << i;
  return out;
}
template <int R>
constexpr unsigned zmask_full() {
  return R >= 32 ? 0xffffffffu : ((1u << R) - 1u);
}

// In-place forward DFT of R points held in registers (natural order in, natural order out).
// Butterflies whose odd half is all zeros degenerate to copies (no instructions).
template <int R, unsigned ZM = 0u>
struct Dft {
  static PDS_HD void run(cplx (&x)[R]) {
    constexpr unsigned ZE = zmask_half<R>(ZM, 0), ZO = zmask_half<R>(ZM, 1);
    cplx e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) {
      e[i] = x[2 * i];
      o[i] = x[2 * i + 1];
    }
    Dft<R / 2, ZE>::run(e);
    if constexpr (ZO == zmask_full<R / 2>()) {
#pragma unroll
      for (int i = 0; i < R / 2; ++i) x[i] = x[i + R / 2] = e[i];
    } else {
      Dft<R / 2, ZO>::run(o);
      Butterflies<R, 0>::run(x, e, o);
    }
  }
};

template <unsigned ZM>
struct Dft<1, ZM> {
  static PDS_HD void run(cplx (&)[1]) {}
};

template <unsigned ZM>
struct Dft<2, ZM> {
  static PDS_HD void run(cplx (&x)[2]) {
    const cplx a = x[0], b = x[1];
    if constexpr ((ZM & 2u) != 0u) {
      x[1] = a;  // b == 0
    } else {
      x[0] = cadd(a, b);
      x[1] = csub(a, b);
    }
  }
};

// Geometry of the two-stage decomposition for a given real DFT size N.
template <int N>
struct FftGeom {
  static_assert(N >= 64 && N <= 2048 && (N & (N - 1)) == 0, "N must be a power of two in [64, 2048]");
  static constexpr int NC = N / 2;                                       // complex points
  static constexpr int G = NC >= 1024 ? 32 : (NC >= 256 ? 16 : (NC >= 64 ? 8 : 4));  // lanes/frame
  static constexpr int R1 = NC / G;                                      // registers per lane
  static constexpr int NSUB = R1 / G;                                    // stage-2 DFTs per lane
  static_assert(R1 % G == 0 && R1 <= 32, "unsupported split");
  static constexpr int SCR_STRIDE = R1 + 1;   // float2 units, +1 kills bank conflicts
  static constexpr int SCR_FLOAT2 = G * SCR_STRIDE;  // scratch per frame
};

// One unit of the real-FFT split: from Z[k] (= a) and Z[NC-k] (= b), with the window already
// scaled by 1/2, produce X[k] and X[NC-k] of the N-point real DFT.  w = e^{-2 pi i k / N}.
//   E = a + conj(b),  D = a - conj(b),  O = -i D,  T = w O = (-i w) D,
//   X[k] = E + T,  X[NC-k] = conj(E - T)   (the conjugate is irrelevant for |.|)
PDS_HD void split_pair(cplx a, cplx b, float2 w, cplx& xk, cplx& xq) {
  const cplx e = cadd_conj(a, b);
  const cplx d = csub_conj(a, b);
  const cplx t = cmul(d, w.y, -w.x);  // -i w = (w.im, -w.re)
  xk = cadd(e, t);
  xq = csub(e, t);
}

}  // namespace pds
