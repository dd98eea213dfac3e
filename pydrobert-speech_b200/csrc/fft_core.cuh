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

#ifdef __CUDACC__
#define PDS_HD __host__ __device__ __forceinline__
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

PDS_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PDS_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PDS_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// o * W_R^K with the trivial cases resolved at compile time (forward transform: W = e^{-2 pi i/R})
template <int R, int K>
PDS_HD float2 mul_twiddle(float2 o) {
  static_assert(R <= 32 && 32 % R == 0, "radix limited to 32");
  constexpr int idx = (K * (32 / R)) % 32;
  if constexpr (idx == 0) {
    return o;
  } else if constexpr (idx == 8) {  // -i
    return make_float2(o.y, -o.x);
  } else if constexpr (idx == 16) {  // -1
    return make_float2(-o.x, -o.y);
  } else if constexpr (idx == 24) {  // +i
    return make_float2(-o.y, o.x);
  } else if constexpr (idx == 4) {  // (1 - i)/sqrt2
    constexpr float c = (float)kCos32[4];
    return make_float2(c * (o.x + o.y), c * (o.y - o.x));
  } else if constexpr (idx == 12) {  // (-1 - i)/sqrt2
    constexpr float c = (float)kCos32[4];
    return make_float2(c * (o.y - o.x), -c * (o.x + o.y));
  } else {
    constexpr float wr = (float)kCos32[idx];
    constexpr float wi = (float)(-kSin32[idx]);
    return make_float2(o.x * wr - o.y * wi, o.x * wi + o.y * wr);
  }
}

template <int R, int K>
struct Butterflies {
  static PDS_HD void run(float2 (&x)[R], const float2 (&e)[R / 2], const float2 (&o)[R / 2]) {
    const float2 t = mul_twiddle<R, K>(o[K]);
    x[K] = cadd(e[K], t);
    x[K + R / 2] = csub(e[K], t);
    if constexpr (K + 1 < R / 2) Butterflies<R, K + 1>::run(x, e, o);
  }
};

// In-place forward DFT of R points held in registers (natural order in, natural order out).
template <int R>
struct Dft {
  static PDS_HD void run(float2 (&x)[R]) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) {
      e[i] = x[2 * i];
      o[i] = x[2 * i + 1];
    }
    Dft<R / 2>::run(e);
    Dft<R / 2>::run(o);
    Butterflies<R, 0>::run(x, e, o);
  }
};

template <>
struct Dft<1> {
  static PDS_HD void run(float2 (&)[1]) {}
};

template <>
struct Dft<2> {
  static PDS_HD void run(float2 (&x)[2]) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
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
PDS_HD void split_pair(float2 a, float2 b, float2 w, float2& xk, float2& xp) {
  const float er = a.x + b.x, ei = a.y - b.y;  // a + conj(b)
  const float dr = a.x - b.x, di = a.y + b.y;  // a - conj(b); O = -i * (dr, di) = (di, -dr)
  const float tr = w.x * di + w.y * dr;        // Re(w * O)
  const float ti = w.y * di - w.x * dr;        // Im(w * O)
  xk = make_float2(er + tr, ei + ti);
  xp = make_float2(er - tr, ti - ei);          // conj(E - T)
}

}  // namespace pds
