// emu_fft.cpp -- CPU emulation of the fft phase of stft_fused_kernel, lane by lane.
//
// Test infrastructure only (built by tests/ with g++, never linked into libpds_b200.so): it runs
// the SAME templates (fft_core.cuh) with the SAME index maps as the device code -- stage-1 load
// pattern, twiddles, the padded exchange layout, the partner/shuffle pairing of the real-FFT
// split -- so the no-GPU test-suite can check them against numpy.fft.rfft.
#include <cmath>
#include <cstring>
#include <vector>

#include "fft_core.cuh"

using namespace pds;

template <int N>
static void emulate(const float* frame, int L, const float* window, float* P, int power) {
  using Geo = FftGeom<N>;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1, NSUB = Geo::NSUB;
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<float> win(N, 0.f), x(N, 0.f);
  for (int i = 0; i < L; ++i) win[i] = 0.5f * window[i], x[i] = frame[i];
  std::vector<cplx> scr(Geo::SCR_FLOAT2);
  std::vector<std::vector<cplx>> Z(G, std::vector<cplx>(R1));
  const int rows_full = L / (2 * G), row_partial = (L % (2 * G)) != 0;
  for (int l = 0; l < G; ++l) {  // stage 1 of every lane
    cplx z[R1];
    for (int r = 0; r < R1; ++r) {
      if (r < rows_full) {
        z[r] = cmul2(cmake(x[2 * (G * r + l)], x[2 * (G * r + l) + 1]),
                     cmake(win[2 * (G * r + l)], win[2 * (G * r + l) + 1]));
      } else if (r == rows_full && row_partial) {
        const int i0 = 2 * (G * r + l);
        const float x0 = i0 < L ? x[i0] : 0.f, x1 = i0 + 1 < L ? x[i0 + 1] : 0.f;
        z[r] = cmake(x0 * win[i0], x1 * win[i0 + 1]);
      } else {
        z[r] = cmake(0.f, 0.f);
      }
    }
    // like the kernels' kRows13 mode: the last 3/16 of the rows are known zeros
    constexpr int ROWS13 = (R1 * 13) / 16;
    constexpr unsigned Z13 = zmask_full<R1>() & ~((1u << ROWS13) - 1u);
    if (rows_full + row_partial == ROWS13)
      Dft<R1, Z13>::run(z);
    else
      Dft<R1>::run(z);
    for (int k1 = 1; k1 < R1; ++k1) {
      const double a = -two_pi * (double)((long long)l * k1 % NC) / NC;
      z[k1] = cmul(z[k1], make_float2((float)std::cos(a), (float)std::sin(a)));
    }
    for (int k1 = 0; k1 < R1; ++k1) scr[l * Geo::SCR_STRIDE + k1] = z[k1];
  }
  for (int l = 0; l < G; ++l) {  // stage 2 of every lane
    for (int j = 0; j < NSUB; ++j) {
      cplx v[G];
      for (int n2 = 0; n2 < G; ++n2) v[n2] = scr[n2 * Geo::SCR_STRIDE + l + G * j];
      Dft<G>::run(v);
      for (int k2 = 0; k2 < G; ++k2) Z[l][j + NSUB * k2] = v[k2];
    }
  }
  for (int l = 0; l < G; ++l) {  // split
    const int partner = (G - l) % G;
    for (int m = 0; m < R1 / 2; ++m) {
      cplx b = Z[partner][R1 - 1 - m];  // the shuffle
      if (l == 0) b = Z[0][(R1 - m) % R1];
      const double a = -two_pi * (double)(l + G * m) / N;
      cplx xk, xq;
      split_pair(Z[l][m], b, make_float2((float)std::cos(a), (float)std::sin(a)), xk, xq);
      float pk = cnorm(xk), pq = cnorm(xq);
      if (!power) pk = std::sqrt(pk), pq = std::sqrt(pq);
      const int k = l + G * m;
      P[k] = pk;
      P[NC - k] = pq;
    }
    if (l == 0) {
      cplx xk, xq;
      split_pair(Z[0][R1 / 2], Z[0][R1 / 2], make_float2(0.f, -1.f), xk, xq);
      float pk = cnorm(xk);
      if (!power) pk = std::sqrt(pk);
      P[NC / 2] = pk;
    }
  }
}

extern "C" int pds_emu_frame_spectrum(int N, const float* frame, int L, const float* window,
                                      float* P, int power) {
  // P must hold N/2+1 floats, pre-filled by the caller with NaN to detect unwritten bins
  switch (N) {
    case 128: emulate<128>(frame, L, window, P, power); return 0;
    case 256: emulate<256>(frame, L, window, P, power); return 0;
    case 512: emulate<512>(frame, L, window, P, power); return 0;
    case 1024: emulate<1024>(frame, L, window, P, power); return 0;
    case 2048: emulate<2048>(frame, L, window, P, power); return 0;
    default: return -1;
  }
}
