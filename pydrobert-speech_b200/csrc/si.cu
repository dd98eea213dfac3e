// si.cu -- short-integration frame computer (K3 of SURVEY.md).
//
// Reference semantics (compute.py:613-999, spec: tests/test_compute.py:129-176): zero-pad the
// signal on the left by pad_left, convolve it with each (max_support-tap, possibly complex) FIR
// filter, take |y|^2 or |y| per sample, and pool 2S samples every S with the integration window;
// floor + log.
//
// This first version evaluates the FIR directly in time (the reference uses overlap-save FFTs):
// every thread owns kSamplesPerThread consecutive output samples of one filter and slides the
// signal through registers, so the inner loop is FFMA on registers with one broadcast
// shared-memory load of the tap and one of the new sample per kSamplesPerThread*2 FMAs.  Pooling
// is fused: |y|^p never leaves registers; partial window sums go to shared-memory accumulators.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace pds {

constexpr int kSiThreads = 256;
constexpr int kSiTileFrames = 8;        // frames per tile
constexpr int kSiSamplesPerThread = 8;  // consecutive outputs per thread

struct SiParams {
  const float* sig;
  const pds_tile* tiles;
  long long n_tiles;
  float* out;
  const float* h_re;    // [C][M]
  const float* h_im;    // [C][M] (unused if is_real)
  const float* window;  // [2S]
  int S, C, M, pad_left;
  int use_log;
  float log_floor;
  // overlap-save path (si_fft_kernel)
  const float2* hc;        // [C][32][32] conj(DFT_1024(h_c)) / 1024, element k = lane + 32 * reg at [reg][lane]
  const float2* tw;        // [32][32]    W_1024^(lane * k1) at [k1][lane]
  int valid_per_fft;       // exact outputs of one 1024-point block: 1024 - (M - 1)
  int ffts_per_tile;       // blocks that cover the (tile_frames + 1) * S pooled samples of a full tile
  int tile_frames;         // frames per tile on this path
  // long supports (si_fft_big_kernel): blocks of 1024 * big_R points, one block per tile
  const float2* hc_big;    // [C][R][1024] conj(DFT_N(h_c)) / N, element k = R * n1 + n2 at [n2][n1]
  const float2* tw_big;    // [R][1024]    W_N^(n2 * k1) at [n2][k1]
  int big_R;
  // real banks: two filters share one complex transform, y1 - i y2 = FFT(conj(X) (Hc1 - i Hc2)); `hc` and
  // `hc_big` then hold ceil(C / 2) combined spectra
  int paired;
};

// y index i of the full convolution reads padded samples i-k; padded index q maps to x[q - pad_left].
// With the taps reversed (g[j] = h[Mp-1-j], Mp = M rounded up to a multiple of 8, zero padded) this
// is the correlation y[i] = sum_j g[j] * xs[i + j] over the staged span xs.
template <bool REAL, bool POWER>
__global__ void __launch_bounds__(kSiThreads, 2) si_direct_kernel(const __grid_constant__ SiParams p) {
  extern __shared__ __align__(16) float smem[];
  const int S = p.S, M = p.M, C = p.C;
  const int Mp = (M + 7) & ~7;
  constexpr int SPT = kSiSamplesPerThread;
  const int ny_max = (kSiTileFrames + 1) * S;  // pooled samples per tile
  const int ny_pad = (ny_max + 32 * SPT - 1) / (32 * SPT) * (32 * SPT);
  const int nx_max = ny_pad + Mp + 8;          // staged samples (idle lanes read in bounds)
  float* s_x = smem;                           // [nx_max]
  float* s_w = s_x + ((nx_max + 3) & ~3);      // [2S]
  float* s_acc = s_w + ((2 * S + 3) & ~3);     // [kSiTileFrames][C]
  float2* s_h = reinterpret_cast<float2*>(s_acc + ((kSiTileFrames * C + 3) & ~3));  // [warps][Mp] reversed taps
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kSiThreads / 32;
  float2* my_h = s_h + warp * Mp;

  for (int i = tid; i < 2 * S; i += kSiThreads) s_w[i] = p.window[i];

  for (long long tile_idx = blockIdx.x; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const pds_tile tile = p.tiles[tile_idx];
    const int nframes = tile.nframes;
    const int ny = (nframes + 1) * S;
    const long long y0 = tile.start;  // first pooled sample (index into the full convolution)
    __syncthreads();
    // xs[j] = padded sample (y0 - (Mp-1) + j); zero outside the signal
    for (int j = tid; j < nx_max; j += kSiThreads) {
      const long long g = y0 - (Mp - 1) + j - p.pad_left;
      s_x[j] = (g >= 0 && g < tile.sig_len) ? p.sig[tile.sig_off + g] : 0.f;
    }
    for (int i = tid; i < kSiTileFrames * C; i += kSiThreads) s_acc[i] = 0.f;
    __syncthreads();

    // warp task = (filter c, chunk of 32*SPT consecutive samples)
    const int chunk = 32 * SPT;
    const int nchunks = (ny + chunk - 1) / chunk;
    for (int c = warp; c < C; c += NW) {
      __syncwarp();
      for (int k = lane; k < Mp; k += 32) {
        const int src = Mp - 1 - k;  // reversed, zero padded at the front of g
        my_h[k] = src < M ? make_float2(p.h_re[c * M + src], REAL ? 0.f : p.h_im[c * M + src])
                          : make_float2(0.f, 0.f);
      }
      __syncwarp();
      for (int ch = 0; ch < nchunks; ++ch) {
        const int i0 = ch * chunk + lane * SPT;  // first output of this thread (multiple of 8)
        // accumulators are packed (re, im) pairs: one FFMA2 per (tap, output) for complex banks
        cplx acc[SPT];
        float w[2 * SPT];
#pragma unroll
        for (int q = 0; q < SPT; ++q) acc[q] = cmake(0.f, 0.f);
        const float4* xs4 = reinterpret_cast<const float4*>(s_x + i0);
        {
          const float4 a = xs4[0], b = xs4[1];
          w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
        }
        for (int j0 = 0; j0 < Mp; j0 += 8) {
          {  // samples j0+8 .. j0+15 of this thread's window
            const float4 a = xs4[j0 / 4 + 2], b = xs4[j0 / 4 + 3];
            w[8] = a.x, w[9] = a.y, w[10] = a.z, w[11] = a.w, w[12] = b.x, w[13] = b.y, w[14] = b.z, w[15] = b.w;
          }
          const ulonglong2* g2 = reinterpret_cast<const ulonglong2*>(my_h + j0);  // two taps per load
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const ulonglong2 gg = g2[jj];
            cplx ga, gb;
            ga.v = gg.x, gb.v = gg.y;
#pragma unroll
            for (int q = 0; q < SPT; ++q) {
              acc[q] = cfma2(ga, cmake(w[q + 2 * jj], w[q + 2 * jj]), acc[q]);
              acc[q] = cfma2(gb, cmake(w[q + 2 * jj + 1], w[q + 2 * jj + 1]), acc[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < SPT; ++q) w[q] = w[q + SPT];
        }
        float re[SPT], im[SPT];
#pragma unroll
        for (int q = 0; q < SPT; ++q) re[q] = cre(acc[q]), im[q] = cim(acc[q]);
        // pooling: sample r = i0 + q feeds frame t = r / S (window half 0) and t - 1 (half 1).
        // When S is a multiple of SPT a thread's samples share one t: sum them in registers, then a
        // segmented warp reduction (lanes with equal t are contiguous) leaves one shared-memory
        // atomic per frame and warp instead of two per sample.
        if (S % SPT == 0) {
          const int t = i0 / S, n0 = i0 - t * S;
          float first_half = 0.f, second_half = 0.f;
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            float u = REAL ? re[q] * re[q] : re[q] * re[q] + im[q] * im[q];
            if (!POWER) u = sqrtf(u);
            if (i0 + q >= ny) u = 0.f;
            first_half = fmaf(s_w[n0 + q], u, first_half);
            second_half = fmaf(s_w[S + n0 + q], u, second_half);
          }
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t_other = __shfl_down_sync(0xffffffffu, t, off);
            const float a = __shfl_down_sync(0xffffffffu, first_half, off);
            const float b = __shfl_down_sync(0xffffffffu, second_half, off);
            if (lane + off < 32 && t_other == t) first_half += a, second_half += b;
          }
          const int t_prev = __shfl_up_sync(0xffffffffu, t, 1);
          if (lane == 0 || t_prev != t) {  // segment leader
            if (t < nframes) atomicAdd(&s_acc[t * C + c], first_half);
            if (t >= 1 && t - 1 < nframes) atomicAdd(&s_acc[(t - 1) * C + c], second_half);
          }
        } else {
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            const int r = i0 + q;
            if (r < ny) {
              float u = REAL ? re[q] * re[q] : re[q] * re[q] + im[q] * im[q];
              if (!POWER) u = sqrtf(u);
              const int t = r / S, n = r - t * S;
              if (t < nframes) atomicAdd(&s_acc[t * C + c], s_w[n] * u);
              if (t >= 1 && t - 1 < nframes) atomicAdd(&s_acc[(t - 1) * C + c], s_w[S + n] * u);
            }
          }
        }
      }
    }
    __syncthreads();
    float* __restrict__ dst = p.out + tile.out_row * C;
    for (int i = tid; i < nframes * C; i += kSiThreads) {
      float v = s_acc[i];
      if (p.use_log) v = __logf(fmaxf(v, p.log_floor));
      dst[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// overlap-save variant (the default whenever three 1024-point blocks cover at least two hops)
//
// y_c = x * h_c is evaluated block-wise in the frequency domain, like the reference does
// (compute.py:893-980), on the in-register FFT of fft_core.cuh: a warp transforms 1024 complex
// points as 32 lanes x 32 registers (element index = lane + 32 * register on the way in AND on
// the way out, so products and inverse transforms need no reordering).
//   geometry: a block of 1024 samples yields V = 1024 - (M - 1) exact outputs; a tile is the
//            largest number of frames whose (frames + 1) * S pooled samples fit in three blocks
//            (C4: V = 637, 10 frames), so no output is wasted on hop alignment: blocks advance
//            by V samples and a pooling hop may straddle two of them.
//   phase 1: warp j computes Xc_j = conj(FFT(block j of the staged samples)) into shared memory.
//   phase 2: warp w owns filters w, w + 14, ...; per block it forms Xc_j * Hc_c (Hc = conj(H)/N,
//            host, double precision) and runs ONE forward FFT: conj(y) = FFT(conj(Y)/N).  |y|^p of
//            all outputs goes to the warp's scratch row; for every hop the block touches, two
//            half-window sums (warp reduction) are added to the frame accumulators of column c.
//            One warp visits a filter's blocks in order, so the sums are bitwise reproducible.
// Cost per frame ~ 0.3 * C transforms of 1024 points instead of 2 * M * C * S multiply-adds.
// ------------------------------------------------------------------------------------------
constexpr int kSiFftN = 1024;
constexpr int kSiFftWarps = 16;
constexpr int kSiFftThreads = 32 * kSiFftWarps;
constexpr int kSiFftBlocks = 3;   // blocks per full tile
constexpr int kSiFftMaxTileFrames = 24;
using SiGeo = FftGeom<2 * kSiFftN>;  // NC = 1024 complex points: G = 32 lanes, R1 = 32 registers
static_assert(SiGeo::G == 32 && SiGeo::R1 == 32 && SiGeo::NSUB == 1, "one warp per transform");

struct SiFftSmem {
  int x, X, tw, scr, w, acc, total;  // float offsets; total in bytes
};
__host__ __device__ inline SiFftSmem si_fft_layout(int S, int C, int valid, int nblocks, int tile_frames) {
  SiFftSmem l;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
  l.x = take((nblocks - 1) * valid + kSiFftN);
  l.X = take(2 * kSiFftN * nblocks);
  l.tw = take(2 * kSiFftN);
  l.scr = take(2 * SiGeo::SCR_FLOAT2 * kSiFftWarps);
  l.w = take(2 * S);
  l.acc = take(2 * nblocks * (tile_frames + 1) * C);  // [block][hop][half][C] partial window sums
  l.total = o * 4;
  return l;
}

__device__ __forceinline__ float approx_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1024-point forward FFT of z (element index = lane + 32 * register, in and out)
__device__ __forceinline__ void si_fft1024(cplx (&z)[32], int lane, const float2* __restrict__ s_tw,
                                           cplx* __restrict__ scr) {
  Dft<32>::run(z);
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) z[k1] = cmul(z[k1], s_tw[k1 * 32 + lane]);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) scr[lane * SiGeo::SCR_STRIDE + k1] = z[k1];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) z[n2] = scr[n2 * SiGeo::SCR_STRIDE + lane];
  __syncwarp();
  Dft<32>::run(z);
}

template <bool POWER, bool PAIRED>
__global__ void __launch_bounds__(kSiFftThreads, 1) si_fft_kernel(const __grid_constant__ SiParams p) {
  extern __shared__ __align__(16) float smem[];
  const int S = p.S, M = p.M, C = p.C, V = p.valid_per_fft, TF = p.tile_frames;
  const SiFftSmem lay = si_fft_layout(S, C, V, p.ffts_per_tile, TF);
  float* s_x = smem + lay.x;
  cplx* s_X = reinterpret_cast<cplx*>(smem + lay.X);
  float2* s_tw = reinterpret_cast<float2*>(smem + lay.tw);
  float* s_w = smem + lay.w;
  float* s_acc = smem + lay.acc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cplx* scr = reinterpret_cast<cplx*>(smem + lay.scr) + warp * SiGeo::SCR_FLOAT2;
  float* s_u = reinterpret_cast<float*>(scr);  // the exchange scratch doubles as the |y|^p row
  const int nx = (p.ffts_per_tile - 1) * V + kSiFftN;

  for (int i = tid; i < 2 * S; i += kSiFftThreads) s_w[i] = p.window[i];
  for (int i = tid; i < kSiFftN; i += kSiFftThreads) s_tw[i] = p.tw[i];

  for (long long tile_idx = blockIdx.x; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const pds_tile tile = p.tiles[tile_idx];
    const int nframes = tile.nframes;
    const int ny = (nframes + 1) * S;      // pooled samples of this tile
    const int nfft = (ny + V - 1) / V;     // blocks that cover them
    const long long y0 = tile.start;       // first pooled sample (index into the full convolution)
    __syncthreads();
    // xs[j] = padded sample (y0 - (M-1) + j); zero outside the signal
    for (int j = tid; j < nx; j += kSiFftThreads) {
      const long long g = y0 - (M - 1) + j - p.pad_left;
      s_x[j] = (g >= 0 && g < tile.sig_len) ? p.sig[tile.sig_off + g] : 0.f;
    }
    for (int i = tid; i < 2 * p.ffts_per_tile * (TF + 1) * C; i += kSiFftThreads) s_acc[i] = 0.f;
    __syncthreads();

    // ---- phase 1: spectra of the sample blocks -----------------------------------------
    if (warp < nfft) {
      const float* xb = s_x + warp * V;
      cplx z[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) z[r] = cmake(xb[lane + 32 * r], 0.f);
      si_fft1024(z, lane, s_tw, scr);
      cplx* X = s_X + warp * kSiFftN;
#pragma unroll
      for (int r = 0; r < 32; ++r) X[r * 32 + lane] = cmake(cre(z[r]), -cim(z[r]));
    }
    __syncthreads();

    // ---- phase 2: one transform per (filter, block), pooling fused -----------------------
    // tasks (c, j) are dealt round-robin over the warps; each task writes the half-window sums
    // of the hops its block touches into slots of its own ([block][hop][half][c]), which the
    // final pass adds in a fixed order: bitwise reproducible without atomics
    const int CT = PAIRED ? (C + 1) / 2 : C;  // transforms per block: filters, or pairs of real filters
    for (int task = warp; task < CT * nfft; task += kSiFftWarps) {
      const int ct = task / nfft, j = task - ct * nfft;
      const int c = PAIRED ? 2 * ct : ct;
      const float2* __restrict__ hc = p.hc + (size_t)ct * kSiFftN + lane;
      const cplx* __restrict__ X = s_X + j * kSiFftN + lane;
      cplx z[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) z[r] = cmul(X[r * 32], __ldg(hc + r * 32));
      si_fft1024(z, lane, s_tw, scr);
      // |y|^p of all 1024 outputs -> s_u (no predicates, coalesced); output n >= M-1 is pooled
      // sample j*V + n - (M-1) of the tile
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (PAIRED) {  // z = y1 - i y2 with y1, y2 real: the second filter's row sits 1024 floats further
          s_u[lane + 32 * r] = POWER ? cre(z[r]) * cre(z[r]) : fabsf(cre(z[r]));
          s_u[kSiFftN + lane + 32 * r] = POWER ? cim(z[r]) * cim(z[r]) : fabsf(cim(z[r]));
        } else {
          float u = cnorm(z[r]);
          if (!POWER) u = approx_sqrt(u);  // |y|: one MUFU, 2 ulp, branch free
          s_u[lane + 32 * r] = u;
        }
      }
      __syncwarp();
      const int lo_blk = j * V, hi_blk = min(lo_blk + V, ny);
      const float* __restrict__ ub = s_u + (M - 1) - lo_blk;  // ub[r] = |y|^p of pooled sample r
      float* __restrict__ part = s_acc + (size_t)j * 2 * (TF + 1) * C + c;
      for (int h = lo_blk / S; h * S < hi_blk; ++h) {
        const int lo = max(h * S, lo_blk), hi = min(h * S + S, hi_blk);
        const float* __restrict__ w1 = s_w - h * S;  // first half-window, indexed by r
        float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
        for (int r = lo + lane; r < hi; r += 32) {
          const float u = ub[r], wa = w1[r], wb = w1[S + r];
          a1 = fmaf(wa, u, a1);
          a2 = fmaf(wb, u, a2);
          if (PAIRED) {
            const float v = ub[kSiFftN + r];
            b1 = fmaf(wa, v, b1);
            b2 = fmaf(wb, v, b2);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          a1 += __shfl_xor_sync(0xffffffffu, a1, off);
          a2 += __shfl_xor_sync(0xffffffffu, a2, off);
          if (PAIRED) {
            b1 += __shfl_xor_sync(0xffffffffu, b1, off);
            b2 += __shfl_xor_sync(0xffffffffu, b2, off);
          }
        }
        if (lane == 0) {
          part[(2 * h) * C] = a1;
          part[(2 * h + 1) * C] = a2;
          if (PAIRED && c + 1 < C) {
            part[(2 * h) * C + 1] = b1;
            part[(2 * h + 1) * C + 1] = b2;
          }
        }
      }
      __syncwarp();
    }
    __syncthreads();
    float* __restrict__ dst = p.out + tile.out_row * C;
    for (int i = tid; i < nframes * C; i += kSiFftThreads) {
      const int t = i / C, c = i - t * C;
      float v = 0.f;
      for (int j = 0; j < nfft; ++j) {  // frame t = first half-window of hop t + second of hop t + 1
        const float* __restrict__ part = s_acc + (size_t)j * 2 * (TF + 1) * C + c;
        v += part[(2 * t) * C] + part[(2 * t + 3) * C];
      }
      if (p.use_log) v = __logf(fmaxf(v, p.log_floor));
      dst[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// overlap-save for long supports (M - 1 + 2 S beyond what three 1024-point blocks cover, e.g. the
// triangular mel bank: 6 987 taps): blocks of N = 1024 R points, R = 2 ... 16, transformed by
// the whole CTA in four steps (n = R n1 + n2, k = k1 + 1024 k2):
//   rows    : warp n2 runs the in-register 1024-point FFT over n1, multiplies by W_N^(n2 k1) and
//             writes row n2 of the 16 x 1024 work buffer Z (which is also its exchange scratch);
//   columns : a thread takes column k1, runs the R-point DFT over n2 and owns outputs k1 + 1024 k2.
// One block per tile: V = N - (M - 1) exact outputs hold V / S - 1 whole frames (both pooling
// halves), so a frame is ONE dot product of the 2 S window with |y|^p -- no partial sums.
// The spectrum of the (real) samples is kept as its lower half, stored by residue k mod R so
// that the rows of the inverse transform read it contiguously; the upper half is the mirrored
// conjugate.  16 / R filters are transformed at a time (16 warps = 16 rows).  The order of all
// sums is fixed, results are bitwise reproducible.
// ------------------------------------------------------------------------------------------
constexpr int kSiBigRowStride = 520;  // complex entries per residue row of the half spectrum (513 used)

struct SiBigSmem {
  int z, xh, tw, w, total;  // float offsets; total in bytes
};
__host__ __device__ inline SiBigSmem si_big_layout(int S, int R) {
  SiBigSmem l;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
  l.z = take(2 * kSiFftWarps * kSiFftN);
  l.xh = take(2 * R * kSiBigRowStride);
  l.tw = take(2 * kSiFftN);
  l.w = take(2 * S);
  l.total = o * 4;
  return l;
}

// 1024-point forward FFT of z (element index = lane + 32 * register, in and out); `row` (1024
// complex values of shared memory owned by this warp) is the exchange scratch, XOR-swizzled so
// that both the row-wise stores and the column-wise loads are conflict free without padding
__device__ __forceinline__ void si_fft1024_row(cplx (&z)[32], int lane, const float2* __restrict__ s_tw,
                                               cplx* __restrict__ row) {
  Dft<32>::run(z);
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) z[k1] = cmul(z[k1], s_tw[k1 * 32 + lane]);
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) row[lane * 32 + (k1 ^ lane)] = z[k1];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) z[n2] = row[n2 * 32 + (lane ^ n2)];
  __syncwarp();
  Dft<32>::run(z);
}

template <bool POWER, int R, bool PAIRED>
__global__ void __launch_bounds__(kSiFftThreads, 1) si_fft_big_kernel(const __grid_constant__ SiParams p) {
  constexpr int N = kSiFftN * R, SLOTS = kSiFftWarps / R;
  extern __shared__ __align__(16) float smem[];
  const int S = p.S, M = p.M, C = p.C;
  const SiBigSmem lay = si_big_layout(S, R);
  cplx* s_z = reinterpret_cast<cplx*>(smem + lay.z);
  float* s_zf = smem + lay.z;
  cplx* s_xh = reinterpret_cast<cplx*>(smem + lay.xh);
  float2* s_tw = reinterpret_cast<float2*>(smem + lay.tw);
  float* s_w = smem + lay.w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = warp / R, n2 = warp % R;
  cplx* row = s_z + warp * kSiFftN;

  for (int i = tid; i < 2 * S; i += kSiFftThreads) s_w[i] = p.window[i];
  for (int i = tid; i < kSiFftN; i += kSiFftThreads) s_tw[i] = p.tw[i];

  for (long long tile_idx = blockIdx.x; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const pds_tile tile = p.tiles[tile_idx];
    const int nframes = tile.nframes;
    __syncthreads();
    // ---- spectrum of the block: rows by warps 0 .. R-1 ------------------------------------
    if (warp < R) {
      // block sample j = R n1 + warp is padded sample tile.start - (M - 1) + j
      const long long g0 = (long long)tile.start - (M - 1) - p.pad_left + warp;
      const float* __restrict__ sig = p.sig + tile.sig_off;
      cplx z[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const long long g = g0 + (long long)R * (lane + 32 * r);
        z[r] = cmake((g >= 0 && g < tile.sig_len) ? __ldg(sig + g) : 0.f, 0.f);
      }
      si_fft1024_row(z, lane, s_tw, row);
      if (warp > 0) {
        const float2* __restrict__ tw = p.tw_big + warp * kSiFftN + lane;
#pragma unroll
        for (int r = 0; r < 32; ++r) z[r] = cmul(z[r], __ldg(tw + 32 * r));
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) row[lane + 32 * r] = z[r];
    }
    __syncthreads();
    for (int col = tid; col < kSiFftN; col += kSiFftThreads) {
      cplx v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) v[q] = s_z[q * kSiFftN + col];
      Dft<R>::run(v);
      // keep conj(X[k]) for k = col + 1024 k2 <= N / 2 at [k mod R][k div R]
      cplx* __restrict__ dst = s_xh + (col % R) * kSiBigRowStride + col / R;
#pragma unroll
      for (int k2 = 0; k2 <= R / 2; ++k2)
        if (k2 < R / 2 || col == 0) dst[k2 * (kSiFftN / R)] = cmake(cre(v[k2]), -cim(v[k2]));
    }
    __syncthreads();

    // ---- 16 / R transforms at a time: filters, or pairs of real filters (y1 - i y2) ------------
    const int CT = PAIRED ? (C + 1) / 2 : C;
    for (int c0 = 0; c0 < CT; c0 += SLOTS) {
      const int c = c0 + slot;
      if (c < CT) {
        const float2* __restrict__ hc = p.hc_big + ((size_t)c * R + n2) * kSiFftN + lane;
        const cplx* __restrict__ lo = s_xh + n2 * kSiBigRowStride + lane;
        // k = R n1 + n2 > N / 2 is the conjugate of entry N - k: residue (R - n2) mod R, index
        // 1023 - n1 (1024 - n1 for residue 0)
        const cplx* __restrict__ hi = s_xh + ((R - n2) % R) * kSiBigRowStride + (1023 + (n2 == 0)) - lane;
        cplx z[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          cplx x;
          if (r < 16) {
            x = lo[32 * r];
          } else {
            const cplx m = hi[-32 * r];
            x = cmake(cre(m), -cim(m));
          }
          z[r] = cmul(x, __ldg(hc + 32 * r));
        }
        si_fft1024_row(z, lane, s_tw, row);
        if (n2 > 0) {
          const float2* __restrict__ tw = p.tw_big + n2 * kSiFftN + lane;
#pragma unroll
          for (int r = 0; r < 32; ++r) z[r] = cmul(z[r], __ldg(tw + 32 * r));
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) row[lane + 32 * r] = z[r];
      }
      __syncthreads();
      // columns: |y|^p of output k1 + 1024 k2 of slot s replaces the real part of Z[s][k2][k1]
      for (int q = tid; q < SLOTS * kSiFftN; q += kSiFftThreads) {
        const int s = q / kSiFftN, col = q - s * kSiFftN;
        if (c0 + s < CT) {
          cplx* __restrict__ zc = s_z + s * N + col;
          cplx v[R];
#pragma unroll
          for (int k = 0; k < R; ++k) v[k] = zc[k * kSiFftN];
          Dft<R>::run(v);
#pragma unroll
          for (int k = 0; k < R; ++k) {
            if (PAIRED) {  // both real filters: |y1|^p replaces the real part, |y2|^p the imaginary part
              zc[k * kSiFftN] = cmake(POWER ? cre(v[k]) * cre(v[k]) : fabsf(cre(v[k])),
                                      POWER ? cim(v[k]) * cim(v[k]) : fabsf(cim(v[k])));
            } else {
              float u = cnorm(v[k]);
              if (!POWER) u = approx_sqrt(u);
              reinterpret_cast<float*>(zc + k * kSiFftN)[0] = u;
            }
          }
        }
      }
      __syncthreads();
      // pooling: frame t of slot s = window . |y|^p[(M - 1) + t S ... + 2 S)
      for (int task = warp; task < SLOTS * nframes; task += kSiFftWarps) {
        const int s = task / nframes, t = task - s * nframes;
        if (c0 + s >= CT) continue;
        const float* __restrict__ u = s_zf + 2 * (s * N + (M - 1) + t * S);
        float acc = 0.f, acc2 = 0.f;
        for (int n = lane; n < 2 * S; n += 32) {
          acc = fmaf(s_w[n], u[2 * n], acc);
          if (PAIRED) acc2 = fmaf(s_w[n], u[2 * n + 1], acc2);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          acc += __shfl_xor_sync(0xffffffffu, acc, off);
          if (PAIRED) acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
        }
        if (lane == 0) {
          if (p.use_log) acc = __logf(fmaxf(acc, p.log_floor)), acc2 = __logf(fmaxf(acc2, p.log_floor));
          const int cf = PAIRED ? 2 * (c0 + s) : c0 + s;
          p.out[(tile.out_row + t) * C + cf] = acc;
          if (PAIRED && cf + 1 < C) p.out[(tile.out_row + t) * C + cf + 1] = acc2;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace pds

using namespace pds;

struct pds_si_plan {
  int device = 0;
  int S = 0, C = 0, M = 0, pad_left = 0, frame_start = 0, frames_lost = 0;
  bool real = false, power = false;
  size_t smem_bytes = 0;
  int grid_limit = 0;
  int tile_frames = kSiTileFrames;
  bool fft = false;  // overlap-save kernel usable
  size_t fft_smem_bytes = 0;
  int fft_grid_limit = 0;
  int big_R = 0;  // > 0: si_fft_big_kernel with blocks of 1024 * big_R points
  bool paired = false;  // real bank on an overlap-save kernel: two filters per complex transform
  size_t big_smem_bytes = 0;
  SiParams params{};
  void* d_blob = nullptr;
};

namespace {
// in-place radix-2 FFT in double precision (filter spectra of the long-support path)
void host_fft(std::vector<std::complex<double>>& a) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  const double two_pi = 6.283185307179586476925286766559;
  for (size_t len = 2; len <= n; len <<= 1) {
    std::vector<std::complex<double>> w(len / 2);
    for (size_t k = 0; k < len / 2; ++k) w[k] = std::polar(1.0, -two_pi * (double)k / (double)len);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w[k];
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

using SiKernel = void (*)(const SiParams);
template <bool PAIRED>
const void* pick_si_big_p(int big_R, bool power) {
  switch (big_R * 2 + (power ? 1 : 0)) {
    case 4: return reinterpret_cast<const void*>(si_fft_big_kernel<false, 2, PAIRED>);
    case 5: return reinterpret_cast<const void*>(si_fft_big_kernel<true, 2, PAIRED>);
    case 8: return reinterpret_cast<const void*>(si_fft_big_kernel<false, 4, PAIRED>);
    case 9: return reinterpret_cast<const void*>(si_fft_big_kernel<true, 4, PAIRED>);
    case 16: return reinterpret_cast<const void*>(si_fft_big_kernel<false, 8, PAIRED>);
    case 17: return reinterpret_cast<const void*>(si_fft_big_kernel<true, 8, PAIRED>);
    case 32: return reinterpret_cast<const void*>(si_fft_big_kernel<false, 16, PAIRED>);
    default: return reinterpret_cast<const void*>(si_fft_big_kernel<true, 16, PAIRED>);
  }
}
const void* pick_si_big(const pds_si_plan* plan) {
  return plan->paired ? pick_si_big_p<true>(plan->big_R, plan->power) : pick_si_big_p<false>(plan->big_R, plan->power);
}
const void* pick_si_fft(const pds_si_plan* plan) {
  if (plan->paired)
    return plan->power ? reinterpret_cast<const void*>(si_fft_kernel<true, true>) : reinterpret_cast<const void*>(si_fft_kernel<false, true>);
  return plan->power ? reinterpret_cast<const void*>(si_fft_kernel<true, false>) : reinterpret_cast<const void*>(si_fft_kernel<false, false>);
}
SiKernel pick_si(const pds_si_plan* plan) {
  if (plan->real) return plan->power ? si_direct_kernel<true, true> : si_direct_kernel<true, false>;
  return plan->power ? si_direct_kernel<false, true> : si_direct_kernel<false, false>;
}
}  // namespace

extern "C" int pds_si_plan_create(const pds_si_desc* d, int device, pds_si_plan** out) {
  if (out) *out = nullptr;
  PDS_REQUIRE(d && out, "null descriptor or output pointer");
  PDS_REQUIRE(d->frame_shift >= 1 && d->num_filts >= 1 && d->max_support >= 1, "bad SI geometry");
  PDS_REQUIRE(d->pad_left >= 0 && d->frame_start >= 0 && d->frames_lost >= 0, "bad SI offsets");
  PDS_REQUIRE(d->h_real && d->window && (d->is_real || d->h_imag), "null table");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) {
    cudaGetLastError();
    set_error("no usable CUDA device %d (found %d); this library has no CPU fallback", device, ndev);
    return PDS_ERR_CUDA;
  }
  DeviceGuard guard(device);
  PDS_CUDA_CHECK(guard.status());
  pds_si_plan* plan = new (std::nothrow) pds_si_plan();
  if (!plan) return PDS_ERR_NOMEM;
  plan->device = device;
  plan->S = d->frame_shift, plan->C = d->num_filts, plan->M = d->max_support;
  plan->pad_left = d->pad_left, plan->frame_start = d->frame_start, plan->frames_lost = d->frames_lost;
  plan->real = d->is_real != 0, plan->power = d->use_power != 0;
  const size_t nh = (size_t)plan->C * plan->M;
  const size_t o_re = 0, o_im = nh * sizeof(float), o_w = 2 * nh * sizeof(float);
  // overlap-save tables: usable when a 1024-point block yields at least one whole hop
  // (PDS_SI_KERNEL=direct forces the time-domain kernel: A/B runs, tests)
  const int valid_per_fft = kSiFftN - (plan->M - 1);
  const int fft_tile_frames =
      valid_per_fft > 0 ? std::min(kSiFftMaxTileFrames, (kSiFftBlocks * valid_per_fft) / plan->S - 1) : 0;
  const char* force = getenv("PDS_SI_KERNEL");
  plan->fft = fft_tile_frames >= 1 && !(force && force[0] == 'd');
  // long supports: the smallest block of 1024 R points (R = 2 ... 16) of which at least 40 % are
  // exact outputs (R = 16 whenever it still holds two hops); preferred over the 1024-point
  // kernel once that one keeps less than 40 % of its outputs
  {
    const bool want_big = !(force && force[0] == 'd') && !(force && force[0] == 's') &&
                          (fft_tile_frames < 1 || valid_per_fft * 5 < kSiFftN * 2 || (force && force[0] == 'b'));
    for (int R = 2; want_big && R <= 16 && plan->big_R == 0; R *= 2) {
      const long long V = (long long)kSiFftN * R - (plan->M - 1);
      if (V >= 2 * plan->S && (V * 5 >= (long long)kSiFftN * R * 2 || R == 16)) plan->big_R = R;
    }
    if (plan->big_R) plan->fft = false;
  }
  // real banks on the overlap-save kernels: filters 2 p and 2 p + 1 share one complex transform
  // (PDS_SI_PAIRS=0 keeps one transform per filter: A/B runs, tests)
  {
    const char* pairs = getenv("PDS_SI_PAIRS");
    plan->paired = plan->real && (plan->fft || plan->big_R) && !(pairs && pairs[0] == '0');
  }
  const int big_N = kSiFftN * plan->big_R;
  const size_t o_hc = (o_w + 2 * (size_t)plan->S * sizeof(float) + 15) & ~(size_t)15;
  const size_t o_tw = o_hc + (plan->fft ? (size_t)plan->C * kSiFftN * sizeof(float2) : 0);
  const size_t o_hcb = o_tw + ((plan->fft || plan->big_R) ? (size_t)kSiFftN * sizeof(float2) : 0);
  const size_t o_twb = o_hcb + (size_t)plan->C * big_N * sizeof(float2);
  const size_t bytes = o_twb + (size_t)big_N * sizeof(float2);
  std::vector<unsigned char> blob(bytes, 0);
  if (plan->big_R) {
    const int R = plan->big_R;
    const double two_pi = 6.283185307179586476925286766559;
    float2* hcb = reinterpret_cast<float2*>(blob.data() + o_hcb);
    std::vector<std::complex<double>> h(big_N), acc_pair(big_N);
    for (int c = 0; c < plan->C; ++c) {
      std::fill(h.begin(), h.end(), std::complex<double>(0.0, 0.0));
      for (int m = 0; m < plan->M; ++m)
        h[m] = std::complex<double>(d->h_real[(size_t)c * plan->M + m],
                                    plan->real ? 0.0 : d->h_imag[(size_t)c * plan->M + m]);
      host_fft(h);
      if (plan->paired) {
        // Hc1 - i Hc2 with Hc = conj(H) / N = (a, b): (a1 + b2, b1 - a2); pair p at index p
        for (int k = 0; k < big_N; ++k) {
          const double a = h[k].real() / big_N, b = -h[k].imag() / big_N;
          acc_pair[k] = (c % 2 == 0) ? std::complex<double>(a, b) : acc_pair[k] + std::complex<double>(b, -a);
        }
        if (c % 2 == 1 || c == plan->C - 1)
          for (int k = 0; k < big_N; ++k)
            hcb[((size_t)(c / 2) * R + k % R) * kSiFftN + k / R] = make_float2((float)acc_pair[k].real(), (float)acc_pair[k].imag());
        continue;
      }
      for (int k = 0; k < big_N; ++k)  // conj(H[k]) / N at [k mod R][k div R]
        hcb[((size_t)c * R + k % R) * kSiFftN + k / R] =
            make_float2((float)(h[k].real() / big_N), (float)(-h[k].imag() / big_N));
    }
    float2* twb = reinterpret_cast<float2*>(blob.data() + o_twb);
    for (int n2 = 0; n2 < R; ++n2)
      for (int k1 = 0; k1 < kSiFftN; ++k1) {
        const double a = two_pi * (double)(((long long)n2 * k1) % big_N) / big_N;
        twb[n2 * kSiFftN + k1] = make_float2((float)std::cos(a), (float)(-std::sin(a)));
      }
    float2* tw = reinterpret_cast<float2*>(blob.data() + o_tw);
    for (int k1 = 0; k1 < 32; ++k1)
      for (int l = 0; l < 32; ++l) {
        const double a = two_pi * ((l * k1) % kSiFftN) / kSiFftN;
        tw[k1 * 32 + l] = make_float2((float)std::cos(a), (float)(-std::sin(a)));
      }
  }
  if (plan->fft) {
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> cs(kSiFftN), sn(kSiFftN);
    for (int i = 0; i < kSiFftN; ++i) cs[i] = std::cos(two_pi * i / kSiFftN), sn[i] = std::sin(two_pi * i / kSiFftN);
    float2* hc = reinterpret_cast<float2*>(blob.data() + o_hc);
    std::vector<std::complex<double>> pair_acc(kSiFftN);
    for (int c = 0; c < plan->C; ++c)
      for (int k = 0; k < kSiFftN; ++k) {
        double re = 0.0, im = 0.0;  // H[k] = sum_m h[m] e^{-2 pi i k m / N}
        for (int m = 0; m < plan->M; ++m) {
          const double hr = d->h_real[(size_t)c * plan->M + m];
          const double hi = plan->real ? 0.0 : d->h_imag[(size_t)c * plan->M + m];
          const int a = (int)(((long long)k * m) % kSiFftN);
          re += hr * cs[a] + hi * sn[a];
          im += hi * cs[a] - hr * sn[a];
        }
        // conj(H) / N at [reg = k / 32][lane = k % 32]; real banks: (a1 + b2, b1 - a2) of filters 2 p, 2 p + 1 at p
        const double a = re / kSiFftN, b = -im / kSiFftN;
        if (!plan->paired) {
          hc[(size_t)c * kSiFftN + k] = make_float2((float)a, (float)b);
        } else if (c % 2 == 0) {
          pair_acc[k] = std::complex<double>(a, b);
          if (c == plan->C - 1) hc[(size_t)(c / 2) * kSiFftN + k] = make_float2((float)a, (float)b);
        } else {
          hc[(size_t)(c / 2) * kSiFftN + k] = make_float2((float)(pair_acc[k].real() + b), (float)(pair_acc[k].imag() - a));
        }
      }
    float2* tw = reinterpret_cast<float2*>(blob.data() + o_tw);
    for (int k1 = 0; k1 < 32; ++k1)
      for (int l = 0; l < 32; ++l) {
        const int a = (l * k1) % kSiFftN;
        tw[k1 * 32 + l] = make_float2((float)cs[a], (float)(-sn[a]));
      }
  }
  memcpy(blob.data() + o_re, d->h_real, nh * sizeof(float));
  if (!plan->real) memcpy(blob.data() + o_im, d->h_imag, nh * sizeof(float));
  memcpy(blob.data() + o_w, d->window, 2 * (size_t)plan->S * sizeof(float));
  cudaError_t err = cudaMalloc(&plan->d_blob, bytes);
  if (err == cudaSuccess) err = cudaMemcpy(plan->d_blob, blob.data(), bytes, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_error("uploading SI tables failed: %s", cudaGetErrorString(err));
    pds_si_plan_destroy(plan);
    return PDS_ERR_CUDA;
  }
  unsigned char* base = static_cast<unsigned char*>(plan->d_blob);
  SiParams& p = plan->params;
  p.h_re = reinterpret_cast<const float*>(base + o_re);
  p.h_im = reinterpret_cast<const float*>(base + o_im);
  p.window = reinterpret_cast<const float*>(base + o_w);
  p.S = plan->S, p.C = plan->C, p.M = plan->M, p.pad_left = plan->pad_left;
  p.use_log = d->use_log ? 1 : 0;
  p.log_floor = d->log_floor;
  p.hc = reinterpret_cast<const float2*>(base + o_hc);
  p.tw = reinterpret_cast<const float2*>(base + o_tw);
  p.valid_per_fft = valid_per_fft;
  p.tile_frames = plan->fft ? fft_tile_frames : kSiTileFrames;
  p.ffts_per_tile = plan->fft ? ((fft_tile_frames + 1) * plan->S + valid_per_fft - 1) / valid_per_fft : 0;
  p.hc_big = reinterpret_cast<const float2*>(base + o_hcb);
  p.tw_big = reinterpret_cast<const float2*>(base + o_twb);
  p.big_R = plan->big_R;
  p.paired = plan->paired ? 1 : 0;
  if (plan->big_R) {
    p.valid_per_fft = big_N - (plan->M - 1);
    p.tile_frames = p.valid_per_fft / plan->S - 1;
    p.ffts_per_tile = 1;
  }
  const int S = plan->S, M = plan->M, C = plan->C;
  const int Mp = (M + 7) & ~7;
  const int ny_max = (kSiTileFrames + 1) * S;
  const int ny_pad = (ny_max + 32 * kSiSamplesPerThread - 1) / (32 * kSiSamplesPerThread) * (32 * kSiSamplesPerThread);
  const int nx_max = ny_pad + Mp + 8;
  plan->smem_bytes = sizeof(float) * (size_t)(((nx_max + 3) & ~3) + ((2 * S + 3) & ~3) +
                                             ((kSiTileFrames * C + 3) & ~3)) +
                     sizeof(float2) * (size_t)Mp * (kSiThreads / 32);
  cudaDeviceProp prop;
  err = cudaGetDeviceProperties(&prop, device);
  if (err == cudaSuccess && plan->big_R) {
    plan->big_smem_bytes = si_big_layout(S, plan->big_R).total;
    if (plan->big_smem_bytes <= prop.sharedMemPerBlockOptin &&
        cudaFuncSetAttribute(pick_si_big(plan), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)plan->big_smem_bytes) == cudaSuccess) {
      plan->fft_grid_limit = prop.multiProcessorCount;
      plan->tile_frames = p.tile_frames;
      *out = plan;
      return PDS_OK;
    }
    cudaGetLastError();
    set_error("SI geometry (S=%d, max_support=%d) needs %zu bytes of shared memory for the long-support kernel", S,
              M, plan->big_smem_bytes);
    pds_si_plan_destroy(plan);
    return PDS_ERR_UNSUPPORTED;
  }
  if (err != cudaSuccess || plan->smem_bytes > prop.sharedMemPerBlockOptin) {
    set_error("SI geometry (S=%d, max_support=%d, %d filters) needs %zu bytes of shared memory", S,
              M, C, plan->smem_bytes);
    pds_si_plan_destroy(plan);
    return err != cudaSuccess ? PDS_ERR_CUDA : PDS_ERR_UNSUPPORTED;
  }
  err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_si(plan)),
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
  if (err != cudaSuccess) {
    set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
    pds_si_plan_destroy(plan);
    return PDS_ERR_CUDA;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void*>(pick_si(plan)),
                                                kSiThreads, plan->smem_bytes);
  plan->grid_limit = prop.multiProcessorCount * std::max(1, occ);
  if (plan->fft) {
    plan->fft_smem_bytes = si_fft_layout(S, C, p.valid_per_fft, p.ffts_per_tile, p.tile_frames).total;
    plan->fft = p.ffts_per_tile <= kSiFftWarps && plan->fft_smem_bytes <= prop.sharedMemPerBlockOptin;
  }
  if (plan->fft) {
    const void* fn = pick_si_fft(plan);
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->fft_smem_bytes);
    if (err != cudaSuccess) {
      cudaGetLastError();
      plan->fft = false;
    } else {
      int occ_fft = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_fft, fn, kSiFftThreads, plan->fft_smem_bytes);
      plan->fft_grid_limit = prop.multiProcessorCount * std::max(1, occ_fft);
    }
  }
  plan->tile_frames = plan->fft ? p.tile_frames : kSiTileFrames;
  p.tile_frames = plan->tile_frames;
  *out = plan;
  return PDS_OK;
}

extern "C" void pds_si_plan_destroy(pds_si_plan* plan) {
  if (!plan) return;
  DeviceGuard guard(plan->device);
  if (plan->d_blob) cudaFree(plan->d_blob);
  delete plan;
}

extern "C" int64_t pds_si_num_frames(const pds_si_plan* plan, int64_t sig_len) {
  if (!plan) return 0;
  const int64_t t = (sig_len + plan->S / 2) / plan->S - plan->frames_lost;
  return t > 0 ? t : 0;
}

extern "C" int pds_si_tile_frames(const pds_si_plan* plan) { return plan ? plan->tile_frames : kSiTileFrames; }

extern "C" int pds_si_layout(const pds_si_plan* plan, int64_t n_utts, const int64_t* sig_len,
                             int64_t* frame_off, int64_t* n_tiles) {
  PDS_REQUIRE(plan && sig_len && frame_off && n_tiles && n_utts >= 0, "bad argument");
  int64_t rows = 0, tiles = 0;
  for (int64_t u = 0; u < n_utts; ++u) {
    PDS_REQUIRE(sig_len[u] >= 0 && sig_len[u] < ((int64_t)1 << 31) - (1 << 20), "utterance %lld too long",
                (long long)u);
    frame_off[u] = rows;
    const int64_t t = pds_si_num_frames(plan, sig_len[u]);
    rows += t;
    tiles += (t + plan->tile_frames - 1) / plan->tile_frames;
  }
  frame_off[n_utts] = rows;
  *n_tiles = tiles;
  return PDS_OK;
}

extern "C" int pds_si_fill_tiles(const pds_si_plan* plan, int64_t n_utts, const int64_t* sig_off,
                                 const int64_t* sig_len, const int64_t* frame_off, pds_tile* tiles) {
  PDS_REQUIRE(plan && sig_off && sig_len && frame_off && (tiles || n_utts == 0), "bad argument");
  int64_t n = 0;
  for (int64_t u = 0; u < n_utts; ++u) {
    const int64_t t_total = frame_off[u + 1] - frame_off[u];
    for (int64_t t0 = 0; t0 < t_total; t0 += plan->tile_frames) {
      pds_tile& tile = tiles[n++];
      tile.sig_off = sig_off[u];
      tile.sig_len = (int32_t)sig_len[u];
      tile.start = (int32_t)(plan->frame_start + t0 * plan->S);
      tile.nframes = (int32_t)std::min<int64_t>(plan->tile_frames, t_total - t0);
      tile.utt = (int32_t)u;
      tile.out_row = frame_off[u] + t0;
    }
  }
  return PDS_OK;
}

extern "C" int pds_si_run(pds_si_plan* plan, const float* d_signal, const pds_tile* d_tiles,
                          int64_t n_tiles, float* d_out, void* stream) {
  PDS_REQUIRE(plan, "null plan");
  if (n_tiles == 0) return PDS_OK;
  PDS_REQUIRE(d_signal && d_tiles && d_out && n_tiles > 0, "null buffer");
  SiParams p = plan->params;
  p.sig = d_signal, p.tiles = d_tiles, p.n_tiles = n_tiles, p.out = d_out;
  if (plan->big_R) {
    const int grid = (int)std::min<int64_t>(n_tiles, plan->fft_grid_limit);
    void* args[] = {&p};
    PDS_CUDA_CHECK(cudaLaunchKernel(pick_si_big(plan), dim3(grid), dim3(kSiFftThreads), args, plan->big_smem_bytes,
                                    static_cast<cudaStream_t>(stream)));
    return PDS_OK;
  }
  if (plan->fft) {
    const int grid = (int)std::min<int64_t>(n_tiles, plan->fft_grid_limit);
    void* args[] = {&p};
    PDS_CUDA_CHECK(cudaLaunchKernel(pick_si_fft(plan), dim3(grid), dim3(kSiFftThreads), args, plan->fft_smem_bytes,
                                    static_cast<cudaStream_t>(stream)));
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  const int grid = (int)std::min<int64_t>(n_tiles, plan->grid_limit);
  pick_si(plan)<<<grid, kSiThreads, plan->smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}
