// stft.cu -- the fused STFT frame-feature kernel (K1+K2 of SURVEY.md) and its C ABI.
//
// One persistent CTA (256 threads) walks a list of tiles; a tile is up to TF = 32 consecutive
// frames of one utterance.  Per tile:
//
//   stage   : the contiguous span of samples the tile touches ((nframes-1)*S + L floats) is
//             copied once from HBM to shared memory.  Interior, 16-byte aligned float32 spans
//             are fetched by ONE thread with a TMA bulk copy (cp.async.bulk + mbarrier) that is
//             issued as soon as the previous tile's fft phase has released the buffer, so the
//             copy overlaps that tile's bank and store phases.  Utterance edges (symmetric
//             reflection), int16 input and fused dither / pre-emphasis (pre.py:90-149) take a
//             cooperative per-element path.
//   fft     : sub-groups of G lanes each take a frame: window multiply, raw-frame energy,
//             R1-point in-register DFT, twiddle, one shared-memory exchange, G-point DFT(s),
//             real-FFT split through warp shuffles, |X|^2 (or |X|) -> s_P[bin][frame].
//   bank    : lane = frame, warp = subset of filters: banded dot product of the power
//             spectrum with the folded weights (compute.py:416-457), floor + log.
//   store   : the (nframes x C) block is contiguous in the packed output; coalesced copy.
//
// Geometry outside the shared-memory FFT's reach (non power-of-two N, N < 256) runs through
// stft_direct_kernel, a plain O(L*K) DFT per frame with the same staging and bank code.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace pds {

constexpr int kThreads = 256;
constexpr int kTileFrames = 32;       // frames per tile on the FFT path (= lanes of the bank phase)
constexpr int kTileStride = 34;       // s_P row stride; 34 = 2 (mod 32) keeps both phases conflict free
constexpr int kDirectTileFrames = 4;  // frames per tile on the direct-DFT path
constexpr int kDirectThreads = 128;

struct StftParams {
  const void* sig;
  const pds_tile* tiles;
  long long n_tiles;
  float* out;
  const float* window;      // [N] zero padded; pre-scaled by 1/2 on the FFT path
  const float2* tw_stage;   // [G][R1]  W_NC^(l*k1)
  const float2* tw_split;   // [G][R1/2] W_N^(l + G*m)
  const float2* tw_direct;  // [N] e^{-2 pi i j / N} (direct path only)
  const int* band_lo;       // [F]
  const int* band_n4;       // [F] taps / 4 after zero padding to a multiple of 8
  const int* band_off;      // [F] offset (floats, multiple of 4) into weights
  const float* weights;     // padded taps, one filter after the other (direct kernel)
  // fused kernels: filters are processed two at a time (ILP); both members of a pair are padded
  // to the same number of 8-tap groups and their weights interleaved group by group
  const int4* pair_desc;    // [npairs] {lo_a * kTileStride, lo_b * kTileStride, groups, weight offset}
  const float* pair_weights;
  int npairs;
  // tensor-core bank (stft_tc_kernel): work items {n0 | m0 << 16, first 16-bin block, blocks,
  // fragment offset}, grouped per warp by tc_wstart[0..8]; weights pre-arranged in mma.m16n8k8
  // B-fragment order, split into tf32 (hi, lo) parts: {b0_hi, b1_hi, b0_lo, b1_lo} per lane and k-step
  const int4* tc_items;
  const int* tc_wstart;
  const float4* tc_frags;
  int tc_nitems;
  int w_probe;              // development probes of stft_w_kernel: 1 = no bank, 2 = no transform
  int w_frag4;              // float4 entries of tc_frags (stft_w_kernel keeps them in shared memory)
  int tc_p_rows;            // rows of the power-spectrum tile the blocks may touch (multiple of 16)
  int weights_total;        // floats in `pair_weights`
  int weights_in_smem;
  int p_rows;               // rows of the power-spectrum tile: K bins + zero rows read by the padding
  int L, S, N, K, F, C;
  int rows_full, row_partial;  // L / (2G) full rows of the stage-1 load, and whether one more is partial
  int span_max;                // floats reserved for the staged samples
  int include_energy, use_log;
  float log_floor, inv_L, preemph, dither;
  int dither_first;
  unsigned long long seed;
};

// ------------------------------------------------------------------------------------------
// sample staging
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float preprocessed_sample(const T* __restrict__ sig, long long base,
                                                     long long g, const StftParams& p, int utt) {
  float x = load_sample(sig, base + g);
  if (p.dither != 0.f && p.dither_first) x += p.dither * philox_normal(p.seed, utt, g);
  if (p.preemph != 0.f && g > 0) {
    float prev = load_sample(sig, base + g - 1);
    if (p.dither != 0.f && p.dither_first) prev += p.dither * philox_normal(p.seed, utt, g - 1);
    x -= p.preemph * prev;
  }
  if (p.dither != 0.f && !p.dither_first) x += p.dither * philox_normal(p.seed, utt, g);
  return x;
}

// ---- mbarrier / TMA bulk copy helpers (sm_90+ PTX) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Which part [a0, a1) of a tile's span can be fetched by a TMA bulk copy?  (uniform across the CTA)
// Interior tiles: all of it.  Utterance-edge tiles: the in-range middle, because the packing puts
// sample `start` of every tile on a 16-byte grid, so shared-memory index and global index are
// congruent mod 4; only the reflected ends are filled by hand.  int16 input or fused
// pre-processing: nothing (a0 == a1).
template <typename T>
__device__ __forceinline__ void bulk_range(const StftParams& p, const pds_tile& tile, int span, int& a0,
                                           int& a1) {
  a0 = a1 = 0;
  if (sizeof(T) != 4 || p.dither != 0.f || p.preemph != 0.f) return;
  const long long first = tile.start;
  const T* src = static_cast<const T*>(p.sig) + tile.sig_off + first;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) != 0) return;
  const int r0 = (int)max(0LL, -first);
  const int r1 = (int)min((long long)span, (long long)tile.sig_len - first);
  const int b0 = (r0 + 3) & ~3, b1 = r1 & ~3;
  if (b1 - b0 >= 64) a0 = b0, a1 = b1;
}

// four samples per thread and trip: aligned vector load, optional pre-emphasis, conversion
template <typename T, int THREADS, bool PRE, bool ALIGNED_DST>
__device__ __forceinline__ void stage_vec4(float* __restrict__ s_x, const T* __restrict__ src, int j0, int nvec,
                                           float c, int tid) {
  for (int v = tid; v < nvec; v += THREADS) {
    const int j = j0 + 4 * v;
    float x0, x1, x2, x3;
    if constexpr (sizeof(T) == 2) {
      const short4 q = *reinterpret_cast<const short4*>(src + j);
      x0 = (float)q.x, x1 = (float)q.y, x2 = (float)q.z, x3 = (float)q.w;
    } else {
      const float4 q = *reinterpret_cast<const float4*>(src + j);
      x0 = q.x, x1 = q.y, x2 = q.z, x3 = q.w;
    }
    if (PRE) {
      const float prev = (float)src[j - 1];
      x3 -= c * x2, x2 -= c * x1, x1 -= c * x0, x0 -= c * prev;
    }
    if (ALIGNED_DST) {
      *reinterpret_cast<float4*>(s_x + j) = make_float4(x0, x1, x2, x3);
    } else {
      s_x[j] = x0, s_x[j + 1] = x1, s_x[j + 2] = x2, s_x[j + 3] = x3;
    }
  }
}

// fused dither (pre.py:90-104), four samples per Philox call: groups are aligned to the
// utterance-relative sample index (the key of the random stream), so loads are per element
template <typename T, int THREADS, bool PRE, bool DITHER_FIRST>
__device__ __forceinline__ void stage_dither4(float* __restrict__ s_x, const T* __restrict__ src, long long first,
                                              int j0, int nvec, float c, float d, uint64_t seed, int utt,
                                              int tid) {
  for (int v = tid; v < nvec; v += THREADS) {
    const int j = j0 + 4 * v;
    const uint64_t group = (uint64_t)(first + j) >> 2;
    float x0 = (float)src[j], x1 = (float)src[j + 1], x2 = (float)src[j + 2], x3 = (float)src[j + 3];
    const float4 n = philox_normal4(seed, utt, group);
    if (PRE && DITHER_FIRST) {  // y[i] = (x[i] + d n[i]) - c (x[i-1] + d n[i-1])
      const float prev = fmaf(d, philox_normal4(seed, utt, group - 1).w, (float)src[j - 1]);
      x0 = fmaf(d, n.x, x0), x1 = fmaf(d, n.y, x1), x2 = fmaf(d, n.z, x2), x3 = fmaf(d, n.w, x3);
      s_x[j] = x0 - c * prev, s_x[j + 1] = x1 - c * x0, s_x[j + 2] = x2 - c * x1, s_x[j + 3] = x3 - c * x2;
    } else {
      if (PRE) {  // pre-emphasis first, then dither
        const float prev = (float)src[j - 1];
        x3 -= c * x2, x2 -= c * x1, x1 -= c * x0, x0 -= c * prev;
      }
      s_x[j] = fmaf(d, n.x, x0), s_x[j + 1] = fmaf(d, n.y, x1), s_x[j + 2] = fmaf(d, n.z, x2), s_x[j + 3] = fmaf(d, n.w, x3);
    }
  }
}

// Cooperative per-element staging of [0, a0) and [a1, span): reflection at the edges, dtype
// conversion, fused pre-processing
template <typename T, int THREADS>
__device__ __forceinline__ void stage_samples_slow(float* __restrict__ s_x, const StftParams& p,
                                                   const pds_tile& tile, int span, int a0, int a1,
                                                   int tid = threadIdx.x) {
  const T* __restrict__ sig = static_cast<const T*>(p.sig);
  const long long first = tile.start;
  // Spans that cannot take the TMA path (16-bit PCM, fused pre-emphasis) but need no random
  // numbers: the in-range middle is converted / filtered four samples at a time with aligned
  // vector loads; only the reflected ends and a few unaligned samples go through the per-element
  // path below.  (pre.py:136-149: y[0] = x[0], y[i] = x[i] - c x[i-1], applied before framing.)
  if (p.dither != 0.f && a1 == a0) {
    const float c = p.preemph;
    int r0 = (int)max(0LL, -first);
    if (c != 0.f) r0 = max(r0, (int)min((long long)span, 4 - first));  // groups 1.. only: sample 0 and its group go per element
    const int r1 = (int)min((long long)span, (long long)tile.sig_len - first);
    if (r1 - r0 >= 64) {
      const T* __restrict__ src = sig + tile.sig_off + first;
      const int j0 = r0 + (int)((4 - ((first + r0) & 3)) & 3);  // first + j0 is a multiple of 4
      const int nvec = (r1 - j0) >> 2;
      const int j1 = j0 + 4 * nvec;
      if (c == 0.f) stage_dither4<T, THREADS, false, false>(s_x, src, first, j0, nvec, c, p.dither, p.seed, tile.utt, tid);
      else if (p.dither_first) stage_dither4<T, THREADS, true, true>(s_x, src, first, j0, nvec, c, p.dither, p.seed, tile.utt, tid);
      else stage_dither4<T, THREADS, true, false>(s_x, src, first, j0, nvec, c, p.dither, p.seed, tile.utt, tid);
      const int rest = j0 + (span - j1);
      for (int e = tid; e < rest; e += THREADS) {
        const int at = e < j0 ? e : e - j0 + j1;
        s_x[at] = preprocessed_sample(sig, tile.sig_off, reflect_index(first + at, tile.sig_len), p, tile.utt);
      }
      return;
    }
  }
  if (p.dither == 0.f && a1 == a0 && (sizeof(T) == 2 || p.preemph != 0.f)) {
    const float c = p.preemph;
    int r0 = (int)max(0LL, -first);
    if (c != 0.f && first + r0 == 0) ++r0;  // sample 0 has no predecessor: per-element path
    const int r1 = (int)min((long long)span, (long long)tile.sig_len - first);
    if (r1 - r0 >= 64) {
      const T* __restrict__ src = sig + tile.sig_off + first;  // src[j] is in range for r0 <= j < r1
      const int mis = (int)((reinterpret_cast<uintptr_t>(src + r0) / sizeof(T)) & 3);
      const int j0 = r0 + ((4 - mis) & 3);
      const int nvec = (r1 - j0) >> 2;
      const int j1 = j0 + 4 * nvec;
      const bool aligned_dst = (j0 & 3) == 0;
      if (c != 0.f) {
        if (aligned_dst) stage_vec4<T, THREADS, true, true>(s_x, src, j0, nvec, c, tid);
        else stage_vec4<T, THREADS, true, false>(s_x, src, j0, nvec, c, tid);
      } else {
        if (aligned_dst) stage_vec4<T, THREADS, false, true>(s_x, src, j0, nvec, c, tid);
        else stage_vec4<T, THREADS, false, false>(s_x, src, j0, nvec, c, tid);
      }
      const int rest = j0 + (span - j1);
      for (int e = tid; e < rest; e += THREADS) {
        const int at = e < j0 ? e : e - j0 + j1;
        s_x[at] = preprocessed_sample(sig, tile.sig_off, reflect_index(first + at, tile.sig_len), p, tile.utt);
      }
      return;
    }
  }
  const int skip = a1 - a0;       // elements covered by the bulk copy
  const int todo = span - skip;   // element e of the hand-filled part sits at e (e < a0) or e + skip
  int e = tid;
  for (; e + 3 * THREADS < todo; e += 4 * THREADS) {  // four independent loads in flight
    float v[4];
    int at[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = e + u * THREADS;
      at[u] = q < a0 ? q : q + skip;
      v[u] = preprocessed_sample(sig, tile.sig_off, reflect_index(first + at[u], tile.sig_len), p, tile.utt);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) s_x[at[u]] = v[u];
  }
  for (; e < todo; e += THREADS) {
    const int at = e < a0 ? e : e + skip;
    s_x[at] = preprocessed_sample(sig, tile.sig_off, reflect_index(first + at, tile.sig_len), p, tile.utt);
  }
}

// natural log of a positive, normal float: one MUFU.LG2 and one FMUL.  The argument has already
// been floored at log_floor, so the denormal rescue of __logf is dead weight.  Absolute error
// < 1e-6 over the range of feature values (|ln x| < 90), against a 1e-3 tolerance.
__device__ __forceinline__ float fast_log(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}

// ------------------------------------------------------------------------------------------
// filter-bank phase: lane = frame, each warp takes every (blockDim/32)-th filter
// ------------------------------------------------------------------------------------------
// `group_warp` of `NWARPS` cooperating warps; lane = frame.  Two filters (2 pi, 2 pi + 1) per trip.
template <int NWARPS, int STRIDE>
__device__ __forceinline__ void bank_pairs(int group_warp, int lane, const float* __restrict__ s_P,
                                           const float* __restrict__ s_e, float* __restrict__ s_out,
                                           const float* __restrict__ weights,
                                           const int4* __restrict__ s_desc, const StftParams& p,
                                           bool power) {
  const bool use_log = p.use_log != 0;
  const float log_floor = p.log_floor;
  float* __restrict__ out_row = s_out + lane * p.C + p.include_energy;
  for (int pi = group_warp; pi < p.npairs; pi += NWARPS) {
    const int4 d = s_desc[pi];  // one broadcast load per pair
    const float4* __restrict__ wt = reinterpret_cast<const float4*>(weights + d.w);
    const float* __restrict__ pa = s_P + d.x + lane;
    const float* __restrict__ pb = s_P + d.y + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
    for (int j = 0; j < d.z; ++j) {  // 16 taps per trip, all 20 loads issued before the FMAs
      const float4 wa0 = wt[0], wa1 = wt[1], wb0 = wt[2], wb1 = wt[3];
      const float x0 = pa[0], x1 = pa[STRIDE], x2 = pa[2 * STRIDE], x3 = pa[3 * STRIDE];
      const float x4 = pa[4 * STRIDE], x5 = pa[5 * STRIDE], x6 = pa[6 * STRIDE], x7 = pa[7 * STRIDE];
      const float y0 = pb[0], y1 = pb[STRIDE], y2 = pb[2 * STRIDE], y3 = pb[3 * STRIDE];
      const float y4 = pb[4 * STRIDE], y5 = pb[5 * STRIDE], y6 = pb[6 * STRIDE], y7 = pb[7 * STRIDE];
      a0 = fmaf(x0, wa0.x, a0);
      a1 = fmaf(x1, wa0.y, a1);
      a2 = fmaf(x2, wa0.z, a2);
      a3 = fmaf(x3, wa0.w, a3);
      b0 = fmaf(y0, wb0.x, b0);
      b1 = fmaf(y1, wb0.y, b1);
      b2 = fmaf(y2, wb0.z, b2);
      b3 = fmaf(y3, wb0.w, b3);
      a0 = fmaf(x4, wa1.x, a0);
      a1 = fmaf(x5, wa1.y, a1);
      a2 = fmaf(x6, wa1.z, a2);
      a3 = fmaf(x7, wa1.w, a3);
      b0 = fmaf(y4, wb1.x, b0);
      b1 = fmaf(y5, wb1.y, b1);
      b2 = fmaf(y6, wb1.z, b2);
      b3 = fmaf(y7, wb1.w, b3);
      wt += 4;
      pa += 8 * STRIDE;
      pb += 8 * STRIDE;
    }
    float va = (a0 + a1) + (a2 + a3), vb = (b0 + b1) + (b2 + b3);
    if (use_log) {
      va = fast_log(fmaxf(va, log_floor));
      vb = fast_log(fmaxf(vb, log_floor));
    }
    out_row[2 * pi] = va;
    if (2 * pi + 1 < p.F) out_row[2 * pi + 1] = vb;
  }
  if (p.include_energy && group_warp == 0) {
    float v = s_e[lane] * p.inv_L;
    if (!power) v = sqrtf(v);
    if (use_log) v = fast_log(fmaxf(v, log_floor));
    s_out[lane * p.C] = v;
  }
}

// How the stage-1 loads are specialised at compile time (no per-row branches in the hot loop):
//   kRows13  : ceil(L / 2G) == 13 R1/16 rows carry data (25 ms frames in a 32 ms DFT and the
//              like); the remaining rows are exact zeros and are never loaded
//   kRows16  : ceil(L / 2G) == R1 (L close or equal to N)
//   kRowsAny : any L <= N: all rows are loaded (the zero-padded window annihilates the tail) and
//              the energy is masked element by element
// In the first two modes only the LAST row can be partially filled; two per-thread predicates
// computed once per kernel mask its samples out of the energy.
enum RowMode { kRows13 = 0, kRows16 = 1, kRowsAny = 2 };

// ------------------------------------------------------------------------------------------
// one frame on one sub-group of G lanes: window, energy, two-stage FFT, split, |X|^p -> pcol
// ------------------------------------------------------------------------------------------
template <int N, bool POWER, int MODE, int NTW, int NTS>
__device__ __forceinline__ void fft_frame(const float* __restrict__ fx, const float* __restrict__ s_w,
                                          float2* __restrict__ scr, float* __restrict__ pcol,
                                          float* __restrict__ e_slot, const float2 (&tw_stage)[NTW],
                                          const float2 (&tw_split)[NTS], int l, bool last_ok0,
                                          bool last_ok1, bool want_energy, const StftParams& p) {
  using Geo = FftGeom<N>;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1, NSUB = Geo::NSUB;
  constexpr bool REGTW = (R1 <= 16);
  constexpr int TS = kTileStride;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;
  const int partner = (G - l) % G;
  const cplx* xp = reinterpret_cast<const cplx*>(fx) + l;   // (x[2n], x[2n+1]) is one 64-bit load
  const cplx* wp = reinterpret_cast<const cplx*>(s_w) + l;
  cplx z[R1];
  cplx energy2 = cmake(0.f, 0.f);
  const cplx last_mask = cmake(last_ok0 ? 1.f : 0.f, last_ok1 ? 1.f : 0.f);
#pragma unroll
  for (int r = 0; r < R1; ++r) {
    if (r < ROWS) {
      cplx x = xp[G * r];
      z[r] = cmul2(x, wp[G * r]);  // window multiply: one FMUL2 per sample pair
      if (MODE == kRowsAny) {
        x = cmul2(x, cmake(2 * (G * r + l) < p.L ? 1.f : 0.f, 2 * (G * r + l) + 1 < p.L ? 1.f : 0.f));
      } else if (r == ROWS - 1) {
        x = cmul2(x, last_mask);
      }
      energy2 = cfma2(x, x, energy2);
    } else {
      z[r] = cmake(0.f, 0.f);
    }
  }
  if (want_energy) {
    float energy = cre(energy2) + cim(energy2);
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, off, G);
    if (l == 0) *e_slot = energy;
  }

  // rows >= ROWS are the frame's zero padding: their first-level butterflies are copies
  constexpr unsigned ZROWS = ROWS >= R1 ? 0u : (zmask_full<R1>() & ~((1u << ROWS) - 1u));
  Dft<R1, ZROWS>::run(z);
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1)
    z[k1] = cmul(z[k1], REGTW ? tw_stage[k1] : __ldg(&p.tw_stage[l * R1 + k1]));
  cplx* cscr = reinterpret_cast<cplx*>(scr);
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) cscr[l * Geo::SCR_STRIDE + k1] = z[k1];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NSUB; ++j) {
    cplx v[G];
#pragma unroll
    for (int n2 = 0; n2 < G; ++n2) v[n2] = cscr[n2 * Geo::SCR_STRIDE + l + G * j];
    Dft<G>::run(v);
#pragma unroll
    for (int k2 = 0; k2 < G; ++k2) z[j + NSUB * k2] = v[k2];
  }
  __syncwarp();

  // real-FFT split: lane l pairs its lower-half registers with the partner's upper half
#pragma unroll
  for (int m = 0; m < R1 / 2; ++m) {
    cplx b;
    b.v = __shfl_sync(0xffffffffu, z[R1 - 1 - m].v, partner, G);
    if (l == 0) b = z[(R1 - m) % R1];
    const float2 w = REGTW ? tw_split[m] : __ldg(&p.tw_split[l * (R1 / 2) + m]);
    cplx xk, xq;
    split_pair(z[m], b, w, xk, xq);
    float pk = cnorm(xk), pq = cnorm(xq);
    if (!POWER) {
      pk = sqrtf(pk);
      pq = sqrtf(pq);
    }
    const int k = l + G * m;
    pcol[k * TS] = pk;
    pcol[(NC - k) * TS] = pq;
  }
  if (l == 0) {  // bin NC/2 pairs with itself; its twiddle is -i
    const cplx a = z[R1 / 2];
    cplx xk, xq;
    split_pair(a, a, make_float2(0.f, -1.f), xk, xq);
    float pk = cnorm(xk);
    if (!POWER) pk = sqrtf(pk);
    pcol[(NC / 2) * TS] = pk;
  }
}

// ------------------------------------------------------------------------------------------
// NF frames at once on one sub-group of G lanes (stft_tc_kernel, R1 <= 16): the same transform as
// fft_frame, with the twiddle tables in shared memory ([k][lane], conflict free) instead of 46
// registers, so that a thread can carry two frames: two independent dependency chains per warp
// hide the latency of the packed butterflies, and window / twiddle loads are shared by the frames.
// The frames take turns in the sub-group's one exchange scratch.
// ------------------------------------------------------------------------------------------
// frames per tile and row stride of the power-spectrum tile (stride = 2 mod 16 keeps the fft-phase
// writes and the A-fragment reads conflict free): 32 frames, 16 for the 1025-bin spectra of N = 2048
template <int N>
struct TcTile {
  static constexpr int kFrames = N <= 1024 ? kTileFrames : 16;
  static constexpr int kStride = kFrames + 2;
};

// first part of the transform (everything that does not touch the power-spectrum tile): window,
// energy, R1-point DFT, twiddle, exchange through the sub-group's scratch, G-point DFT(s).  On
// return lane l holds Z[k] for k = l (mod G) in z[f][(k - l) / G].
template <int N, int MODE, int NF>
__device__ __forceinline__ void fft_front(const float* const (&fx)[NF], const float* __restrict__ s_w,
                                          const float2* __restrict__ s_tws, float2* __restrict__ scr,
                                          cplx (&z)[NF][FftGeom<N>::R1], float (&energy)[NF], int l,
                                          bool last_ok0, bool last_ok1, bool want_energy, const StftParams& p) {
  using Geo = FftGeom<N>;
  constexpr int G = Geo::G, R1 = Geo::R1, NSUB = Geo::NSUB;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;
  const bool odd_shift = (p.S & 1) != 0;  // kernel-uniform
  const cplx* wp = reinterpret_cast<const cplx*>(s_w) + l;
  cplx energy2[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) energy2[f] = cmake(0.f, 0.f);
  const cplx last_mask = cmake(last_ok0 ? 1.f : 0.f, last_ok1 ? 1.f : 0.f);
#pragma unroll
  for (int r = 0; r < R1; ++r) {
    if (r < ROWS) {
      const cplx w = wp[G * r];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        cplx x;
        if (MODE == kRowsAny && odd_shift) {  // odd frame shift (always the generic row mode): frames may
                                              // start on an odd sample, no 8-byte loads
          const float* q = fx[f] + 2 * (l + G * r);
          x = cmake(q[0], q[1]);
        } else {
          x = reinterpret_cast<const cplx*>(fx[f])[l + G * r];
        }
        z[f][r] = cmul2(x, w);  // window multiply: one FMUL2 per sample pair
        if (MODE == kRowsAny) {
          x = cmul2(x, cmake(2 * (G * r + l) < p.L ? 1.f : 0.f, 2 * (G * r + l) + 1 < p.L ? 1.f : 0.f));
        } else if (r == ROWS - 1) {
          x = cmul2(x, last_mask);
        }
        energy2[f] = cfma2(x, x, energy2[f]);
      }
    } else {
#pragma unroll
      for (int f = 0; f < NF; ++f) z[f][r] = cmake(0.f, 0.f);
    }
  }
  if (want_energy) {
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      float e = cre(energy2[f]) + cim(energy2[f]);
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off, G);
      energy[f] = e;
    }
  }
  constexpr unsigned ZROWS = ROWS >= R1 ? 0u : (zmask_full<R1>() & ~((1u << ROWS) - 1u));
#pragma unroll
  for (int f = 0; f < NF; ++f) Dft<R1, ZROWS>::run(z[f]);
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) {
    const float2 t = s_tws[k1 * G + l];
#pragma unroll
    for (int f = 0; f < NF; ++f) z[f][k1] = cmul(z[f][k1], t);
  }
  cplx* cscr = reinterpret_cast<cplx*>(scr);
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) cscr[l * Geo::SCR_STRIDE + k1] = z[f][k1];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      cplx v[G];
#pragma unroll
      for (int n2 = 0; n2 < G; ++n2) v[n2] = cscr[n2 * Geo::SCR_STRIDE + l + G * j];
      Dft<G>::run(v);
#pragma unroll
      for (int k2 = 0; k2 < G; ++k2) z[f][j + NSUB * k2] = v[k2];
    }
    __syncwarp();
  }
}

// second part: real-FFT split (lane l pairs its lower-half registers with the partner's upper
// half) and |X|^p -> pcol[f][bin * TS]
template <int N, bool POWER, int NF>
__device__ __forceinline__ void fft_back(cplx (&z)[NF][FftGeom<N>::R1], const float2* __restrict__ s_twp,
                                         float* const (&pcol)[NF], int l) {
  using Geo = FftGeom<N>;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1;
  constexpr int TS = TcTile<N>::kStride;
  const int partner = (G - l) % G;
#pragma unroll
  for (int m = 0; m < R1 / 2; ++m) {
    const float2 w = s_twp[m * G + l];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      cplx b;
      b.v = __shfl_sync(0xffffffffu, z[f][R1 - 1 - m].v, partner, G);
      if (l == 0) b = z[f][(R1 - m) % R1];
      cplx xk, xq;
      split_pair(z[f][m], b, w, xk, xq);
      float pk = cnorm(xk), pq = cnorm(xq);
      if (!POWER) {
        pk = sqrtf(pk);
        pq = sqrtf(pq);
      }
      const int k = l + G * m;
      pcol[f][k * TS] = pk;
      pcol[f][(NC - k) * TS] = pq;
    }
  }
  if (l == 0) {  // bin NC/2 pairs with itself; its twiddle is -i
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const cplx a = z[f][R1 / 2];
      cplx xk, xq;
      split_pair(a, a, make_float2(0.f, -1.f), xk, xq);
      float pk = cnorm(xk);
      if (!POWER) pk = sqrtf(pk);
      pcol[f][(NC / 2) * TS] = pk;
    }
  }
}

template <int N, bool POWER, int MODE, int NF>
__device__ __forceinline__ void fft_frames(const float* const (&fx)[NF], const float* __restrict__ s_w,
                                           const float2* __restrict__ s_tws, const float2* __restrict__ s_twp,
                                           float2* __restrict__ scr, float* const (&pcol)[NF],
                                           float (&energy)[NF], int l, bool last_ok0, bool last_ok1,
                                           bool want_energy, const StftParams& p) {
  cplx z[NF][FftGeom<N>::R1];
  fft_front<N, MODE, NF>(fx, s_w, s_tws, scr, z, energy, l, last_ok0, last_ok1, want_energy, p);
  fft_back<N, POWER, NF>(z, s_twp, pcol, l);
}

// shared-memory carve-up shared by host (size computation) and device
struct SmemLayout {
  int x, w, scr, P, e, out, bar, desc, wt, total;  // offsets in floats; total in bytes
};

__host__ __device__ inline int take_floats(int& cursor, int n) {
  const int at = cursor;
  cursor += (n + 3) & ~3;  // keep every region 16-byte aligned
  return at;
}

__host__ __device__ inline SmemLayout fused_layout(int N, int G, int R1, int span_max, int p_rows,
                                                   int npairs, int C, int weights_floats) {
  SmemLayout s;
  int o = 0;
  // every frame reads N samples from its start (the window is zero past L): N floats of slack
  s.x = take_floats(o, span_max + N);
  s.w = take_floats(o, N);
  s.scr = take_floats(o, 2 * (kThreads / G) * G * (R1 + 1));
  s.P = take_floats(o, p_rows * kTileStride);
  s.e = take_floats(o, kTileFrames);
  s.out = take_floats(o, kTileFrames * C);
  s.bar = take_floats(o, 4);
  s.desc = take_floats(o, 4 * npairs);
  s.wt = take_floats(o, weights_floats);
  s.total = o * 4;
  return s;
}

// ------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------
template <int N, bool POWER, typename T, int MODE>
__global__ void __launch_bounds__(kThreads, (N <= 512 ? 2 : 1))
    stft_fused_kernel(const __grid_constant__ StftParams p) {
  using Geo = FftGeom<N>;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1, NSUB = Geo::NSUB;
  constexpr int K = NC + 1;
  constexpr int FPR = kThreads / G;  // frames per round
  constexpr bool REGTW = (R1 <= 16);
  constexpr int TS = kTileStride;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;

  extern __shared__ __align__(16) float smem[];
  const SmemLayout lay =
      fused_layout(N, G, R1, p.span_max, p.p_rows, p.npairs, p.C, p.weights_in_smem ? p.weights_total : 0);
  float* s_x = smem + lay.x;
  float* s_w = smem + lay.w;
  float2* s_scr = reinterpret_cast<float2*>(smem + lay.scr);
  float* s_P = smem + lay.P;
  float* s_e = smem + lay.e;
  float* s_out = smem + lay.out;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + lay.bar);
  int4* s_desc = reinterpret_cast<int4*>(smem + lay.desc);
  float* s_wt = smem + lay.wt;

  const int tid = threadIdx.x;
  const int sub = tid / G, l = tid % G;

  // ---- one-time CTA set-up -------------------------------------------------------------
  for (int i = tid; i < N; i += kThreads) s_w[i] = p.window[i];
  for (int i = tid; i < p.npairs; i += kThreads) s_desc[i] = p.pair_desc[i];
  if (p.weights_in_smem)
    for (int i = tid; i < p.weights_total; i += kThreads) s_wt[i] = p.pair_weights[i];
  for (int i = tid; i < (p.p_rows - K) * TS; i += kThreads) s_P[K * TS + i] = 0.f;  // padding rows
  for (int i = tid; i < p.span_max + N; i += kThreads) s_x[i] = 0.f;   // slack must stay finite
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const float* bank_weights = p.weights_in_smem ? s_wt : p.pair_weights;

  float2 tw_stage[REGTW ? R1 : 1], tw_split[REGTW ? R1 / 2 : 1];
  if (REGTW) {
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) tw_stage[k1] = p.tw_stage[l * R1 + k1];
#pragma unroll
    for (int m = 0; m < R1 / 2; ++m) tw_split[m] = p.tw_split[l * (R1 / 2) + m];
  }
  // validity of this lane's two samples in the last loaded row (modes kRows13 / kRows16)
  const bool last_ok0 = 2 * (G * (ROWS - 1) + l) < p.L;
  const bool last_ok1 = 2 * (G * (ROWS - 1) + l) + 1 < p.L;
  float2* scr = s_scr + sub * Geo::SCR_FLOAT2;
  const bool want_energy = p.include_energy != 0;
  __syncthreads();

  // ---- stage the first tile --------------------------------------------------------------
  uint32_t bar_parity = 0;
  bool pending_bulk = false;  // uniform: the current tile's samples arrive through the mbarrier
  long long tile_idx = blockIdx.x;
  pds_tile tile;
  if (tile_idx < p.n_tiles) {
    tile = p.tiles[tile_idx];
    const int span = (tile.nframes - 1) * p.S + p.L;
    int a0, a1;
    bulk_range<T>(p, tile, span, a0, a1);
    pending_bulk = a1 > a0;
    if (pending_bulk && tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // zero-fill above -> async proxy
      mbar_expect_tx(s_bar, (a1 - a0) * 4);
      bulk_copy_g2s(s_x + a0, static_cast<const T*>(p.sig) + tile.sig_off + tile.start + a0, (a1 - a0) * 4, s_bar);
    }
    if (a1 - a0 < span) {
      stage_samples_slow<T, kThreads>(s_x, p, tile, span, a0, a1);
      __syncthreads();
    }
  }

  for (; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const int nframes = tile.nframes;
    // fetch the next descriptor now; it is consumed after the fft phase
    const long long next_idx = tile_idx + gridDim.x;
    pds_tile next_tile = tile;
    if (next_idx < p.n_tiles) next_tile = p.tiles[next_idx];
    if (pending_bulk) {
      mbar_wait(s_bar, bar_parity);
      bar_parity ^= 1;
    }

    // ---- fft phase ---------------------------------------------------------------------
    for (int t0 = 0; t0 < nframes; t0 += FPR) {
      // sub-groups past the end recompute the last frame (identical writes) so that the
      // shuffles below always run with full warps
      const int t = min(t0 + sub, nframes - 1);
      fft_frame<N, POWER, MODE>(s_x + t * p.S, s_w, scr, s_P + t, s_e + t, tw_stage, tw_split, l, last_ok0,
                                last_ok1, want_energy, p);
    }
    __syncthreads();  // s_x is free again, s_P / s_e are complete

    // ---- prefetch the next tile's samples while this tile goes through bank + store ----
    const pds_tile cur = tile;
    bool next_slow = false;
    int next_span = 0, a0 = 0, a1 = 0;
    pending_bulk = false;
    if (next_idx < p.n_tiles) {
      tile = next_tile;
      next_span = (tile.nframes - 1) * p.S + p.L;
      bulk_range<T>(p, tile, next_span, a0, a1);
      pending_bulk = a1 > a0;
      next_slow = a1 - a0 < next_span;
      if (pending_bulk && tid == 0) {
        // order the generic-proxy reads of s_x above before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(s_bar, (a1 - a0) * 4);
        bulk_copy_g2s(s_x + a0, static_cast<const T*>(p.sig) + tile.sig_off + tile.start + a0, (a1 - a0) * 4,
                      s_bar);
      }
    }

    // ---- filter bank + log -------------------------------------------------------------
    bank_pairs<kThreads / 32, TS>(tid >> 5, tid & 31, s_P, s_e, s_out, bank_weights, s_desc, p, POWER);
    if (next_slow) stage_samples_slow<T, kThreads>(s_x, p, tile, next_span, a0, a1);
    __syncthreads();

    // ---- coalesced store ---------------------------------------------------------------
    float* __restrict__ dst = p.out + cur.out_row * p.C;
    const int total = nframes * p.C;
    for (int i = tid; i < total; i += kThreads) dst[i] = s_out[i];
    // no barrier here: the next fft phase only touches s_x / s_P / s_e, and s_out is not written
    // again before the barrier that follows that phase
  }
}

// ------------------------------------------------------------------------------------------
// tensor-core variant of the fused kernel (the default)
//
// Same staging and fft phases as stft_fused_kernel; the filter bank is a block-sparse GEMM
// feat(32 frames x F) = P(32 x K) * W^T(K x F) on the tensor cores (legacy warp-level
// mma.sync.m16n8k8, tf32 inputs, fp32 accumulate -- the tile is far too small for tcgen05).
// Precision: P and W are each split into two tf32 terms (hi = top 11 significand bits, lo = the
// exact remainder) and the three products hi*hi, lo*hi, hi*lo are accumulated: relative error
// ~2^-20 on sums of non-negative terms, i.e. float32-class (tests pin it against the float64
// oracle at the same tolerance as the scalar bank).
//
// Work split: an item is (8 filters) x (16 frames) x (the 16-bin blocks covering the union of
// the eight bands); items are dealt to the eight warps by the host (longest first).  A 16-bin
// block is two k-steps; k-step s takes bins b0 + 4t + 2s (+1) for t = 0..3, which makes the
// A-fragment loads from s_P[bin][frame] (row stride 34) bank-conflict free.  Results go straight
// from the accumulator fragments to global memory (floor, log, masked by nframes / F): no output
// staging, no store phase.  The energy column is written by the fft phase.
// ------------------------------------------------------------------------------------------
// weight fragments are re-read by every tile: keep them in L1; outputs are written once: stream them
__device__ __forceinline__ float4 ldg_keep(const float4* ptr) {
  float4 v;
  asm("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
  return v;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Shared-memory carve-up of stft_tc_kernel.  Everything but the sample buffer has a compile-time
// offset (no address arithmetic, no registers); the samples come last.
template <int N>
struct TcSmem {
  using Geo = FftGeom<N>;
  static constexpr int kProws = ((Geo::NC + 1 + 15) / 16) * 16;  // power-spectrum rows incl. zero padding
  static constexpr int kMaxItems = 96;
  static constexpr int oW = 0;                                                    // window [N]
  static constexpr int oScr = oW + N;                                             // fft exchange scratch
  static constexpr int oP = oScr + 2 * (kThreads / Geo::G) * Geo::SCR_FLOAT2;     // s_P [kProws][34]
  static constexpr int oBar = oP + kProws * TcTile<N>::kStride;                    // mbarrier
  static constexpr int oCtl = oBar + 4;                                           // 2 control blocks x 16 ints
  static constexpr int oRaw = oCtl + 32;                                          // 2 raw tile descriptors
  static constexpr int oWstart = oRaw + 16;                                       // item ranges per warp
  static constexpr int oItems = oWstart + 12;                                     // bank work items (int4)
  static constexpr int oTws = oItems + 4 * kMaxItems;                             // W_NC^(lane*k1) at [k1][lane]
  static constexpr int oTwp = oTws + 2 * Geo::R1 * Geo::G;                        // W_N^(lane+G*m) at [m][lane]
  static constexpr int oX = oTwp + Geo::R1 * Geo::G;                              // samples [span_max + N]
  static_assert(oScr % 4 == 0 && oP % 4 == 0 && oBar % 4 == 0 && oItems % 4 == 0 && oX % 4 == 0, "16-byte regions");
  // every frame reads up to N samples from its start (the window is zero past L), the span covers
  // L of them: N - L (+ a vector of margin) floats of slack
  static __host__ __device__ constexpr int x_floats(int span_max, int L) { return (span_max + (N - L) + 32 + 3) & ~3; }
  static __host__ __device__ constexpr size_t bytes(int span_max, int L) {
    return sizeof(float) * (size_t)(oX + x_floats(span_max, L));
  }
};

// control block of a tile (ints): what every thread needs, prepared once by thread 0
enum TcCtl { kCtlFrames = 0, kCtlFlags = 1, kCtlOutLo = 2, kCtlOutHi = 3, kCtlA0 = 4, kCtlA1 = 5, kCtlSpan = 6,
             kCtlUtt = 7, kCtlStart = 8, kCtlSigLen = 9, kCtlSigOffLo = 10, kCtlSigOffHi = 11 };
constexpr int kFlagHandStaged = 1;  // some samples (reflected edges, int16, fused pre-processing) are staged by hand
constexpr int kFlagBulk = 2;        // part of the span arrives by TMA bulk copy

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// thread 0: raw descriptor -> control block
template <typename T>
__device__ __noinline__ void tc_prepare(const StftParams& p, const pds_tile& tile, int* __restrict__ c) {
  const int span = (tile.nframes - 1) * p.S + p.L;
  int a0, a1;
  bulk_range<T>(p, tile, span, a0, a1);
  const long long out_off = tile.out_row * p.C;
  c[kCtlFrames] = tile.nframes;
  c[kCtlFlags] = (a1 - a0 < span ? kFlagHandStaged : 0) | (a1 > a0 ? kFlagBulk : 0);
  c[kCtlOutLo] = (int)(unsigned)(out_off & 0xffffffffll);
  c[kCtlOutHi] = (int)(out_off >> 32);
  c[kCtlA0] = a0;
  c[kCtlA1] = a1;
  c[kCtlSpan] = span;
  c[kCtlUtt] = tile.utt;
  c[kCtlStart] = tile.start;
  c[kCtlSigLen] = tile.sig_len;
  c[kCtlSigOffLo] = (int)(unsigned)(tile.sig_off & 0xffffffffll);
  c[kCtlSigOffHi] = (int)(tile.sig_off >> 32);
}

__device__ __forceinline__ pds_tile tc_tile_of(const int* __restrict__ c) {
  pds_tile t;
  t.sig_off = ((long long)c[kCtlSigOffHi] << 32) | (unsigned)c[kCtlSigOffLo];
  t.sig_len = c[kCtlSigLen];
  t.start = c[kCtlStart];
  t.nframes = c[kCtlFrames];
  t.utt = c[kCtlUtt];
  t.out_row = 0;
  return t;
}

// thread 0: start the TMA copy of a prepared tile (or complete the phase by hand when it has none)
template <typename T>
__device__ __forceinline__ void tc_issue(const StftParams& p, const int* __restrict__ c, float* s_x, uint64_t* s_bar) {
  if (c[kCtlFlags] & kFlagBulk) {
    const int a0 = c[kCtlA0], a1 = c[kCtlA1];
    const long long sig_off = ((long long)c[kCtlSigOffHi] << 32) | (unsigned)c[kCtlSigOffLo];
    // order the generic-proxy accesses of s_x (made visible by the CTA barrier) before the async write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(s_bar, (a1 - a0) * 4);
    bulk_copy_g2s(s_x + a0, static_cast<const T*>(p.sig) + sig_off + c[kCtlStart] + a0, (a1 - a0) * 4, s_bar);
  } else {
    mbar_arrive(s_bar);
  }
}

template <int TS>
__device__ __forceinline__ void bank_tc(int warp, int lane, const float* __restrict__ s_P,
                                        const int4* __restrict__ s_items, const int* __restrict__ s_wstart,
                                        const float4* __restrict__ frags, const StftParams& p,
                                        float* __restrict__ out_tile, int nframes) {
  const int g = lane >> 2, t = lane & 3;
  const bool use_log = p.use_log != 0;
  const float log_floor = p.log_floor;
  const int C = p.C;
  const int it_end = s_wstart[warp + 1];
  int lane_p = 4 * t * TS + g;                    // this lane's corner of an A fragment
  int lane_o = g * C + p.include_energy + 2 * t;  // ... and of a C fragment in the output tile
  // keep both in registers: re-deriving them from the thread index for every item costs more
  asm volatile("" : "+r"(lane_p), "+r"(lane_o));
  for (int it = s_wstart[warp]; it < it_end; ++it) {
    const int4 d = s_items[it];
    const int n0 = d.x & 0xffff, m0 = d.x >> 16;
    if (m0 >= nframes) continue;
    const float* __restrict__ pa = s_P + d.y * (16 * TS) + m0 + lane_p;
    const float4* __restrict__ fr = frags + d.w + lane;
    float acc[2][2][4];  // [k-step parity][main | correction][fragment]
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[s][q][i] = 0.f;
    int left = d.z;  // >= 1 (host)
#pragma unroll 2
    do {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float4 f = ldg_keep(fr + 32 * s);
        const float a[4] = {pa[(2 * s) * TS], pa[(2 * s) * TS + 8], pa[(2 * s + 1) * TS], pa[(2 * s + 1) * TS + 8]};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          hi[i] = __float_as_uint(a[i]) & 0xffffe000u;
          lo[i] = __float_as_uint(a[i] - __uint_as_float(hi[i]));
        }
        const uint32_t whi0 = __float_as_uint(f.x), whi1 = __float_as_uint(f.y);
        const uint32_t wlo0 = __float_as_uint(f.z), wlo1 = __float_as_uint(f.w);
        mma_tf32(acc[s][0], hi[0], hi[1], hi[2], hi[3], whi0, whi1);
        mma_tf32(acc[s][1], lo[0], lo[1], lo[2], lo[3], whi0, whi1);
        mma_tf32(acc[s][1], hi[0], hi[1], hi[2], hi[3], wlo0, wlo1);
      }
      pa += 16 * TS;
      fr += 64;
    } while (--left > 0);
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = (acc[0][1][i] + acc[1][1][i]) + (acc[0][0][i] + acc[1][0][i]);  // small terms first
      if (use_log) v[i] = fast_log(fmaxf(v[i], log_floor));
    }
    float* __restrict__ r0 = out_tile + (m0 * C + n0 + lane_o);
    float* __restrict__ r1 = r0 + 8 * C;
    const bool c0 = n0 + 2 * t < p.F, c1 = n0 + 2 * t + 1 < p.F;
    if (m0 + g < nframes) {
      if (c0) __stcs(r0, v[0]);
      if (c1) __stcs(r0 + 1, v[1]);
    }
    if (m0 + g + 8 < nframes) {
      if (c0) __stcs(r1, v[2]);
      if (c1) __stcs(r1 + 1, v[3]);
    }
  }
}

template <int N, bool POWER, typename T, int MODE, int NF>
__global__ void __launch_bounds__(kThreads, (N <= 512 ? 2 : 1))
    stft_tc_kernel(const __grid_constant__ StftParams p) {
  using Geo = FftGeom<N>;
  using Lay = TcSmem<N>;
  constexpr int G = Geo::G, R1 = Geo::R1;
  constexpr int FPR = kThreads / G;  // frames per round
  // twiddles live in shared memory ([k][lane]); NF frames per sub-group at once (see fft_frames)
  constexpr int TS = TcTile<N>::kStride;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;

  extern __shared__ __align__(16) float smem[];
  float* const s_x = smem + Lay::oX;
  float* const s_w = smem + Lay::oW;
  float2* const s_scr = reinterpret_cast<float2*>(smem + Lay::oScr);
  float* const s_P = smem + Lay::oP;
  uint64_t* const s_bar = reinterpret_cast<uint64_t*>(smem + Lay::oBar);
  int* const s_ctl = reinterpret_cast<int*>(smem + Lay::oCtl);
  int4* const s_raw = reinterpret_cast<int4*>(smem + Lay::oRaw);
  int* const s_wstart = reinterpret_cast<int*>(smem + Lay::oWstart);
  int4* const s_items = reinterpret_cast<int4*>(smem + Lay::oItems);

  const int tid = threadIdx.x;
  const int sub = tid / G, l = tid % G;
  const int n_tiles = (int)p.n_tiles, stride = gridDim.x;

  // ---- one-time CTA set-up -------------------------------------------------------------
  for (int i = tid; i < N; i += kThreads) s_w[i] = p.window[i];
  for (int i = tid; i < p.tc_nitems; i += kThreads) s_items[i] = p.tc_items[i];
  if (tid <= kThreads / 32) s_wstart[tid] = p.tc_wstart[tid];
  for (int i = tid; i < Lay::kProws * TS; i += kThreads) s_P[i] = 0.f;  // incl. the padding rows
  for (int i = tid; i < Lay::x_floats(p.span_max, p.L); i += kThreads) s_x[i] = 0.f;  // slack must stay finite
  float2* const s_tws = reinterpret_cast<float2*>(smem + Lay::oTws);
  float2* const s_twp = reinterpret_cast<float2*>(smem + Lay::oTwp);
  for (int i = tid; i < R1 * G; i += kThreads) s_tws[i] = p.tw_stage[(i % G) * R1 + i / G];
  for (int i = tid; i < (R1 / 2) * G; i += kThreads) s_twp[i] = p.tw_split[(i % G) * (R1 / 2) + i / G];
  int ti = blockIdx.x;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (ti < n_tiles) {
      const pds_tile first = p.tiles[ti];
      tc_prepare<T>(p, first, s_ctl);
      if (ti + stride < n_tiles) {
        const int4* src = reinterpret_cast<const int4*>(p.tiles + ti + stride);
        cp_async16(s_raw + 2, src);
        cp_async16(s_raw + 3, src + 1);
      }
      cp_async_commit();
    }
  }

  const bool last_ok0 = 2 * (G * (ROWS - 1) + l) < p.L;
  const bool last_ok1 = 2 * (G * (ROWS - 1) + l) + 1 < p.L;
  float2* const scr = s_scr + sub * Geo::SCR_FLOAT2;
  const bool want_energy = p.include_energy != 0;
  __syncthreads();
  if (ti >= n_tiles) return;

  // ---- stage the first tile --------------------------------------------------------------
  if (tid == 0) tc_issue<T>(p, s_ctl, s_x, s_bar);
  if (s_ctl[kCtlFlags] & kFlagHandStaged) {
    stage_samples_slow<T, kThreads>(s_x, p, tc_tile_of(s_ctl), s_ctl[kCtlSpan], s_ctl[kCtlA0], s_ctl[kCtlA1]);
    __syncthreads();
  }

  for (int it = 0; ti < n_tiles; ti += stride, ++it) {
    const int* __restrict__ c = s_ctl + (it & 1) * 16;
    const int* __restrict__ cn = s_ctl + ((it + 1) & 1) * 16;
    const int4 c0 = *reinterpret_cast<const int4*>(c);
    const int nframes = c0.x;
    float* __restrict__ out_tile = p.out + (((long long)c0.w << 32) | (unsigned)c0.z);
    const bool has_next = ti + stride < n_tiles;
    if (tid == 0 && has_next) {
      // the next tile's descriptor was fetched (cp.async) an iteration ago; fetch the one after it
      cp_async_wait_all();
      const int4* raw = s_raw + 2 * ((it + 1) & 1);
      const int4 r0 = raw[0], r1 = raw[1];
      pds_tile nt;
      nt.sig_off = ((long long)r0.y << 32) | (unsigned)r0.x;
      nt.sig_len = r0.z;
      nt.start = r0.w;
      nt.nframes = r1.x;
      nt.utt = r1.y;
      nt.out_row = ((long long)r1.w << 32) | (unsigned)r1.z;
      tc_prepare<T>(p, nt, s_ctl + ((it + 1) & 1) * 16);
      if (ti + 2 * stride < n_tiles) {
        const int4* src = reinterpret_cast<const int4*>(p.tiles + ti + 2 * stride);
        cp_async16(s_raw + 2 * (it & 1), src);
        cp_async16(s_raw + 2 * (it & 1) + 1, src + 1);
      }
      cp_async_commit();
    }
    mbar_wait(s_bar, it & 1);

    // ---- fft phase ---------------------------------------------------------------------
    // sub-groups past the end recompute the last frame (identical writes): full-warp shuffles
    if (NF == 2 && nframes > FPR) {  // two frames per sub-group, FPR apart: the tile is one pass
      const int ta = min(sub, nframes - 1), tb = min(sub + FPR, nframes - 1);
      const float* const fx[2] = {s_x + ta * p.S, s_x + tb * p.S};
      float* const pc[2] = {s_P + ta, s_P + tb};
      float en[2] = {0.f, 0.f};
      fft_frames<N, POWER, MODE, 2>(fx, s_w, s_tws, s_twp, scr, pc, en, l, last_ok0, last_ok1, want_energy, p);
      if (want_energy && l == 0) {  // energy column (compute.py:392-398)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          float v = en[f] * p.inv_L;
          if (!POWER) v = sqrtf(v);
          if (p.use_log) v = fast_log(fmaxf(v, p.log_floor));
          __stcs(out_tile + (f ? tb : ta) * p.C, v);
        }
      }
    } else {
      for (int t0 = 0; t0 < nframes; t0 += FPR) {
        const int ta = min(t0 + sub, nframes - 1);
        const float* const fx[1] = {s_x + ta * p.S};
        float* const pc[1] = {s_P + ta};
        float en[1] = {0.f};
        fft_frames<N, POWER, MODE, 1>(fx, s_w, s_tws, s_twp, scr, pc, en, l, last_ok0, last_ok1, want_energy, p);
        if (want_energy && l == 0) {
          float v = en[0] * p.inv_L;
          if (!POWER) v = sqrtf(v);
          if (p.use_log) v = fast_log(fmaxf(v, p.log_floor));
          __stcs(out_tile + ta * p.C, v);
        }
      }
    }
    __syncthreads();  // s_x is free again, s_P is complete, the next control block is visible

    // ---- start the next tile's TMA copy: it overlaps the bank phase -----------------------
    if (has_next && tid == 0) tc_issue<T>(p, cn, s_x, s_bar);

    // ---- filter bank on the tensor cores, results straight to global memory --------------
    bank_tc<TS>(tid >> 5, tid & 31, s_P, s_items, s_wstart, p.tc_frags, p, out_tile, nframes);
    if (has_next && (cn[kCtlFlags] & kFlagHandStaged))
      stage_samples_slow<T, kThreads>(s_x, p, tc_tile_of(cn), cn[kCtlSpan], cn[kCtlA0], cn[kCtlA1]);
    __syncthreads();  // s_P may be overwritten, hand-staged samples are visible
  }
}

// ------------------------------------------------------------------------------------------
// stft_tc2_kernel: the same tile pipeline with the phases cut differently (two frames per
// sub-group, R1 <= 16).  The transform is split where it first touches the power-spectrum tile:
//
//   phase A : filter bank of the PREVIOUS tile (reads s_P), then window / energy / both DFT
//             stages of THIS tile (reads s_x, private scratch; spectra stay in registers)
//   barrier : s_x is free (the next TMA copy starts), every warp is done with the old s_P
//   phase B : real-FFT split and |X|^p of this tile -> s_P
//   barrier : s_P is complete
//
// In stft_tc_kernel the bank phase stands alone between two barriers: its ten unequal work items
// leave warps idle (15 % of all warp time is spent at the barriers) and, being latency bound, it
// keeps the FMA pipe idle while it lasts.  Here the bank items run side by side with other warps'
// butterflies, and their imbalance is diluted in a phase three times as long.
// ------------------------------------------------------------------------------------------
template <int N, bool POWER, typename T, int MODE, int PROBE = 0>
__global__ void __launch_bounds__(kThreads, 2) stft_tc2_kernel(const __grid_constant__ StftParams p) {
  using Geo = FftGeom<N>;
  using Lay = TcSmem<N>;
  constexpr int G = Geo::G, R1 = Geo::R1;
  static_assert(R1 <= 16 && TcTile<N>::kFrames == 2 * (kThreads / G), "two frames per sub-group cover one tile");
  constexpr int FPR = kThreads / G;
  constexpr int TS = TcTile<N>::kStride;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;

  extern __shared__ __align__(16) float smem[];
  float* const s_x = smem + Lay::oX;
  float* const s_w = smem + Lay::oW;
  float2* const s_scr = reinterpret_cast<float2*>(smem + Lay::oScr);
  float* const s_P = smem + Lay::oP;
  uint64_t* const s_bar = reinterpret_cast<uint64_t*>(smem + Lay::oBar);
  int* const s_ctl = reinterpret_cast<int*>(smem + Lay::oCtl);
  int4* const s_raw = reinterpret_cast<int4*>(smem + Lay::oRaw);
  int* const s_wstart = reinterpret_cast<int*>(smem + Lay::oWstart);
  int4* const s_items = reinterpret_cast<int4*>(smem + Lay::oItems);

  const int tid = threadIdx.x;
  const int sub = tid / G, l = tid % G;
  const int n_tiles = (int)p.n_tiles, stride = gridDim.x;

  // ---- one-time CTA set-up (as in stft_tc_kernel) ------------------------------------------
  for (int i = tid; i < N; i += kThreads) s_w[i] = p.window[i];
  for (int i = tid; i < p.tc_nitems; i += kThreads) s_items[i] = p.tc_items[i];
  if (tid <= kThreads / 32) s_wstart[tid] = p.tc_wstart[tid];
  for (int i = tid; i < Lay::kProws * TS; i += kThreads) s_P[i] = 0.f;
  for (int i = tid; i < Lay::x_floats(p.span_max, p.L); i += kThreads) s_x[i] = 0.f;
  float2* const s_tws = reinterpret_cast<float2*>(smem + Lay::oTws);
  float2* const s_twp = reinterpret_cast<float2*>(smem + Lay::oTwp);
  for (int i = tid; i < R1 * G; i += kThreads) s_tws[i] = p.tw_stage[(i % G) * R1 + i / G];
  for (int i = tid; i < (R1 / 2) * G; i += kThreads) s_twp[i] = p.tw_split[(i % G) * (R1 / 2) + i / G];
  int ti = blockIdx.x;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (ti < n_tiles) {
      const pds_tile first = p.tiles[ti];
      tc_prepare<T>(p, first, s_ctl);
      if (ti + stride < n_tiles) {
        const int4* src = reinterpret_cast<const int4*>(p.tiles + ti + stride);
        cp_async16(s_raw + 2, src);
        cp_async16(s_raw + 3, src + 1);
      }
      cp_async_commit();
    }
  }
  const bool last_ok0 = 2 * (G * (ROWS - 1) + l) < p.L;
  const bool last_ok1 = 2 * (G * (ROWS - 1) + l) + 1 < p.L;
  float2* const scr = s_scr + sub * Geo::SCR_FLOAT2;
  const bool want_energy = p.include_energy != 0;
  __syncthreads();
  if (ti >= n_tiles) return;

  if (tid == 0) tc_issue<T>(p, s_ctl, s_x, s_bar);
  if (s_ctl[kCtlFlags] & kFlagHandStaged) {
    stage_samples_slow<T, kThreads>(s_x, p, tc_tile_of(s_ctl), s_ctl[kCtlSpan], s_ctl[kCtlA0], s_ctl[kCtlA1]);
    __syncthreads();
  }

  float* prev_out = nullptr;  // the tile whose power spectra sit in s_P
  int prev_frames = 0;
  for (int it = 0; ti < n_tiles; ti += stride, ++it) {
    const int* __restrict__ c = s_ctl + (it & 1) * 16;
    const int* __restrict__ cn = s_ctl + ((it + 1) & 1) * 16;
    const int4 c0 = *reinterpret_cast<const int4*>(c);
    const int nframes = c0.x;
    float* __restrict__ out_tile = p.out + (((long long)c0.w << 32) | (unsigned)c0.z);
    const bool has_next = ti + stride < n_tiles;
    if (tid == 0 && has_next) {
      cp_async_wait_all();
      const int4* raw = s_raw + 2 * ((it + 1) & 1);
      const int4 r0 = raw[0], r1 = raw[1];
      pds_tile nt;
      nt.sig_off = ((long long)r0.y << 32) | (unsigned)r0.x;
      nt.sig_len = r0.z;
      nt.start = r0.w;
      nt.nframes = r1.x;
      nt.utt = r1.y;
      nt.out_row = ((long long)r1.w << 32) | (unsigned)r1.z;
      tc_prepare<T>(p, nt, s_ctl + ((it + 1) & 1) * 16);
      if (ti + 2 * stride < n_tiles) {
        const int4* src = reinterpret_cast<const int4*>(p.tiles + ti + 2 * stride);
        cp_async16(s_raw + 2 * (it & 1), src);
        cp_async16(s_raw + 2 * (it & 1) + 1, src + 1);
      }
      cp_async_commit();
    }

    // ---- phase A: bank of the previous tile, then the front of this tile's transform ------
    if (it > 0 && PROBE != 1) bank_tc<TS>(tid >> 5, tid & 31, s_P, s_items, s_wstart, p.tc_frags, p, prev_out, prev_frames);
    mbar_wait(s_bar, it & 1);
    // sub-groups past the end recompute the last frame (identical writes): full-warp shuffles
    const int ta = min(sub, nframes - 1), tb = min(sub + FPR, nframes - 1);
    cplx z[2][R1];
    if (PROBE != 2) {
      const float* const fx[2] = {s_x + ta * p.S, s_x + tb * p.S};
      float en[2] = {0.f, 0.f};
      fft_front<N, MODE, 2>(fx, s_w, s_tws, scr, z, en, l, last_ok0, last_ok1, want_energy, p);
      if (want_energy && l == 0) {  // energy column (compute.py:392-398)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          float v = en[f] * p.inv_L;
          if (!POWER) v = sqrtf(v);
          if (p.use_log) v = fast_log(fmaxf(v, p.log_floor));
          __stcs(out_tile + (f ? tb : ta) * p.C, v);
        }
      }
    }
    __syncthreads();  // s_x is free again, nobody reads the previous s_P any more, next control block visible

    if (has_next && tid == 0) tc_issue<T>(p, cn, s_x, s_bar);

    // ---- phase B: split + |X|^p -> s_P ------------------------------------------------------
    if (PROBE != 2) {
      float* const pc[2] = {s_P + ta, s_P + tb};
      fft_back<N, POWER, 2>(z, s_twp, pc, l);
    }
    if (has_next && (cn[kCtlFlags] & kFlagHandStaged))
      stage_samples_slow<T, kThreads>(s_x, p, tc_tile_of(cn), cn[kCtlSpan], cn[kCtlA0], cn[kCtlA1]);
    prev_out = out_tile;
    prev_frames = nframes;
    __syncthreads();  // s_P is complete, hand-staged samples are visible
  }
  if (PROBE != 1) bank_tc<TS>(tid >> 5, tid & 31, s_P, s_items, s_wstart, p.tc_frags, p, prev_out, prev_frames);
}

// ------------------------------------------------------------------------------------------
// stft_w_kernel: warp-specialised pipeline, one 512-thread CTA per SM, no CTA-wide barriers.
//
//   warps 0..11  transform : three groups of four warps.  A group owns a stream of 16-frame tiles
//                            (one m16 MMA tile); a warp transforms four of the frames (two per
//                            half-warp, as in stft_tc2_kernel) and writes their power spectra to
//                            the group's tile P[g][stage] (two stages).
//   warps 12..14 bank      : warp 12 + g applies the filter bank to the tiles of group g on the
//                            tensor cores and stores the features (block-major: the A fragments
//                            of a 16-bin block are loaded and split into tf32 hi / lo once and
//                            used for every filter group whose band covers the block).
//   warp  15     producer  : lane g fetches the tile descriptors of group g, prepares the control
//                            blocks and issues the TMA bulk copies into the group's two sample
//                            stages.
//
// Hand-over is by mbarriers only (x_full / x_empty per sample stage, p_full / p_empty per spectrum
// stage).  The transform is bound by the FMA pipe and the bank by latency; in the phased kernels
// the two alternate (4.7 ms + 3.1 ms when timed alone, 7.1 ms together), here the bank's
// instructions fill the issue slots the butterflies leave free.  Shared memory: the exchange
// scratch of a half-warp aliases the two P columns it is about to write (they are dead between
// the bank warp's p_empty and this warp's own split), which is what makes room for double-buffered
// spectra, double-buffered samples and the weight fragments (220 KB).
//
// Used for float32 input without fused pre-processing, dft_size 512 geometry (G = R1 = 16), at
// most 64 filters whose weight fragments fit the budget; everything else runs stft_tc2_kernel.
// ------------------------------------------------------------------------------------------
constexpr int kWThreads = 512;
constexpr int kWGroups = 3;          // groups of four transform warps
constexpr int kWTile = 16;           // frames per tile
constexpr int kWStride = 18;         // floats per row of a spectrum tile (conflict free, see TcTile)
constexpr int kWRows = 272;          // rows per spectrum tile: 257 bins padded to 17 blocks of 16
constexpr int kWBlocks = kWRows / 16;
constexpr int kWRing = 8;            // control blocks per group
constexpr int kWMaxNT = 8;           // filter groups of eight (F <= 64)

struct WLayout {  // offsets in floats; every region 16-byte aligned
  static constexpr int oW = 0;                                  // window [512]
  static constexpr int oTws = oW + 512;                         // stage twiddles [16][16] float2
  static constexpr int oTwp = oTws + 512;                       // split twiddles [8][16] float2
  static constexpr int oBar = oTwp + 256;                       // 24 mbarriers
  static constexpr int oCtl = oBar + 64;                        // control blocks [3][8][16] ints
  static constexpr int oTab = oCtl + kWGroups * kWRing * 16;    // block masks [32] + per-filter-group offsets [8]
  static constexpr int oP = oTab + 48;                          // spectra [3][2][272][18]
  static constexpr int oX = oP + kWGroups * 2 * kWRows * kWStride;  // samples [3][2][xstride]
  static_assert(oBar % 4 == 0 && oCtl % 4 == 0 && oTab % 4 == 0 && oP % 4 == 0 && oX % 4 == 0, "16-byte regions");
  static __host__ __device__ constexpr int xstride(int span_max, int L) { return (span_max + (512 - L) + 32 + 3) & ~3; }
  static __host__ __device__ constexpr int o_frag(int span_max, int L) { return oX + kWGroups * 2 * xstride(span_max, L); }
  static __host__ __device__ constexpr size_t bytes(int span_max, int L, int frag_float4) {
    return sizeof(float) * ((size_t)o_frag(span_max, L) + 4 * (size_t)frag_float4);
  }
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// bank warp: features of one 16-frame tile from its power spectra (compute.py:416-460)
template <int NT>
__device__ __forceinline__ void bank_w(int lane, const float* __restrict__ P, const int* __restrict__ s_mask,
                                       const int* __restrict__ s_adj, const float4* __restrict__ s_frag,
                                       const StftParams& p, float* __restrict__ out_tile, int nframes) {
  constexpr int TS = kWStride;
  constexpr int NS = NT <= 5 ? 2 : 1;  // separate accumulators per k-step parity while the registers last
  const int g = lane >> 2, t = lane & 3;
  float acc[NT][NS][2][4];  // [filter group][k-step parity][main | correction][fragment]
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][s2][q][i] = 0.f;
  const float* __restrict__ pa = P + 4 * t * TS + g;
  const float4* __restrict__ frl = s_frag + lane;
  int adj[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) adj[n] = s_adj[n];
  float a[2][4];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    a[s2][0] = pa[(2 * s2) * TS], a[s2][1] = pa[(2 * s2) * TS + 8];
    a[s2][2] = pa[(2 * s2 + 1) * TS], a[s2][3] = pa[(2 * s2 + 1) * TS + 8];
  }
#pragma unroll
  for (int b = 0; b < kWBlocks; ++b) {
    const int mask = s_mask[b];
    uint32_t hi[2][4], lo[2][4];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hi[s2][i] = __float_as_uint(a[s2][i]) & 0xffffe000u;
        lo[s2][i] = __float_as_uint(a[s2][i] - __uint_as_float(hi[s2][i]));
      }
    if (b + 1 < kWBlocks) {  // the next block's spectra are in flight while this block's MMAs run
      const float* __restrict__ pn = pa + (b + 1) * 16 * TS;
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        a[s2][0] = pn[(2 * s2) * TS], a[s2][1] = pn[(2 * s2) * TS + 8];
        a[s2][2] = pn[(2 * s2 + 1) * TS], a[s2][3] = pn[(2 * s2 + 1) * TS + 8];
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (mask & (1 << n)) {
        const float4* __restrict__ fr = frl + (adj[n] + b * 64);
        const float4 f0 = fr[0], f1 = fr[32];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const float4 f = s2 ? f1 : f0;
          const uint32_t whi0 = __float_as_uint(f.x), whi1 = __float_as_uint(f.y);
          const uint32_t wlo0 = __float_as_uint(f.z), wlo1 = __float_as_uint(f.w);
          float(&am)[4] = acc[n][NS == 2 ? s2 : 0][0];
          float(&ac)[4] = acc[n][NS == 2 ? s2 : 0][1];
          mma_tf32(am, hi[s2][0], hi[s2][1], hi[s2][2], hi[s2][3], whi0, whi1);
          mma_tf32(ac, lo[s2][0], lo[s2][1], lo[s2][2], lo[s2][3], whi0, whi1);
          mma_tf32(ac, hi[s2][0], hi[s2][1], hi[s2][2], hi[s2][3], wlo0, wlo1);
        }
      }
    }
  }
  const bool use_log = p.use_log != 0;
  const float log_floor = p.log_floor;
  const int C = p.C;
  float* __restrict__ r0 = out_tile + (g * C + p.include_energy + 2 * t);
  float* __restrict__ r1 = r0 + 8 * C;
  const bool row0 = g < nframes, row1 = g + 8 < nframes;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (NS == 2) v[i] = (acc[n][0][1][i] + acc[n][NS - 1][1][i]) + (acc[n][0][0][i] + acc[n][NS - 1][0][i]);
      else v[i] = acc[n][0][1][i] + acc[n][0][0][i];
      if (use_log) v[i] = fast_log(fmaxf(v[i], log_floor));
    }
    const bool c0 = 8 * n + 2 * t < p.F, c1 = 8 * n + 2 * t + 1 < p.F;
    if (row0) {
      if (c0) __stcs(r0 + 8 * n, v[0]);
      if (c1) __stcs(r0 + 8 * n + 1, v[1]);
    }
    if (row1) {
      if (c0) __stcs(r1 + 8 * n, v[2]);
      if (c1) __stcs(r1 + 8 * n + 1, v[3]);
    }
  }
}

template <bool POWER, int MODE, int NT>
__global__ void __launch_bounds__(kWThreads, 1) stft_w_kernel(const __grid_constant__ StftParams p) {
  constexpr int N = 512;
  using Geo = FftGeom<N>;
  using Lay = WLayout;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1;
  static_assert(G == 16 && R1 == 16, "one frame pair per half-warp");
  constexpr int TS = kWStride;
  constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;

  extern __shared__ __align__(16) float smem[];
  float* const s_w = smem + Lay::oW;
  float2* const s_tws = reinterpret_cast<float2*>(smem + Lay::oTws);
  float2* const s_twp = reinterpret_cast<float2*>(smem + Lay::oTwp);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Lay::oBar);
  uint64_t* const x_full = bars;        // [3][2]
  uint64_t* const x_empty = bars + 6;   // [3][2]
  uint64_t* const p_full = bars + 12;   // [3][2]
  uint64_t* const p_empty = bars + 18;  // [3][2]
  int* const s_ctl = reinterpret_cast<int*>(smem + Lay::oCtl);
  int* const s_mask = reinterpret_cast<int*>(smem + Lay::oTab);
  int* const s_adj = s_mask + 32;
  float* const s_P = smem + Lay::oP;
  float* const s_x = smem + Lay::oX;
  const int xstride = Lay::xstride(p.span_max, p.L);
  float4* const s_frag = reinterpret_cast<float4*>(smem + Lay::o_frag(p.span_max, p.L));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time CTA set-up -------------------------------------------------------------
  for (int i = tid; i < N; i += kWThreads) s_w[i] = p.window[i];
  for (int i = tid; i < R1 * G; i += kWThreads) s_tws[i] = p.tw_stage[(i % G) * R1 + i / G];
  for (int i = tid; i < (R1 / 2) * G; i += kWThreads) s_twp[i] = p.tw_split[(i % G) * (R1 / 2) + i / G];
  for (int i = tid; i < kWGroups * 2 * kWRows * TS; i += kWThreads) s_P[i] = 0.f;
  for (int i = tid; i < kWGroups * 2 * xstride; i += kWThreads) s_x[i] = 0.f;  // slack must stay finite
  for (int i = tid; i < p.w_frag4; i += kWThreads) s_frag[i] = p.tc_frags[i];
  if (tid < 40) s_mask[tid] = 0;
  __syncthreads();
  if (tid < p.tc_nitems) {  // the bank items with m0 == 0 describe each filter group once
    const int4 d = p.tc_items[tid];
    if ((d.x >> 16) == 0) {
      const int n = (d.x & 0xffff) >> 3;
      s_adj[n] = d.w - d.y * 64;
      for (int b = d.y; b < d.y + d.z; ++b) atomicOr(&s_mask[b], 1 << n);
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_tiles = (int)p.n_tiles;
  const int tile_step = kWGroups * gridDim.x;

  if (warp < 4 * kWGroups) {
    // =============================== transform warps ====================================
    const int g = warp >> 2, wg = warp & 3;
    const int h = lane >> 4, l = lane & 15;
    const int col = 4 * wg + 2 * h;  // this half-warp's frames are col and col + 1 of the tile
    const int first = blockIdx.x * kWGroups + g;
    const int n_iter = first < n_tiles ? (n_tiles - first + tile_step - 1) / tile_step : 0;
    const bool last_ok0 = 2 * (G * (ROWS - 1) + l) < p.L;
    const bool last_ok1 = 2 * (G * (ROWS - 1) + l) + 1 < p.L;
    const bool want_energy = p.include_energy != 0;
    const cplx* wp = reinterpret_cast<const cplx*>(s_w) + l;
    const cplx last_mask = cmake(last_ok0 ? 1.f : 0.f, last_ok1 ? 1.f : 0.f);
    const int partner = (G - l) % G;
    for (int j = 0; j < n_iter; ++j) {
      const int st = g * 2 + (j & 1);
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&x_full[st], ph);
      const int* __restrict__ c = s_ctl + (g * kWRing + (j & (kWRing - 1))) * 16;
      const int4 c0 = *reinterpret_cast<const int4*>(c);
      const int nframes = c0.x;
      float* __restrict__ out_tile = p.out + (((long long)c0.w << 32) | (unsigned)c0.z);
      float* const sx = s_x + st * xstride;
      if (c0.y & kFlagHandStaged) {  // utterance edges: the four warps fill in the reflected samples
        stage_samples_slow<float, 128>(sx, p, tc_tile_of(c), c[kCtlSpan], c[kCtlA0], c[kCtlA1], tid & 127);
        named_bar_sync(1 + g, 128);
      }
      // Frames col and col + 1 are transformed even when the tile is shorter (they then read stale,
      // finite samples); their energies are not stored and the bank masks their rows.
      const float* const fx[2] = {sx + col * p.S, sx + (col + 1) * p.S};
      cplx z[2][R1];
      if (p.w_probe == 2) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[st]);
        mbar_wait(&p_empty[st], ph ^ 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
        continue;
      }
      // ---- window, energy (compute.py:392-398), first DFT stage, twiddle -----------------
      {
        cplx energy2[2] = {cmake(0.f, 0.f), cmake(0.f, 0.f)};
        // S == 2 G SH: row r of frame col + 1 is row r + SH of frame col -- one load serves both
        constexpr int SH = 5;
        if (MODE != kRowsAny && p.S == 2 * G * SH) {
          cplx x[ROWS + SH];
#pragma unroll
          for (int r = 0; r < ROWS + SH; ++r) x[r] = reinterpret_cast<const cplx*>(fx[0])[l + G * r];
#pragma unroll
          for (int r = 0; r < R1; ++r) {
            if (r < ROWS) {
              const cplx w = wp[G * r];
#pragma unroll
              for (int f = 0; f < 2; ++f) {
                cplx xv = x[r + f * SH];
                z[f][r] = cmul2(xv, w);
                if (r == ROWS - 1) xv = cmul2(xv, last_mask);
                energy2[f] = cfma2(xv, xv, energy2[f]);
              }
            } else {
              z[0][r] = z[1][r] = cmake(0.f, 0.f);
            }
          }
        } else {
          const bool odd_shift = (p.S & 1) != 0;
#pragma unroll
          for (int r = 0; r < R1; ++r) {
            if (r < ROWS) {
              const cplx w = wp[G * r];
#pragma unroll
              for (int f = 0; f < 2; ++f) {
                cplx x;
                if (MODE == kRowsAny && odd_shift) {
                  const float* q = fx[f] + 2 * (l + G * r);
                  x = cmake(q[0], q[1]);
                } else {
                  x = reinterpret_cast<const cplx*>(fx[f])[l + G * r];
                }
                z[f][r] = cmul2(x, w);
                if (MODE == kRowsAny) {
                  x = cmul2(x, cmake(2 * (G * r + l) < p.L ? 1.f : 0.f, 2 * (G * r + l) + 1 < p.L ? 1.f : 0.f));
                } else if (r == ROWS - 1) {
                  x = cmul2(x, last_mask);
                }
                energy2[f] = cfma2(x, x, energy2[f]);
              }
            } else {
#pragma unroll
              for (int f = 0; f < 2; ++f) z[f][r] = cmake(0.f, 0.f);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[st]);  // the samples are in registers
        if (want_energy) {
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            float e = cre(energy2[f]) + cim(energy2[f]);
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off, G);
            if (l == 0 && col + f < nframes) {
              float v = e * p.inv_L;
              if (!POWER) v = sqrtf(v);
              if (p.use_log) v = fast_log(fmaxf(v, p.log_floor));
              __stcs(out_tile + (col + f) * p.C, v);
            }
          }
        }
        constexpr unsigned ZROWS = ROWS >= R1 ? 0u : (zmask_full<R1>() & ~((1u << ROWS) - 1u));
#pragma unroll
        for (int f = 0; f < 2; ++f) Dft<R1, ZROWS>::run(z[f]);
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) {
          const float2 tw = s_tws[k1 * G + l];
#pragma unroll
          for (int f = 0; f < 2; ++f) z[f][k1] = cmul(z[f][k1], tw);
        }
      }
      // ---- exchange through this half-warp's two columns of P[g][stage], second DFT stage ----
      mbar_wait(&p_empty[st], ph ^ 1);  // the bank warp is done with the tile two iterations back
      float* const Pst = s_P + st * (kWRows * TS);
      {
        cplx* const cs = reinterpret_cast<cplx*>(Pst + col);  // slot j = row j, columns col / col + 1
        constexpr int SLOT = TS / 2;                          // cplx units per row
        // the frames take turns in the one scratch; the second frame's stores and loads are issued
        // before the first frame's butterflies so that their latency hides behind the arithmetic
        cplx v0[G], v1[G];
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) cs[(17 * l + k1) * SLOT] = z[0][k1];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < G; ++n2) v0[n2] = cs[(17 * n2 + l) * SLOT];
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) cs[(17 * l + k1) * SLOT] = z[1][k1];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < G; ++n2) v1[n2] = cs[(17 * n2 + l) * SLOT];
        Dft<G>::run(v0);
        Dft<G>::run(v1);
#pragma unroll
        for (int k2 = 0; k2 < G; ++k2) z[0][k2] = v0[k2], z[1][k2] = v1[k2];
        __syncwarp();
        if (l < kWRows - 257) cs[(257 + l) * SLOT] = cmake(0.f, 0.f);  // the padding rows are zeros again
      }
      // ---- real-FFT split, |X|^p of both frames -> P[bin][col .. col + 1] --------------------
      {
        float2* const pc = reinterpret_cast<float2*>(Pst + col);
        constexpr int SLOT = TS / 2;
#pragma unroll
        for (int m = 0; m < R1 / 2; ++m) {
          const float2 w = s_twp[m * G + l];
          float pk[2], pq[2];
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            cplx b;
            b.v = __shfl_sync(0xffffffffu, z[f][R1 - 1 - m].v, partner, G);
            if (l == 0) b = z[f][(R1 - m) % R1];
            cplx xk, xq;
            split_pair(z[f][m], b, w, xk, xq);
            pk[f] = cnorm(xk), pq[f] = cnorm(xq);
            if (!POWER) {
              pk[f] = sqrtf(pk[f]);
              pq[f] = sqrtf(pq[f]);
            }
          }
          const int k = l + G * m;
          pc[k * SLOT] = make_float2(pk[0], pk[1]);
          pc[(NC - k) * SLOT] = make_float2(pq[0], pq[1]);
        }
        if (l == 0) {  // bin NC/2 pairs with itself; its twiddle is -i
          float pk[2];
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            const cplx a = z[f][R1 / 2];
            cplx xk, xq;
            split_pair(a, a, make_float2(0.f, -1.f), xk, xq);
            pk[f] = cnorm(xk);
            if (!POWER) pk[f] = sqrtf(pk[f]);
          }
          pc[(NC / 2) * SLOT] = make_float2(pk[0], pk[1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
  } else if (warp < 4 * kWGroups + kWGroups) {
    // =============================== bank warps =========================================
    const int g = warp - 4 * kWGroups;
    const int first = blockIdx.x * kWGroups + g;
    const int n_iter = first < n_tiles ? (n_tiles - first + tile_step - 1) / tile_step : 0;
    for (int j = 0; j < n_iter; ++j) {
      const int st = g * 2 + (j & 1);
      mbar_wait(&p_full[st], (j >> 1) & 1);
      const int* __restrict__ c = s_ctl + (g * kWRing + (j & (kWRing - 1))) * 16;
      const int4 c0 = *reinterpret_cast<const int4*>(c);
      float* __restrict__ out_tile = p.out + (((long long)c0.w << 32) | (unsigned)c0.z);
      if (p.w_probe != 1) bank_w<NT>(lane, s_P + st * (kWRows * TS), s_mask, s_adj, s_frag, p, out_tile, c0.x);
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_empty[st]);
    }
  } else if (lane < kWGroups) {
    // =============================== producer lanes =====================================
    const int g = lane;
    const int first = blockIdx.x * kWGroups + g;
    const int n_iter = first < n_tiles ? (n_tiles - first + tile_step - 1) / tile_step : 0;
    for (int j = 0; j < n_iter; ++j) {
      const int st = g * 2 + (j & 1);
      const int4* src = reinterpret_cast<const int4*>(p.tiles + first + (long long)j * tile_step);
      const int4 r0 = __ldg(src), r1 = __ldg(src + 1);
      pds_tile nt;
      nt.sig_off = ((long long)r0.y << 32) | (unsigned)r0.x;
      nt.sig_len = r0.z;
      nt.start = r0.w;
      nt.nframes = r1.x;
      nt.utt = r1.y;
      nt.out_row = ((long long)r1.w << 32) | (unsigned)r1.z;
      int* c = s_ctl + (g * kWRing + (j & (kWRing - 1))) * 16;
      tc_prepare<float>(p, nt, c);
      if (j >= 2) mbar_wait(&x_empty[st], ((j - 2) >> 1) & 1);  // all four warps have consumed the stage
      tc_issue<float>(p, c, s_x + st * xstride, &x_full[st]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// software-pipelined variant (one CTA per SM, no CTA-wide barriers in steady state)
//
//   warps 0..15  compute : every iteration  A) fft of this warp's two frames of tile i
//                                           B) its share of the filter bank of tile i-1
//   warp  16     producer: TMA bulk copies of the sample spans, two tiles ahead (the in-range
//                          middle of utterance-edge tiles included; only the reflected ends are
//                          filled by hand)
//   warps 17..19 store   : coalesced global store of finished tiles
//
// Stages are handed over through mbarriers only (sample ring x_full/x_empty, power-spectrum ring
// p_full/p_empty, output ring o_full/o_empty), so step B always finds its input completed an
// iteration earlier and never stalls; the single tight hand-over is p_empty (the fft of tile i+1
// re-uses the P stage that the bank step of tile i-1 read).  Compared with the phased kernel this
// removes both __syncthreads per tile and the exposed staging / store phases.  setmaxnreg moves
// 4096 registers from the producer/store warpgroup (96 -> 64) to the compute warpgroups (96 -> 104).
// Used for float32 input without fused pre-processing and G = R1 = 16 (N = 512).
// ------------------------------------------------------------------------------------------
constexpr int kWsComputeWarps = 16;
constexpr int kWsStoreWarps = 3;
constexpr int kWsThreads = 32 * (kWsComputeWarps + 1 + kWsStoreWarps);
constexpr int kWsOutStages = 3;

struct WsLayout {
  int x, xstride, w, scr, P, pstride, e, out, bar, desc, wt, total;  // floats; total in bytes
};

__host__ __device__ inline WsLayout ws_layout(int N, int G, int R1, int span_max, int p_rows,
                                              int npairs, int C, int weights_floats) {
  WsLayout s;
  int o = 0;
  s.xstride = (span_max + N + 3) & ~3;
  s.x = take_floats(o, 2 * s.xstride);
  s.w = take_floats(o, N);
  s.scr = take_floats(o, 2 * (2 * kWsComputeWarps) * G * (R1 + 1));
  s.pstride = (p_rows * kTileStride + 3) & ~3;
  s.P = take_floats(o, 2 * s.pstride);
  s.e = take_floats(o, 2 * kTileFrames);
  s.out = take_floats(o, kWsOutStages * kTileFrames * C);
  s.bar = take_floats(o, 32);
  s.desc = take_floats(o, 4 * npairs);
  s.wt = take_floats(o, weights_floats);
  s.total = o * 4;
  return s;
}

template <int N, bool POWER, int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) stft_ws_kernel(const __grid_constant__ StftParams p) {
  using Geo = FftGeom<N>;
  constexpr int NC = Geo::NC, G = Geo::G, R1 = Geo::R1;
  static_assert(G == 16 && R1 == 16, "the pipelined kernel maps one frame to each half-warp");
  constexpr int K = NC + 1;
  constexpr int TS = kTileStride;
  constexpr int NW = kWsComputeWarps;

  extern __shared__ __align__(16) float smem[];
  const WsLayout lay = ws_layout(N, G, R1, p.span_max, p.p_rows, p.npairs, p.C, p.weights_total);
  float* s_x = smem + lay.x;
  float* s_w = smem + lay.w;
  float2* s_scr = reinterpret_cast<float2*>(smem + lay.scr);
  float* s_P = smem + lay.P;
  float* s_e = smem + lay.e;
  float* s_out = smem + lay.out;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar);
  uint64_t* x_full = bars;       // [2]
  uint64_t* x_empty = bars + 2;  // [2]
  uint64_t* p_full = bars + 4;   // [2]
  uint64_t* p_empty = bars + 6;  // [2]
  uint64_t* o_full = bars + 8;   // [3]
  uint64_t* o_empty = bars + 11; // [3]
  int4* s_desc = reinterpret_cast<int4*>(smem + lay.desc);
  float* s_wt = smem + lay.wt;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time CTA set-up (all roles) ---------------------------------------------------
  for (int i = tid; i < N; i += kWsThreads) s_w[i] = p.window[i];
  for (int i = tid; i < p.npairs; i += kWsThreads) s_desc[i] = p.pair_desc[i];
  for (int i = tid; i < p.weights_total; i += kWsThreads) s_wt[i] = p.pair_weights[i];  // always resident
  for (int st = 0; st < 2; ++st)
    for (int i = tid; i < (p.p_rows - K) * TS; i += kWsThreads) s_P[st * lay.pstride + K * TS + i] = 0.f;
  for (int i = tid; i < 2 * lay.xstride; i += kWsThreads) s_x[i] = 0.f;  // slack must stay finite
  if (tid == 0) {
    for (int st = 0; st < 2; ++st) {
      mbar_init(&x_full[st], 1);
      mbar_init(&x_empty[st], NW);
      mbar_init(&p_full[st], NW);
      mbar_init(&p_empty[st], NW);
    }
    for (int st = 0; st < kWsOutStages; ++st) {
      mbar_init(&o_full[st], NW);
      mbar_init(&o_empty[st], kWsStoreWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long first_tile = blockIdx.x, step = gridDim.x;
  const int n_iter = first_tile < p.n_tiles ? (int)((p.n_tiles - first_tile + step - 1) / step) : 0;

  if (warp < NW) {
    // =============================== compute warps ======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int l = lane & 15;
    float2 tw_stage[R1], tw_split[R1 / 2];
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) tw_stage[k1] = p.tw_stage[l * R1 + k1];
#pragma unroll
    for (int m = 0; m < R1 / 2; ++m) tw_split[m] = p.tw_split[l * (R1 / 2) + m];
    constexpr int ROWS = MODE == kRows13 ? (R1 * 13) / 16 : R1;
    const bool last_ok0 = 2 * (G * (ROWS - 1) + l) < p.L;
    const bool last_ok1 = 2 * (G * (ROWS - 1) + l) + 1 < p.L;
    const bool want_energy = p.include_energy != 0;
    const bool use_log = p.use_log != 0;
    const float log_floor = p.log_floor;
    float2* scr = s_scr + (2 * warp + (lane >> 4)) * Geo::SCR_FLOAT2;

    // frame count of tile `it`; the next descriptor is fetched an iteration ahead so that its
    // latency never sits on the critical path
    int nf0 = 0;
    int nf_next = n_iter > 0 ? p.tiles[first_tile].nframes : 0;

    for (int it = 0; it < n_iter + 1; ++it) {
      nf0 = nf_next;
      if (it + 1 < n_iter) nf_next = p.tiles[first_tile + (long long)(it + 1) * step].nframes;
      // ---- A: fft of tile `it` ---------------------------------------------------------
      if (it < n_iter) {
        const int stage = it & 1;
        const uint32_t phase = (it >> 1) & 1;
        mbar_wait(&x_full[stage], phase);
        if (2 * warp < nf0) {
          mbar_wait(&p_empty[stage], phase ^ 1);  // every warp is done with the bank step of tile it-2
          const int t = min(2 * warp + (lane >> 4), nf0 - 1);
          fft_frame<N, POWER, MODE>(s_x + stage * lay.xstride + t * p.S, s_w, scr,
                                    s_P + stage * lay.pstride + t, s_e + stage * kTileFrames + t, tw_stage,
                                    tw_split, l, last_ok0, last_ok1, want_energy, p);
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[stage]);
          mbar_arrive(&x_empty[stage]);
        }
      }
      // ---- B: this warp's filter pairs of tile `it - 1` --------------------------------
      if (it >= 1 && it - 1 < n_iter) {
        const int j = it - 1, stage = j & 1, ostage = j % kWsOutStages;
        mbar_wait(&p_full[stage], (j >> 1) & 1);
        mbar_wait(&o_empty[ostage], ((j / kWsOutStages) & 1) ^ 1);  // the store of tile j-3 has drained
        const float* P = s_P + stage * lay.pstride;
        float* out = s_out + ostage * kTileFrames * p.C;
        float* __restrict__ out_row = out + lane * p.C + p.include_energy;
        // rotate the pair -> warp assignment from tile to tile so that the uneven split
        // (npairs is rarely a multiple of 16) averages out
        const int first = (warp + NW - (j % NW)) % NW;
        for (int pi = first; pi < p.npairs; pi += NW) {
          const int4 d = s_desc[pi];
          const float4* __restrict__ wt = reinterpret_cast<const float4*>(s_wt + d.w);
          const float* __restrict__ pa = P + d.x + lane;
          const float* __restrict__ pb = P + d.y + lane;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll 1
          for (int g = 0; g < d.z; ++g) {
            const float4 wa0 = wt[0], wa1 = wt[1], wb0 = wt[2], wb1 = wt[3];
            const float x0 = pa[0], x1 = pa[TS], x2 = pa[2 * TS], x3 = pa[3 * TS];
            const float x4 = pa[4 * TS], x5 = pa[5 * TS], x6 = pa[6 * TS], x7 = pa[7 * TS];
            const float y0 = pb[0], y1 = pb[TS], y2 = pb[2 * TS], y3 = pb[3 * TS];
            const float y4 = pb[4 * TS], y5 = pb[5 * TS], y6 = pb[6 * TS], y7 = pb[7 * TS];
            a0 = fmaf(x0, wa0.x, a0);
            a1 = fmaf(x1, wa0.y, a1);
            a2 = fmaf(x2, wa0.z, a2);
            a3 = fmaf(x3, wa0.w, a3);
            b0 = fmaf(y0, wb0.x, b0);
            b1 = fmaf(y1, wb0.y, b1);
            b2 = fmaf(y2, wb0.z, b2);
            b3 = fmaf(y3, wb0.w, b3);
            a0 = fmaf(x4, wa1.x, a0);
            a1 = fmaf(x5, wa1.y, a1);
            a2 = fmaf(x6, wa1.z, a2);
            a3 = fmaf(x7, wa1.w, a3);
            b0 = fmaf(y4, wb1.x, b0);
            b1 = fmaf(y5, wb1.y, b1);
            b2 = fmaf(y6, wb1.z, b2);
            b3 = fmaf(y7, wb1.w, b3);
            wt += 4;
            pa += 8 * TS;
            pb += 8 * TS;
          }
          float va = (a0 + a1) + (a2 + a3), vb = (b0 + b1) + (b2 + b3);
          if (use_log) {
            va = fast_log(fmaxf(va, log_floor));
            vb = fast_log(fmaxf(vb, log_floor));
          }
          out_row[2 * pi] = va;
          if (2 * pi + 1 < p.F) out_row[2 * pi + 1] = vb;
        }
        if (want_energy && first == 0) {
          float v = s_e[stage * kTileFrames + lane] * p.inv_L;
          if (!POWER) v = sqrtf(v);
          if (use_log) v = fast_log(fmaxf(v, log_floor));
          out[lane * p.C] = v;
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_empty[stage]);
          mbar_arrive(&o_full[ostage]);
        }
      }
    }
  } else {
    // one setmaxnreg for the whole warpgroup, before its two roles part ways
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  }
  if (warp > NW) {
    // =============================== store warps ========================================
    const int st = tid - 32 * (NW + 1);  // 0..95
    for (int j = 0; j < n_iter; ++j) {
      const int ostage = j % kWsOutStages;
      const pds_tile tile = p.tiles[first_tile + (long long)j * step];
      mbar_wait(&o_full[ostage], (j / kWsOutStages) & 1);
      const float* out = s_out + ostage * kTileFrames * p.C;
      float* __restrict__ dst = p.out + tile.out_row * p.C;
      const int total = tile.nframes * p.C;
      int i = st;
      for (; i + 3 * 32 * kWsStoreWarps < total; i += 4 * 32 * kWsStoreWarps) {  // four stores in flight
        const float v0 = out[i], v1 = out[i + 96], v2 = out[i + 192], v3 = out[i + 288];
        dst[i] = v0, dst[i + 96] = v1, dst[i + 192] = v2, dst[i + 288] = v3;
      }
      for (; i < total; i += 32 * kWsStoreWarps) dst[i] = out[i];
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ostage]);
    }
  } else if (warp == NW) {
    // =============================== producer warp ======================================
    const float* __restrict__ sig = static_cast<const float*>(p.sig);
    for (int it = 0; it < n_iter; ++it) {
      const int stage = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      const pds_tile tile = p.tiles[first_tile + (long long)it * step];
      const int span = (tile.nframes - 1) * p.S + p.L;
      float* dst = s_x + stage * lay.xstride;
      const long long first = tile.start;
      const float* src = sig + tile.sig_off + first;
      const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
      const bool bulk = first >= 0 && first + span <= (long long)tile.sig_len && (span & 3) == 0 && aligned;
      mbar_wait(&x_empty[stage], phase ^ 1);  // the compute warps are done with tile it-2
      if (bulk) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&x_full[stage], span * 4);
          bulk_copy_g2s(dst, src, span * 4, &x_full[stage]);
        }
      } else {
        // utterance edge: the in-range middle of the span still goes through TMA (when the
        // packing put it on a 16-byte grid); only the reflected ends are filled by hand
        const int r0 = (int)max(0LL, -first);
        const int r1 = (int)min((long long)span, (long long)tile.sig_len - first);
        int a0 = (r0 + 3) & ~3, a1 = r1 & ~3;
        if (!aligned || a1 - a0 < 64) a0 = a1 = 0;
        if (a1 > a0 && lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          bulk_copy_g2s(dst + a0, src + a0, (a1 - a0) * 4, &x_full[stage]);
        }
        for (int i = lane; i < a0; i += 32) dst[i] = sig[tile.sig_off + reflect_index(first + i, tile.sig_len)];
        for (int i = a1 + lane; i < span; i += 32)
          dst[i] = sig[tile.sig_off + reflect_index(first + i, tile.sig_len)];
        __syncwarp();
        if (lane == 0) {
          if (a1 > a0) mbar_expect_tx(&x_full[stage], (a1 - a0) * 4);  // arrival + the bytes in flight
          else mbar_arrive(&x_full[stage]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// generic fallback: direct DFT, any N / L / S
// ------------------------------------------------------------------------------------------
template <bool POWER, typename T>
__global__ void __launch_bounds__(kDirectThreads)
    stft_direct_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ __align__(16) float smem[];
  // layout: frame [L] | twiddles [2N] | P [K + 7] | energy scratch [4]
  float* s_f = smem;
  float2* s_tw = reinterpret_cast<float2*>(smem + ((p.L + 3) & ~3));
  float* s_P = reinterpret_cast<float*>(s_tw + p.N);
  float* s_red = s_P + ((p.K + 7 + 3) & ~3);
  const int tid = threadIdx.x;
  const T* __restrict__ sig = static_cast<const T*>(p.sig);
  for (int i = tid; i < p.N; i += kDirectThreads) s_tw[i] = p.tw_direct[i];
  for (int i = tid; i < 7; i += kDirectThreads) s_P[p.K + i] = 0.f;
  __syncthreads();

  for (long long tile_idx = blockIdx.x; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const pds_tile tile = p.tiles[tile_idx];
    for (int t = 0; t < tile.nframes; ++t) {
      // windowed frame + raw energy
      float e = 0.f;
      for (int i = tid; i < p.L; i += kDirectThreads) {
        const long long g = reflect_index((long long)tile.start + (long long)t * p.S + i, tile.sig_len);
        const float x = preprocessed_sample(sig, tile.sig_off, g, p, tile.utt);
        e = fmaf(x, x, e);
        s_f[i] = x * p.window[i];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
      if ((tid & 31) == 0) s_red[tid >> 5] = e;
      __syncthreads();
      for (int k = tid; k < p.K; k += kDirectThreads) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < p.L; ++n) {
          const float2 w = s_tw[idx];
          const float x = s_f[n];
          re = fmaf(x, w.x, re);
          im = fmaf(x, w.y, im);
          idx += k;
          if (idx >= p.N) idx -= p.N;
        }
        const float pw = re * re + im * im;
        s_P[k] = POWER ? pw : sqrtf(pw);
      }
      __syncthreads();
      float* __restrict__ dst = p.out + (tile.out_row + t) * p.C;
      for (int f = tid; f < p.F; f += kDirectThreads) {
        const float* wt = p.weights + p.band_off[f];
        const float* pp = s_P + p.band_lo[f];
        const int n = p.band_n4[f] * 4;
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(pp[j], wt[j], acc);
        if (p.use_log) acc = __logf(fmaxf(acc, p.log_floor));
        dst[p.include_energy + f] = acc;
      }
      if (p.include_energy && tid == 0) {
        float v = 0.f;
        for (int w = 0; w < kDirectThreads / 32; ++w) v += s_red[w];
        v *= p.inv_L;
        if (!POWER) v = sqrtf(v);
        if (p.use_log) v = __logf(fmaxf(v, p.log_floor));
        dst[0] = v;
      }
      __syncthreads();
    }
  }
}

}  // namespace pds

// ==========================================================================================
// host side
// ==========================================================================================
using namespace pds;

struct pds_stft_plan {
  int device = 0;
  int L = 0, S = 0, N = 0, K = 0, F = 0, C = 0, pad_left = 0;
  bool fast = false;
  bool ws = false;  // warp-specialised kernel available (N = 512 geometry, plain float32 input)
  size_t ws_smem_bytes = 0;
  bool fused = false;  // scalar-bank kernel compiled for this size (N = 512 only: A/B runs, > 96 work items)
  bool tc = false;  // tensor-core bank kernel available (the default fast path)
  // PDS_STFT_KERNEL, read once when the plan is made: unset = stft_w_kernel where it applies
  // (else stft_tc2_kernel), '2' = stft_tc2_kernel, '1' = stft_tc_kernel, 's' = scalar-bank
  // kernel, 'w' = the round-1 software-pipelined kernel
  int variant = 5;
  bool want_ws = false, want_scalar = false;
  bool w = false;  // stft_w_kernel usable (float32 input; 16-frame tiles)
  size_t w_smem_bytes = 0;
  int w_nt = 0;
  size_t tc_smem_bytes = 0;
  int tc_grid_limit = 0;
  bool power = false;
  int row_mode = 2;
  int tile_frames = 0;
  int G = 0, R1 = 0;
  size_t smem_bytes = 0;
  int grid_limit = 0;
  int num_sms = 0;
  StftParams params{};
  void* d_blob = nullptr;  // all constant tables, one allocation
  // scratch for pds_stft_compute_host
  void* d_sig = nullptr;
  size_t d_sig_bytes = 0;
  pds_tile* d_tiles = nullptr;
  size_t d_tiles_cap = 0;
  float* d_out = nullptr;
  size_t d_out_bytes = 0;
};

namespace {

using KernelFn = void (*)(const StftParams);

template <int N, int MODE>
KernelFn pick_fused_mode(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_fused_kernel<N, true, short, MODE> : stft_fused_kernel<N, true, float, MODE>;
  return dtype == PDS_I16 ? stft_fused_kernel<N, false, short, MODE> : stft_fused_kernel<N, false, float, MODE>;
}

template <int N>
KernelFn pick_fused(bool power, int dtype, int mode) {
  switch (mode) {
    case kRows13: return pick_fused_mode<N, kRows13>(power, dtype);
    case kRows16: return pick_fused_mode<N, kRows16>(power, dtype);
    default: return pick_fused_mode<N, kRowsAny>(power, dtype);
  }
}

// frames per sub-group and pass: 2 where the registers allow it (R1 <= 16), see fft_frames
template <int N>
constexpr int tc_frames() {
  return FftGeom<N>::R1 <= 16 ? 2 : 1;
}

template <int N, int MODE>
KernelFn pick_tc2_mode(bool power, int dtype) {
  if constexpr (FftGeom<N>::R1 <= 16 && TcTile<N>::kFrames == 2 * (kThreads / FftGeom<N>::G)) {
    if (power) return dtype == PDS_I16 ? stft_tc2_kernel<N, true, short, MODE> : stft_tc2_kernel<N, true, float, MODE>;
    return dtype == PDS_I16 ? stft_tc2_kernel<N, false, short, MODE> : stft_tc2_kernel<N, false, float, MODE>;
  } else {
    return nullptr;
  }
}

template <int N, int MODE>
KernelFn pick_tc_mode(bool power, int dtype, int variant) {
  if (variant == 2) {
    KernelFn fn = pick_tc2_mode<N, MODE>(power, dtype);
    if (fn) return fn;
  }
#ifdef PDS_DEV_N512_ONLY
  if constexpr (N == 512 && MODE == kRows13) {
    if (variant == 3) return stft_tc2_kernel<512, true, float, kRows13, 1>;  // probe: no bank phase
    if (variant == 4) return stft_tc2_kernel<512, true, float, kRows13, 2>;  // probe: no fft phase
  }
#endif
  constexpr int NF = tc_frames<N>();
  if (power) return dtype == PDS_I16 ? stft_tc_kernel<N, true, short, MODE, NF> : stft_tc_kernel<N, true, float, MODE, NF>;
  return dtype == PDS_I16 ? stft_tc_kernel<N, false, short, MODE, NF> : stft_tc_kernel<N, false, float, MODE, NF>;
}

template <int N>
KernelFn pick_tc_n(bool power, int dtype, int mode, int variant) {
  switch (mode) {
    case kRows13: return pick_tc_mode<N, kRows13>(power, dtype, variant);
    case kRows16: return pick_tc_mode<N, kRows16>(power, dtype, variant);
    default: return pick_tc_mode<N, kRowsAny>(power, dtype, variant);
  }
}

KernelFn pick_tc(const pds_stft_plan* plan, int dtype) {
  switch (plan->N) {
#ifndef PDS_DEV_N512_ONLY
    case 256: return pick_tc_n<256>(plan->power, dtype, plan->row_mode, plan->variant);
#endif
    case 512: return pick_tc_n<512>(plan->power, dtype, plan->row_mode, plan->variant);
#ifndef PDS_DEV_N512_ONLY
    case 1024: return pick_tc_n<1024>(plan->power, dtype, plan->row_mode, plan->variant);
    case 2048: return pick_tc_n<2048>(plan->power, dtype, plan->row_mode, plan->variant);
#endif
    default: return nullptr;
  }
}

template <int MODE>
KernelFn pick_w_mode(bool power, int nt) {
  if (power) return nt <= 5 ? stft_w_kernel<true, MODE, 5> : stft_w_kernel<true, MODE, 8>;
  return nt <= 5 ? stft_w_kernel<false, MODE, 5> : stft_w_kernel<false, MODE, 8>;
}

KernelFn pick_w(const pds_stft_plan* plan) {
  switch (plan->row_mode) {
    case kRows13: return pick_w_mode<kRows13>(plan->power, plan->w_nt);
    case kRows16: return pick_w_mode<kRows16>(plan->power, plan->w_nt);
    default: return pick_w_mode<kRowsAny>(plan->power, plan->w_nt);
  }
}

// round-to-nearest split of a weight into two tf32-representable terms
void split_tf32(float w, float* hi, float* lo) {
  uint32_t bits;
  std::memcpy(&bits, &w, 4);
  bits = (bits + 0x1000u) & 0xffffe000u;
  std::memcpy(hi, &bits, 4);
  *lo = w - *hi;
}

template <int N>
KernelFn pick_ws(bool power, int mode) {
  switch (mode) {
    case kRows13: return power ? stft_ws_kernel<N, true, kRows13> : stft_ws_kernel<N, false, kRows13>;
    case kRows16: return power ? stft_ws_kernel<N, true, kRows16> : stft_ws_kernel<N, false, kRows16>;
    default: return power ? stft_ws_kernel<N, true, kRowsAny> : stft_ws_kernel<N, false, kRowsAny>;
  }
}

KernelFn pick_kernel(const pds_stft_plan* plan, int dtype) {
  if (plan->fast) {
    // the scalar-bank kernel is instantiated for N = 512 only; other sizes run stft_tc_kernel
    return plan->fused ? pick_fused<512>(plan->power, dtype, plan->row_mode) : nullptr;
  }
  if (plan->power) return dtype == PDS_I16 ? stft_direct_kernel<true, short> : stft_direct_kernel<true, float>;
  return dtype == PDS_I16 ? stft_direct_kernel<false, short> : stft_direct_kernel<false, float>;
}

void geometry_for(int N, int* G, int* R1) {
  const int NC = N / 2;
  *G = NC >= 1024 ? 32 : (NC >= 256 ? 16 : (NC >= 64 ? 8 : 4));
  *R1 = NC / *G;
}

size_t dtype_size(int dtype) { return dtype == PDS_I16 ? 2 : 4; }

}  // namespace

extern "C" int pds_stft_plan_create(const pds_stft_desc* d, int device, pds_stft_plan** out) {
  if (out) *out = nullptr;
  PDS_REQUIRE(d && out, "null descriptor or output pointer");
  PDS_REQUIRE(d->frame_length >= 1 && d->frame_shift >= 1, "frame_length/frame_shift must be positive");
  PDS_REQUIRE(d->dft_size >= d->frame_length, "dft_size (%d) < frame_length (%d)", d->dft_size,
              d->frame_length);
  PDS_REQUIRE(d->num_filts >= 1, "need at least one filter");
  PDS_REQUIRE(d->pad_left >= 0, "pad_left must be non-negative");
  PDS_REQUIRE(d->window && d->band_lo && d->band_len && d->band_off && d->weights, "null table");
  const int N = d->dft_size, L = d->frame_length, S = d->frame_shift, F = d->num_filts;
  const int K = N % 2 ? (N + 1) / 2 : N / 2 + 1;
  for (int f = 0; f < F; ++f)
    PDS_REQUIRE(d->band_lo[f] >= 0 && d->band_len[f] >= 0 && d->band_lo[f] + d->band_len[f] <= K,
                "band %d = [%d, %d) outside the %d bins", f, d->band_lo[f],
                d->band_lo[f] + d->band_len[f], K);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) {
    cudaGetLastError();
    set_error("no usable CUDA device %d (found %d); this library has no CPU fallback", device, ndev);
    return PDS_ERR_CUDA;
  }
  DeviceGuard guard(device);
  PDS_CUDA_CHECK(guard.status());

  pds_stft_plan* plan = new (std::nothrow) pds_stft_plan();
  if (!plan) return PDS_ERR_NOMEM;
  plan->device = device;
  plan->L = L, plan->S = S, plan->N = N, plan->K = K, plan->F = F;
  plan->C = F + (d->include_energy ? 1 : 0);
  plan->pad_left = d->pad_left;
  plan->power = d->use_power != 0;
  if (const char* force = getenv("PDS_STFT_KERNEL")) {
    plan->want_ws = force[0] == 'w';
    plan->want_scalar = force[0] == 's';
    if (force[0] == '1' || force[0] == 'w' || force[0] == 's') plan->variant = 1;
    if (force[0] == '2') plan->variant = 2;
    if (force[0] == 'x') plan->variant = 3;
    if (force[0] == 'y') plan->variant = 4;
  }
  cudaDeviceProp prop;
  cudaError_t err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(err));
    delete plan;
    return PDS_ERR_CUDA;
  }
  const size_t smem_cap = prop.sharedMemPerBlockOptin;

  // band tables, zero padded to whole groups of eight taps (the kernels' inner-loop trip)
  std::vector<int> lo(F), n4(F), off(F);
  int wtotal = 0;
  for (int f = 0; f < F; ++f) {
    lo[f] = d->band_lo[f];
    n4[f] = 2 * ((d->band_len[f] + 7) / 8);
    off[f] = wtotal;
    wtotal += n4[f] * 4;
  }
  // fused kernels: filter pairs (2i, 2i+1) padded to a common number of 8-tap groups, weights
  // interleaved per group as [8 taps of a | 8 taps of b]; an odd last filter pairs with zeros
  const int npairs = (F + 1) / 2;
  std::vector<int> desc(4 * (size_t)npairs);
  std::vector<float> pair_wt;
  int p_rows = K;
  for (int pi = 0; pi < npairs; ++pi) {
    const int fa = 2 * pi, fb = std::min(2 * pi + 1, F - 1);
    const bool has_b = 2 * pi + 1 < F;
    const int groups = std::max((d->band_len[fa] + 7) / 8, has_b ? (d->band_len[fb] + 7) / 8 : 0);
    desc[4 * pi + 0] = d->band_lo[fa] * kTileStride;
    desc[4 * pi + 1] = d->band_lo[fb] * kTileStride;
    desc[4 * pi + 2] = groups;
    desc[4 * pi + 3] = (int)pair_wt.size();
    p_rows = std::max(p_rows, std::max(d->band_lo[fa], d->band_lo[fb]) + 8 * groups);
    for (int g = 0; g < groups; ++g) {
      for (int j = 0; j < 8; ++j) {
        const int t = 8 * g + j;
        pair_wt.push_back(t < d->band_len[fa] ? d->weights[d->band_off[fa] + t] : 0.f);
      }
      for (int j = 0; j < 8; ++j) {
        const int t = 8 * g + j;
        pair_wt.push_back(has_b && t < d->band_len[fb] ? d->weights[d->band_off[fb] + t] : 0.f);
      }
    }
  }
  if (pair_wt.empty()) pair_wt.resize(4, 0.f);
  const int pair_total = (int)pair_wt.size();

  // tensor-core bank: items (8 filters x 16 frames x the group's 16-bin blocks), B fragments
  const int n_ntiles = (F + 7) / 8;
  const int n_warps = kThreads / 32;
  std::vector<int> tc_items;   // 4 ints per item: n0 | m0 << 16, first block, blocks, fragment offset
  std::vector<int> tc_wstart(n_warps + 1, 0);
  std::vector<float> tc_frags; // 4 floats per (block, k-step, lane)
  int tc_p_rows = ((K + 15) / 16) * 16;
  {
    struct Run { int n0, slot, blk0, nblk, frag; };
    std::vector<Run> runs;
    for (int j = 0; j < n_ntiles; ++j) {
      int lo_bin = K, hi_bin = 0;
      for (int f = 8 * j; f < std::min(F, 8 * j + 8); ++f)
        if (d->band_len[f] > 0) {
          lo_bin = std::min(lo_bin, d->band_lo[f]);
          hi_bin = std::max(hi_bin, d->band_lo[f] + d->band_len[f]);
        }
      const int blk0 = hi_bin > lo_bin ? lo_bin / 16 : 0;
      const int nblk = hi_bin > lo_bin ? (hi_bin + 15) / 16 - blk0 : 1;  // all-zero group: one block of zeros
      const int frag = (int)(tc_frags.size() / 4);
      for (int b = blk0; b < blk0 + nblk; ++b)
        for (int s2 = 0; s2 < 2; ++s2)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            const int f = 8 * j + g;
            float w[2] = {0.f, 0.f};
            for (int c = 0; c < 2; ++c) {
              const int bin = 16 * b + 4 * t + 2 * s2 + c;
              if (f < F && bin >= d->band_lo[f] && bin < d->band_lo[f] + d->band_len[f])
                w[c] = d->weights[d->band_off[f] + (bin - d->band_lo[f])];
            }
            float h0, l0, h1, l1;
            split_tf32(w[0], &h0, &l0);
            split_tf32(w[1], &h1, &l1);
            tc_frags.push_back(h0);
            tc_frags.push_back(h1);
            tc_frags.push_back(l0);
            tc_frags.push_back(l1);
          }
      for (int m0 = 0; m0 < (N > 1024 ? 16 : kTileFrames); m0 += 16) runs.push_back({8 * j, m0, blk0, nblk, frag});
      tc_p_rows = std::max(tc_p_rows, 16 * (blk0 + nblk));
    }
    // longest-processing-time-first deal to the warps
    std::vector<int> order(runs.size());
    for (size_t i = 0; i < runs.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return runs[a].nblk > runs[b].nblk; });
    std::vector<std::vector<int>> per_warp(n_warps);
    std::vector<int> load(n_warps, 0);
    for (int i : order) {
      int w = 0;
      for (int q = 1; q < n_warps; ++q)
        if (load[q] < load[w]) w = q;
      per_warp[w].push_back(i);
      load[w] += runs[i].nblk + 2;  // +2: per-item prologue / epilogue
    }
    // warp 0 also prepares the next tile and issues its TMA copy: give it the lightest share
    std::vector<int> by_load(n_warps);
    for (int w = 0; w < n_warps; ++w) by_load[w] = w;
    std::stable_sort(by_load.begin(), by_load.end(), [&](int a, int b) { return load[a] < load[b]; });
    for (int w = 0; w < n_warps; ++w) {
      tc_wstart[w] = (int)(tc_items.size() / 4);
      for (int i : per_warp[by_load[w]]) {
        const Run& r = runs[i];
        tc_items.push_back(r.n0 | (r.slot << 16));
        tc_items.push_back(r.blk0);
        tc_items.push_back(r.nblk);
        tc_items.push_back(r.frag);  // float4 index
      }
    }
    tc_wstart[n_warps] = (int)(tc_items.size() / 4);
    if (tc_frags.empty()) tc_frags.resize(4, 0.f);
  }
  const int tc_nitems = (int)(tc_items.size() / 4);

  // ---- pick the kernel: shared-memory FFT when the geometry allows, else direct DFT ------
  StftParams& p = plan->params;
  const bool pow2 = (N & (N - 1)) == 0;
  plan->fast = pow2 && N >= 256 && N <= 2048;  // odd frame shifts: tensor-core kernel only
  if (plan->fast) {
    geometry_for(N, &plan->G, &plan->R1);
    const int G = plan->G, R1 = plan->R1;
    const int rows = (L + 2 * G - 1) / (2 * G);  // stage-1 rows that carry samples
    plan->row_mode = rows == R1 ? kRows16 : (rows == (R1 * 13) / 16 ? kRows13 : kRowsAny);
    if (S % 2) plan->row_mode = kRowsAny;  // odd frame shift: the generic mode loads sample pairs with 4-byte loads
    p.rows_full = L / (2 * G);
    p.row_partial = (L % (2 * G)) != 0;
    const int fast_tile_frames = N > 1024 ? 16 : kTileFrames;  // TcTile<N>::kFrames
    p.span_max = (fast_tile_frames - 1) * S + L;
    // keep the weights in shared memory while that still leaves room for two CTAs per SM
    const SmemLayout with = fused_layout(N, G, R1, p.span_max, p_rows, npairs, plan->C, pair_total);
    const SmemLayout without = fused_layout(N, G, R1, p.span_max, p_rows, npairs, plan->C, 0);
    p.weights_in_smem = (size_t)with.total <= std::min<size_t>(smem_cap, 110 * 1024) ? 1 : 0;
    plan->smem_bytes = p.weights_in_smem ? with.total : without.total;
    plan->fused = N == 512 && S % 2 == 0 && plan->smem_bytes <= smem_cap;
    {
      size_t tc_bytes = 0;
      int tc_rows_max = 0, tc_items_max = 0;
      switch (N) {
        case 256: tc_bytes = TcSmem<256>::bytes(p.span_max, L), tc_rows_max = TcSmem<256>::kProws, tc_items_max = TcSmem<256>::kMaxItems; break;
        case 512: tc_bytes = TcSmem<512>::bytes(p.span_max, L), tc_rows_max = TcSmem<512>::kProws, tc_items_max = TcSmem<512>::kMaxItems; break;
        case 1024: tc_bytes = TcSmem<1024>::bytes(p.span_max, L), tc_rows_max = TcSmem<1024>::kProws, tc_items_max = TcSmem<1024>::kMaxItems; break;
        default: tc_bytes = TcSmem<2048>::bytes(p.span_max, L), tc_rows_max = TcSmem<2048>::kProws, tc_items_max = TcSmem<2048>::kMaxItems; break;
      }
      plan->tc = tc_bytes <= smem_cap && F < 65536 && tc_p_rows <= tc_rows_max && tc_nitems <= tc_items_max;
      plan->tc_smem_bytes = tc_bytes;
    }
    // warp-specialised kernel: dft_size 512, no fused pre-processing, <= 64 filters, everything in
    // shared memory (16-frame tiles)
    if (plan->tc && plan->variant == 5 && N == 512 && d->preemph == 0.f && d->dither == 0.f && F <= 8 * kWMaxNT) {
      const int w_span = (kWTile - 1) * S + L;
      const size_t w_bytes = WLayout::bytes(w_span, L, (int)(tc_frags.size() / 4));
      if (w_bytes <= smem_cap) {
        plan->w = true;
        plan->w_smem_bytes = w_bytes;
        plan->w_nt = (F + 7) / 8;
        p.span_max = w_span;
      }
    }
    if (!plan->fused && !plan->tc) plan->fast = false;  // huge frame shift / dft_size 2048: direct kernel
    if (plan->fast && plan->fused && d->preemph == 0.f && d->dither == 0.f) {
      const WsLayout ws = ws_layout(N, G, R1, p.span_max, p_rows, npairs, plan->C, pair_total);
      plan->ws = (size_t)ws.total <= smem_cap;
      plan->ws_smem_bytes = ws.total;
    }
  }
  if (!plan->fast) {
    plan->smem_bytes = sizeof(float) * (((L + 3) & ~3) + 2 * (size_t)N + ((K + 7 + 3) & ~3) + 8);
    if (plan->smem_bytes > smem_cap) {
      set_error("dft_size %d needs %zu bytes of shared memory (> %zu)", N, plan->smem_bytes, smem_cap);
      delete plan;
      return PDS_ERR_UNSUPPORTED;
    }
  }
  plan->tile_frames = plan->fast ? (N > 1024 || plan->w ? 16 : kTileFrames) : kDirectTileFrames;

  // ---- build the constant tables on the host (double precision trig) --------------------
  std::vector<float> wt(std::max(wtotal, 4), 0.f);
  for (int f = 0; f < F; ++f)
    for (int j = 0; j < d->band_len[f]; ++j) wt[off[f] + j] = d->weights[d->band_off[f] + j];
  std::vector<float> win(N, 0.f);
  for (int i = 0; i < L; ++i) win[i] = plan->fast ? 0.5f * d->window[i] : d->window[i];
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<float2> tw_stage, tw_split, tw_direct;
  if (plan->fast) {
    const int G = plan->G, R1 = plan->R1, NC = N / 2;
    tw_stage.resize((size_t)G * R1);
    for (int l = 0; l < G; ++l)
      for (int k1 = 0; k1 < R1; ++k1) {
        const double a = -two_pi * (double)((long long)l * k1 % NC) / NC;
        tw_stage[(size_t)l * R1 + k1] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    tw_split.resize((size_t)G * (R1 / 2));
    for (int l = 0; l < G; ++l)
      for (int m = 0; m < R1 / 2; ++m) {
        const double a = -two_pi * (double)(l + G * m) / N;
        tw_split[(size_t)l * (R1 / 2) + m] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
  } else {
    tw_direct.resize(N);
    for (int j = 0; j < N; ++j) {
      const double a = -two_pi * j / N;
      tw_direct[j] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  }

  // ---- one device blob -----------------------------------------------------------------
  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t o_win = 0, o_tws = align16(o_win + sizeof(float) * N);
  size_t o_twp = align16(o_tws + sizeof(float2) * tw_stage.size());
  size_t o_twd = align16(o_twp + sizeof(float2) * tw_split.size());
  size_t o_lo = align16(o_twd + sizeof(float2) * tw_direct.size());
  size_t o_n4 = align16(o_lo + sizeof(int) * F);
  size_t o_off = align16(o_n4 + sizeof(int) * F);
  size_t o_desc = align16(o_off + sizeof(int) * F);
  size_t o_pwt = align16(o_desc + sizeof(int) * 4 * npairs);
  size_t o_wt = align16(o_pwt + sizeof(float) * pair_wt.size());
  size_t o_tci = align16(o_wt + sizeof(float) * wt.size());
  size_t o_tcw = align16(o_tci + sizeof(int) * tc_items.size());
  size_t o_tcf = align16(o_tcw + sizeof(int) * tc_wstart.size());
  size_t blob_bytes = align16(o_tcf + sizeof(float) * tc_frags.size());
  std::vector<unsigned char> blob(blob_bytes, 0);
  std::memcpy(blob.data() + o_win, win.data(), sizeof(float) * N);
  if (!tw_stage.empty()) std::memcpy(blob.data() + o_tws, tw_stage.data(), sizeof(float2) * tw_stage.size());
  if (!tw_split.empty()) std::memcpy(blob.data() + o_twp, tw_split.data(), sizeof(float2) * tw_split.size());
  if (!tw_direct.empty()) std::memcpy(blob.data() + o_twd, tw_direct.data(), sizeof(float2) * tw_direct.size());
  std::memcpy(blob.data() + o_lo, lo.data(), sizeof(int) * F);
  std::memcpy(blob.data() + o_n4, n4.data(), sizeof(int) * F);
  std::memcpy(blob.data() + o_off, off.data(), sizeof(int) * F);
  std::memcpy(blob.data() + o_desc, desc.data(), sizeof(int) * 4 * npairs);
  std::memcpy(blob.data() + o_pwt, pair_wt.data(), sizeof(float) * pair_wt.size());
  std::memcpy(blob.data() + o_wt, wt.data(), sizeof(float) * wt.size());
  if (!tc_items.empty()) std::memcpy(blob.data() + o_tci, tc_items.data(), sizeof(int) * tc_items.size());
  std::memcpy(blob.data() + o_tcw, tc_wstart.data(), sizeof(int) * tc_wstart.size());
  std::memcpy(blob.data() + o_tcf, tc_frags.data(), sizeof(float) * tc_frags.size());
  err = cudaMalloc(&plan->d_blob, blob_bytes);
  if (err == cudaSuccess) err = cudaMemcpy(plan->d_blob, blob.data(), blob_bytes, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    set_error("uploading plan tables failed: %s", cudaGetErrorString(err));
    pds_stft_plan_destroy(plan);
    return PDS_ERR_CUDA;
  }
  unsigned char* base = static_cast<unsigned char*>(plan->d_blob);
  p.window = reinterpret_cast<const float*>(base + o_win);
  p.tw_stage = reinterpret_cast<const float2*>(base + o_tws);
  p.tw_split = reinterpret_cast<const float2*>(base + o_twp);
  p.tw_direct = reinterpret_cast<const float2*>(base + o_twd);
  p.band_lo = reinterpret_cast<const int*>(base + o_lo);
  p.band_n4 = reinterpret_cast<const int*>(base + o_n4);
  p.band_off = reinterpret_cast<const int*>(base + o_off);
  p.pair_desc = reinterpret_cast<const int4*>(base + o_desc);
  p.pair_weights = reinterpret_cast<const float*>(base + o_pwt);
  p.npairs = npairs;
  p.tc_items = reinterpret_cast<const int4*>(base + o_tci);
  p.tc_wstart = reinterpret_cast<const int*>(base + o_tcw);
  p.tc_frags = reinterpret_cast<const float4*>(base + o_tcf);
  p.tc_nitems = tc_nitems;
  p.w_frag4 = (int)(tc_frags.size() / 4);
  p.w_probe = getenv("PDS_W_PROBE") ? atoi(getenv("PDS_W_PROBE")) : 0;
  p.tc_p_rows = tc_p_rows;
  p.p_rows = p_rows;
  p.weights = reinterpret_cast<const float*>(base + o_wt);
  p.weights_total = pair_total;
  p.L = L, p.S = S, p.N = N, p.K = K, p.F = F, p.C = plan->C;
  p.include_energy = d->include_energy ? 1 : 0;
  p.use_log = d->use_log ? 1 : 0;
  p.log_floor = d->log_floor;
  p.inv_L = 1.0f / (float)L;
  p.preemph = d->preemph;
  p.dither = d->dither;
  p.dither_first = d->dither_first ? 1 : 0;

  // opt in to the dynamic shared memory for every instantiation this plan may launch
  for (int dt = 0; dt < 2; ++dt) {
    KernelFn fn = pick_kernel(plan, dt);
    if (!fn) continue;  // tensor-core kernel only
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", plan->smem_bytes, cudaGetErrorString(err));
      pds_stft_plan_destroy(plan);
      return PDS_ERR_CUDA;
    }
  }
  if (plan->tc) {
    for (int dt = 0; dt < 2 && plan->tc; ++dt) {
      err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_tc(plan, dt)),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->tc_smem_bytes);
      if (err != cudaSuccess) {
        cudaGetLastError();
        plan->tc = false;
      }
    }
  }
  if (plan->w) {
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_w(plan)),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->w_smem_bytes);
    if (err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", plan->w_smem_bytes, cudaGetErrorString(err));
      pds_stft_plan_destroy(plan);
      return PDS_ERR_CUDA;
    }
  }
  if (plan->ws) {
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_ws<512>(plan->power, plan->row_mode)),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->ws_smem_bytes);
    if (err != cudaSuccess) {
      cudaGetLastError();
      plan->ws = false;
    }
  }
  plan->num_sms = prop.multiProcessorCount;
  int occ = 1;
  if (pick_kernel(plan, PDS_F32))
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void*>(pick_kernel(plan, PDS_F32)),
                                                  plan->fast ? kThreads : kDirectThreads, plan->smem_bytes);
  plan->grid_limit = prop.multiProcessorCount * std::max(1, occ);
  if (plan->tc) {
    int occ_tc = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tc, reinterpret_cast<const void*>(pick_tc(plan, PDS_F32)),
                                                  kThreads, plan->tc_smem_bytes);
    plan->tc_grid_limit = prop.multiProcessorCount * std::max(1, occ_tc);
  }
  *out = plan;
  return PDS_OK;
}

extern "C" void pds_stft_plan_destroy(pds_stft_plan* plan) {
  if (!plan) return;
  DeviceGuard guard(plan->device);
  if (plan->d_blob) cudaFree(plan->d_blob);
  if (plan->d_sig) cudaFree(plan->d_sig);
  if (plan->d_tiles) cudaFree(plan->d_tiles);
  if (plan->d_out) cudaFree(plan->d_out);
  delete plan;
}

extern "C" int pds_stft_num_coeffs(const pds_stft_plan* plan) { return plan ? plan->C : 0; }
extern "C" int pds_stft_tile_frames(const pds_stft_plan* plan) { return plan ? plan->tile_frames : 0; }
extern "C" int pds_stft_is_fast_path(const pds_stft_plan* plan) { return plan && plan->fast ? 1 : 0; }

extern "C" int64_t pds_stft_num_frames(const pds_stft_plan* plan, int64_t sig_len) {
  if (!plan || sig_len < plan->L / 2 + 1) return 0;
  return (sig_len + plan->S / 2) / plan->S;
}

extern "C" int pds_stft_layout(const pds_stft_plan* plan, int64_t n_utts, const int64_t* sig_len,
                               int64_t* frame_off, int64_t* n_tiles) {
  PDS_REQUIRE(plan && sig_len && frame_off && n_tiles && n_utts >= 0, "bad argument");
  int64_t rows = 0, tiles = 0;
  const int64_t tf = plan->tile_frames;
  for (int64_t u = 0; u < n_utts; ++u) {
    PDS_REQUIRE(sig_len[u] >= 0 && sig_len[u] < (int64_t)1 << 31, "utterance %lld has length %lld",
                (long long)u, (long long)sig_len[u]);
    frame_off[u] = rows;
    const int64_t t = pds_stft_num_frames(plan, sig_len[u]);
    rows += t;
    tiles += (t + tf - 1) / tf;
  }
  frame_off[n_utts] = rows;
  *n_tiles = tiles;
  return PDS_OK;
}

extern "C" int pds_stft_fill_tiles(const pds_stft_plan* plan, int64_t n_utts, const int64_t* sig_off,
                                   const int64_t* sig_len, const int64_t* frame_off, pds_tile* tiles) {
  PDS_REQUIRE(plan && sig_off && sig_len && frame_off && (tiles || n_utts == 0), "bad argument");
  const int64_t tf = plan->tile_frames;
  int64_t n = 0;
  for (int64_t u = 0; u < n_utts; ++u) {
    const int64_t t_total = frame_off[u + 1] - frame_off[u];
    for (int64_t t0 = 0; t0 < t_total; t0 += tf) {
      pds_tile& tile = tiles[n++];
      tile.sig_off = sig_off[u];
      tile.sig_len = (int32_t)sig_len[u];
      tile.start = (int32_t)(t0 * plan->S - plan->pad_left);
      tile.nframes = (int32_t)std::min<int64_t>(tf, t_total - t0);
      tile.utt = (int32_t)u;
      tile.out_row = frame_off[u] + t0;
    }
  }
  return PDS_OK;
}

extern "C" int pds_stft_fill_tiles_range(const pds_stft_plan* plan, int64_t sig_off, int64_t buf_len,
                                         int64_t buf_origin, int64_t first_frame, int64_t nframes,
                                         int64_t out_row, pds_tile* tiles, int64_t* n_tiles) {
  PDS_REQUIRE(plan && n_tiles && nframes >= 0 && buf_len >= 0, "bad argument");
  const int64_t tf = plan->tile_frames;
  int64_t n = 0;
  for (int64_t t0 = 0; t0 < nframes; t0 += tf) {
    if (tiles) {
      pds_tile& tile = tiles[n];
      tile.sig_off = sig_off;
      tile.sig_len = (int32_t)buf_len;
      tile.start = (int32_t)((first_frame + t0) * plan->S - plan->pad_left - buf_origin);
      tile.nframes = (int32_t)std::min<int64_t>(tf, nframes - t0);
      tile.utt = 0;
      tile.out_row = out_row + t0;
    }
    ++n;
  }
  *n_tiles = n;
  return PDS_OK;
}

extern "C" int pds_stft_run(pds_stft_plan* plan, const void* d_signal, int sig_dtype,
                            const pds_tile* d_tiles, int64_t n_tiles, float* d_out, uint64_t seed,
                            void* stream) {
  PDS_REQUIRE(plan, "null plan");
  PDS_REQUIRE(sig_dtype == PDS_F32 || sig_dtype == PDS_I16,
              "sample dtype %d is not float32 / int16 (convert float64 on the host side)", sig_dtype);
  if (n_tiles == 0) return PDS_OK;
  PDS_REQUIRE(d_signal && d_tiles && d_out && n_tiles > 0, "null buffer");
  StftParams p = plan->params;
  p.sig = d_signal;
  p.tiles = d_tiles;
  p.n_tiles = n_tiles;
  p.out = d_out;
  p.seed = seed;
  // PDS_STFT_KERNEL=ws opts in to the warp-specialised kernel (A/B runs, tests).  It is not the
  // default yet: its three bank warps are the bottleneck (profiles/), the phased kernel is faster.
  if (plan->ws && sig_dtype == PDS_F32 && plan->want_ws) {
    const int grid = (int)std::min<int64_t>(n_tiles, plan->num_sms);
    pick_ws<512>(plan->power, plan->row_mode)<<<grid, kWsThreads, plan->ws_smem_bytes,
                                                static_cast<cudaStream_t>(stream)>>>(p);
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  // PDS_STFT_KERNEL=scalar falls back to the CUDA-core bank kernel (A/B runs, tests)
  if (plan->w && sig_dtype == PDS_F32) {
    PDS_REQUIRE((reinterpret_cast<uintptr_t>(d_tiles) & 15u) == 0, "d_tiles must be 16-byte aligned");
    PDS_REQUIRE(n_tiles < ((int64_t)1 << 30), "at most 2^30 tiles per launch (got %lld)", (long long)n_tiles);
    const int grid = (int)std::min<int64_t>((n_tiles + kWGroups - 1) / kWGroups, plan->num_sms);
    pick_w(plan)<<<grid, kWThreads, plan->w_smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  if (plan->tc && !(plan->want_scalar && plan->fused)) {
    PDS_REQUIRE((reinterpret_cast<uintptr_t>(d_tiles) & 15u) == 0, "d_tiles must be 16-byte aligned");
    PDS_REQUIRE(n_tiles < ((int64_t)1 << 30), "at most 2^30 tiles per launch (got %lld)", (long long)n_tiles);
    const int grid = (int)std::min<int64_t>(n_tiles, plan->tc_grid_limit);
    pick_tc(plan, sig_dtype)<<<grid, kThreads, plan->tc_smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  KernelFn fn = pick_kernel(plan, sig_dtype);
  PDS_REQUIRE(fn, "no kernel for this plan");
  const int grid = (int)std::min<int64_t>(n_tiles, plan->grid_limit);
  const int threads = plan->fast ? kThreads : kDirectThreads;
  fn<<<grid, threads, plan->smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}

namespace {
int ensure(void** ptr, size_t* cap, size_t need) {
  if (*cap >= need) return PDS_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  PDS_CUDA_CHECK(cudaMalloc(ptr, need));
  *cap = need;
  return PDS_OK;
}
}  // namespace

extern "C" int pds_stft_compute_host(pds_stft_plan* plan, const void* h_signal, int sig_dtype,
                                     int64_t total_samples, int64_t n_utts, const int64_t* sig_off,
                                     const int64_t* sig_len, float* h_out, int64_t out_capacity_rows,
                                     int64_t* frame_off, uint64_t seed) {
  PDS_REQUIRE(plan && sig_off && sig_len && frame_off && n_utts >= 0 && total_samples >= 0, "bad argument");
  PDS_REQUIRE(sig_dtype == PDS_F32 || sig_dtype == PDS_I16,
              "sample dtype %d is not float32 / int16 (convert float64 on the host side)", sig_dtype);
  for (int64_t u = 0; u < n_utts; ++u)
    PDS_REQUIRE(sig_off[u] >= 0 && sig_off[u] + sig_len[u] <= total_samples,
                "utterance %lld lies outside the packed buffer", (long long)u);
  int64_t n_tiles = 0;
  int rc = pds_stft_layout(plan, n_utts, sig_len, frame_off, &n_tiles);
  if (rc != PDS_OK) return rc;
  const int64_t rows = frame_off[n_utts];
  PDS_REQUIRE(rows <= out_capacity_rows, "output holds %lld rows, need %lld",
              (long long)out_capacity_rows, (long long)rows);
  if (rows == 0) return PDS_OK;
  PDS_REQUIRE(h_signal && h_out, "null buffer");
  DeviceGuard guard(plan->device);
  PDS_CUDA_CHECK(guard.status());
  std::vector<pds_tile> tiles((size_t)n_tiles);
  rc = pds_stft_fill_tiles(plan, n_utts, sig_off, sig_len, frame_off, tiles.data());
  if (rc != PDS_OK) return rc;
  size_t tiles_bytes = plan->d_tiles_cap * sizeof(pds_tile);
  rc = ensure(&plan->d_sig, &plan->d_sig_bytes, (size_t)total_samples * dtype_size(sig_dtype) + 16);
  if (rc == PDS_OK) {
    void* t = plan->d_tiles;
    rc = ensure(&t, &tiles_bytes, (size_t)n_tiles * sizeof(pds_tile));
    plan->d_tiles = static_cast<pds_tile*>(t);
    plan->d_tiles_cap = tiles_bytes / sizeof(pds_tile);
  }
  if (rc == PDS_OK) {
    void* o = plan->d_out;
    rc = ensure(&o, &plan->d_out_bytes, (size_t)rows * plan->C * sizeof(float));
    plan->d_out = static_cast<float*>(o);
  }
  if (rc != PDS_OK) return rc;
  PDS_CUDA_CHECK(cudaMemcpyAsync(plan->d_sig, h_signal, (size_t)total_samples * dtype_size(sig_dtype),
                                 cudaMemcpyHostToDevice, 0));
  PDS_CUDA_CHECK(cudaMemcpyAsync(plan->d_tiles, tiles.data(), (size_t)n_tiles * sizeof(pds_tile),
                                 cudaMemcpyHostToDevice, 0));
  rc = pds_stft_run(plan, plan->d_sig, sig_dtype, plan->d_tiles, n_tiles, plan->d_out, seed, nullptr);
  if (rc != PDS_OK) return rc;
  PDS_CUDA_CHECK(cudaMemcpyAsync(h_out, plan->d_out, (size_t)rows * plan->C * sizeof(float),
                                 cudaMemcpyDeviceToHost, 0));
  PDS_CUDA_CHECK(cudaStreamSynchronize(0));
  return PDS_OK;
}
