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

#pragma once
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
  // Bluestein path (stft_bluestein_kernel): window * conj(chirp) [L], DFT_1024 of the chirp kernel / 1024,
  // and the 1024-point transform's twiddles W_1024^(lane * k1) at [k1][lane]
  const float2* blue_aw;
  const float2* blue_b;
  const float2* blue_tw;
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
  // stft_umma_kernel (stft_umma.cuh): the constant DFT operand as fp16 (hi, lo) terms in its shared-memory
  // layout, the unscaled zero-padded window [N], bank weights as bf16 m16n8k16 A fragments in the kernel's
  // bin order ({a0 .. a3} of the high terms for 32 lanes, then of the low terms), and per (16 values of k1,
  // bin group) the mask of 16-filter tiles with non-zero weights and the index of their first fragment
  const void* um_bmat;
  const float* um_window;
  const void* um_frags;
  const int* um_masks;
  const int* um_offs;
  int um_kch;                  // 8-sample groups of a residue that carry samples (K / 8), even
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
    if (spins > (1u << 22)) {
#ifdef PDS_DEBUG_WAIT
      printf("mbar_wait timeout: block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
#endif
      __trap();
    }
}
// The same for warps that expect to wait (idle consumers): back off between polls so that the
// spinning does not take issue slots from the warps that do the work.
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
  for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    __nanosleep(40);
    if (spins > (1u << 24)) {
#ifdef PDS_DEBUG_WAIT
      printf("mbar_wait_idle timeout: block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
#endif
      __trap();
    }
  }
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

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x0 (lower half) and x1 (upper half) as bf16, and the same for their rounding residuals
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& p1, uint32_t& p2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(x1), "f"(x0));
  const float r0 = x0 - __uint_as_float(p1 << 16);
  const float r1 = x1 - __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
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

// One bank item on bf16 m16n8k16 MMAs: NM 16-frame halves (frames m_first + 16 m) x eight filters x `left`
// sixteen-bin blocks.  A 16-bin block is ONE MMA k-step (k = 2t, 2t+1, 2t+8, 2t+9 <-> bins 4t .. 4t+3: the
// same conflict-free loads as the tf32 path); spectra and weights are split into two bf16 terms each and
// the products p1 w1, p2 w1, p1 w2 are accumulated: three MMAs per block instead of six, weight fragments
// half the size; relative error <= 3 * 2^-17 on sums of non-negative terms (measured 7e-6; tolerance 1e-4).
template <int NM, int TS>
__device__ __forceinline__ void bank_item_bf16(const float* __restrict__ pa, const float4* __restrict__ fr, int left,
                                               float* __restrict__ out, int m_first, int n0, int g, int t, int nframes,
                                               const StftParams& p, bool use_log, float log_floor) {
  float acc[NM][2][4];  // [half][main | correction][fragment]
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][q][i] = 0.f;
#pragma unroll 2
  do {
    const float4 f = ldg_keep(fr);
    const uint32_t w1b0 = __float_as_uint(f.x), w1b1 = __float_as_uint(f.y);
    const uint32_t w2b0 = __float_as_uint(f.z), w2b1 = __float_as_uint(f.w);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const float* __restrict__ pm = pa + 16 * m;
      uint32_t p1[4], p2[4];
      split_bf16x2(pm[0], pm[TS], p1[0], p2[0]);                   // row g,     bins 4t, 4t+1
      split_bf16x2(pm[8], pm[TS + 8], p1[1], p2[1]);               // row g + 8
      split_bf16x2(pm[2 * TS], pm[3 * TS], p1[2], p2[2]);          // row g,     bins 4t+2, 4t+3
      split_bf16x2(pm[2 * TS + 8], pm[3 * TS + 8], p1[3], p2[3]);  // row g + 8
      mma_bf16(acc[m][0], p1[0], p1[1], p1[2], p1[3], w1b0, w1b1);
      mma_bf16(acc[m][1], p2[0], p2[1], p2[2], p2[3], w1b0, w1b1);
      mma_bf16(acc[m][1], p1[0], p1[1], p1[2], p1[3], w2b0, w2b1);
    }
    pa += 16 * TS;
    fr += 32;
  } while (--left > 0);
  const int C = p.C;
  const bool c0 = n0 + 2 * t < p.F, c1 = n0 + 2 * t + 1 < p.F;
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = acc[m][1][i] + acc[m][0][i];
      if (use_log) v[i] = fast_log(fmaxf(v[i], log_floor));
    }
    const int row = m_first + 16 * m;
    float* __restrict__ r0 = out + row * C;
    float* __restrict__ r1 = r0 + 8 * C;
    if (row + g < nframes) {
      if (c0) __stcs(r0, v[0]);
      if (c1) __stcs(r0 + 1, v[1]);
    }
    if (row + g + 8 < nframes) {
      if (c0) __stcs(r1, v[2]);
      if (c1) __stcs(r1 + 1, v[3]);
    }
  }
}

// An item with this m0 covers BOTH 16-frame halves of a 32-frame tile: one fetch of a weight
// fragment feeds two A tiles.  Used when there are at least as many filter groups as warps (dense
// banks: gammatone-64 has 8 groups x 17 blocks = 136 KB of fragments, far beyond the L1; with one
// item per (group, half) every tile pulled them through the L2 twice: 8.5 KB per frame).
constexpr int kBothHalves = 0x7fff;
constexpr int kBothHalvesBf16 = 0x7ffe;  // the same with bf16 fragments (bank_item_bf16)
constexpr int kHalfBf16 = 0x4000;        // flag on m0: one half, bf16 fragments

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
    if (m0 == kBothHalvesBf16) {
      bank_item_bf16<2, TS>(s_P + d.y * (16 * TS) + lane_p, frags + d.w + lane, d.z, out_tile + n0 + lane_o, 0, n0, g, t,
                            nframes, p, use_log, log_floor);
      continue;
    }
    if (m0 & kHalfBf16) {
      const int mh = m0 & ~kHalfBf16;
      if (mh >= nframes) continue;
      bank_item_bf16<1, TS>(s_P + d.y * (16 * TS) + mh + lane_p, frags + d.w + lane, d.z, out_tile + n0 + lane_o, mh, n0,
                            g, t, nframes, p, use_log, log_floor);
      continue;
    }
    if (m0 == kBothHalves) {
      const float* __restrict__ pa = s_P + d.y * (16 * TS) + lane_p;
      const float4* __restrict__ fr = frags + d.w + lane;
      float acc[2][2][2][4];  // [half][k-step parity][main | correction][fragment]
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][s][q][i] = 0.f;
      int left = d.z;  // >= 1 (host)
      do {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float4 f = ldg_keep(fr + 32 * s);
          const uint32_t whi0 = __float_as_uint(f.x), whi1 = __float_as_uint(f.y);
          const uint32_t wlo0 = __float_as_uint(f.z), wlo1 = __float_as_uint(f.w);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const float* __restrict__ pm = pa + 16 * m;
            const float a[4] = {pm[(2 * s) * TS], pm[(2 * s) * TS + 8], pm[(2 * s + 1) * TS], pm[(2 * s + 1) * TS + 8]};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              hi[i] = __float_as_uint(a[i]) & 0xffffe000u;
              lo[i] = __float_as_uint(a[i] - __uint_as_float(hi[i]));
            }
            mma_tf32(acc[m][s][0], hi[0], hi[1], hi[2], hi[3], whi0, whi1);
            mma_tf32(acc[m][s][1], lo[0], lo[1], lo[2], lo[3], whi0, whi1);
            mma_tf32(acc[m][s][1], hi[0], hi[1], hi[2], hi[3], wlo0, wlo1);
          }
        }
        pa += 16 * TS;
        fr += 64;
      } while (--left > 0);
      const bool c0 = n0 + 2 * t < p.F, c1 = n0 + 2 * t + 1 < p.F;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = (acc[m][0][1][i] + acc[m][1][1][i]) + (acc[m][0][0][i] + acc[m][1][0][i]);  // small terms first
          if (use_log) v[i] = fast_log(fmaxf(v[i], log_floor));
        }
        float* __restrict__ r0 = out_tile + (16 * m * C + n0 + lane_o);
        float* __restrict__ r1 = r0 + 8 * C;
        if (16 * m + g < nframes) {
          if (c0) __stcs(r0, v[0]);
          if (c1) __stcs(r0 + 1, v[1]);
        }
        if (16 * m + g + 8 < nframes) {
          if (c0) __stcs(r1, v[2]);
          if (c1) __stcs(r1 + 1, v[3]);
        }
      }
      continue;
    }
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
constexpr int kWHalves = 1;           // bank warps per spectrum stage (2: each takes half of the filter groups)
constexpr int kWBankWarps = 6 * kWHalves;
constexpr int kWThreads = 32 * (12 + kWBankWarps + 1);  // 12 transform warps + bank warps + 1 producer warp
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
  static constexpr int oTab = oCtl + kWGroups * kWRing * 16;    // block masks [2][32], per-filter-group offsets [8], owned groups [2]
  static constexpr int oP = oTab + 96;                          // spectra [3][2][272][18]
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
                                       const StftParams& p, float* __restrict__ out_tile, int nframes, int own) {
  constexpr int TS = kWStride;
  constexpr int NS = 1;  // one accumulator pair per filter group (104 registers per thread)
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
    const bool mine = (own >> n) & 1;
    const bool c0 = mine && 8 * n + 2 * t < p.F, c1 = mine && 8 * n + 2 * t + 1 < p.F;
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
  int* const s_adj = s_mask + 64;
  int* const s_own = s_mask + 72;
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
  if (tid < 96) s_mask[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    // The bank items with m0 == 0 describe each filter group (eight filters) once.  The groups are
    // dealt to the two bank warps of a spectrum stage, longest band first to the lighter warp.
    int blk0[kWMaxNT], nblk[kWMaxNT];
    for (int n = 0; n < kWMaxNT; ++n) blk0[n] = 0, nblk[n] = 0;
    for (int i = 0; i < p.tc_nitems; ++i) {
      const int4 d = p.tc_items[i];
      if ((d.x >> 16) == 0 || (d.x >> 16) == kBothHalves) {
        const int n = (d.x & 0xffff) >> 3;
        blk0[n] = d.y, nblk[n] = d.z;
        s_adj[n] = d.w - d.y * 64;
      }
    }
    int load[2] = {0, 0};
    unsigned done = 0;
    for (int round = 0; round < kWMaxNT; ++round) {
      int best = -1;
      for (int n = 0; n < kWMaxNT; ++n)
        if (!(done & (1u << n)) && nblk[n] > 0 && (best < 0 || nblk[n] > nblk[best])) best = n;
      if (best < 0) break;
      done |= 1u << best;
      const int half = kWHalves > 1 && load[1] < load[0] ? 1 : 0;
      load[half] += nblk[best] + 1;
      s_own[half] |= 1 << best;
      for (int b = blk0[best]; b < blk0[best] + nblk[best]; ++b) s_mask[half * 32 + b] |= 1 << best;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], kWHalves);
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
      mbar_wait_idle(&x_full[st], ph);
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
        mbar_wait_idle(&p_empty[st], ph ^ 1);
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
      mbar_wait_idle(&p_empty[st], ph ^ 1);  // the bank warp is done with the tile two iterations back
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
  } else if (warp < 4 * kWGroups + kWBankWarps) {
    // =============================== bank warps =========================================
    // kWHalves bank warps serve one spectrum stage of one group (iterations j = stage, stage + 2,
    // ...), each with its share of the filter groups.  A tile keeps a warp busy for longer than a group
    // needs to produce the next one (the tf32 MMA chain is latency bound: 8 cycles per MMA and
    // warp, 20 cycles per dependent pair, tools/ubench/mma_rate.cu), so both stages of a group are
    // in work at once.  One fixed pair of warps per stage also keeps every parity wait at most one
    // phase away from its barrier.
    const int b = warp - 4 * kWGroups;
    const int st = b / kWHalves, half = b % kWHalves;  // st = g * 2 + (j & 1)
    const int g = st >> 1;
    const int first = blockIdx.x * kWGroups + g;
    const int n_iter = first < n_tiles ? (n_tiles - first + tile_step - 1) / tile_step : 0;
    const int own = s_own[half];
    for (int j = st & 1; j < n_iter; j += 2) {
      mbar_wait_idle(&p_full[st], (j >> 1) & 1);
      const int* __restrict__ c = s_ctl + (g * kWRing + (j & (kWRing - 1))) * 16;
      const int4 c0 = *reinterpret_cast<const int4*>(c);
      float* __restrict__ out_tile = p.out + (((long long)c0.w << 32) | (unsigned)c0.z);
      if (p.w_probe != 1 && own)
        bank_w<NT>(lane, s_P + st * (kWRows * TS), s_mask + half * 32, s_adj, s_frag, p, out_tile, c0.x, own);
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_empty[st]);
    }
  } else if (warp == 4 * kWGroups + kWBankWarps && lane < kWGroups) {
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
      if (j >= 2) mbar_wait_idle(&x_empty[st], ((j - 2) >> 1) & 1);  // all four warps have consumed the stage
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
// DFT sizes that are not a power of two (pad_to_nearest_power_of_two = false, compute.py:344-347;
// N = 400 for 25 ms at 16 kHz is the common case), N <= 512: Bluestein's algorithm on the in-register
// 1024-point FFT.  With the chirp c[n] = exp(i pi n^2 / N),
//     X[k] = conj(c[k]) * sum_n (x[n] w[n] conj(c[n])) * c[k - n],
// a circular convolution of length 1024 >= 2 N - 1: FFT of a[n] = x[n] w[n] conj(c[n]), product with
// the precomputed transform of the chirp kernel (scaled by 1/1024), inverse FFT -- as a forward FFT
// of the conjugate, since only |X[k]| = |y[k]| is needed the final conj(c[k]) drops out as well.
// One warp per frame (32 lanes x 32 registers, element index = lane + 32 * register in and out, so
// the product needs no reordering); eight frames per tile; the bank is the banded dot product of
// the direct kernel.  About 20x fewer flops than the O(L K) direct DFT it replaces.
// ------------------------------------------------------------------------------------------
constexpr int kBlueThreads = 256;
constexpr int kBlueTileFrames = kBlueThreads / 32;
constexpr int kBlueM = 1024;
using BlueGeo = FftGeom<2 * kBlueM>;  // 1024 complex points: G = 32 lanes, R1 = 32 registers
static_assert(BlueGeo::G == 32 && BlueGeo::R1 == 32 && BlueGeo::NSUB == 1, "one warp per transform");

struct BlueSmem {
  int x, tw, aw, scr, P, total;  // float offsets; total in bytes
};
__host__ __device__ inline BlueSmem blue_layout(int span_max, int L, int K) {
  BlueSmem l;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
  l.x = take(span_max + 8);
  l.tw = take(2 * kBlueM);
  l.aw = take(2 * L);
  l.scr = take(2 * BlueGeo::SCR_FLOAT2 * kBlueTileFrames);
  l.P = take(kBlueTileFrames * ((K + 3) & ~3));
  l.total = o * 4;
  return l;
}

// 1024-point forward FFT of z (element index = lane + 32 * register, in and out)
__device__ __forceinline__ void fft1024_warp(cplx (&z)[32], int lane, const float2* __restrict__ s_tw,
                                             cplx* __restrict__ scr) {
  Dft<32>::run(z);
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) z[k1] = cmul(z[k1], s_tw[k1 * 32 + lane]);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) scr[lane * BlueGeo::SCR_STRIDE + k1] = z[k1];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) z[n2] = scr[n2 * BlueGeo::SCR_STRIDE + lane];
  __syncwarp();
  Dft<32>::run(z);
}

template <bool POWER, typename T>
__global__ void __launch_bounds__(kBlueThreads, 2) stft_bluestein_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ __align__(16) float smem[];
  const BlueSmem lay = blue_layout(p.span_max, p.L, p.K);
  float* const s_x = smem + lay.x;
  float2* const s_tw = reinterpret_cast<float2*>(smem + lay.tw);
  float2* const s_aw = reinterpret_cast<float2*>(smem + lay.aw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cplx* const scr = reinterpret_cast<cplx*>(smem + lay.scr) + warp * BlueGeo::SCR_FLOAT2;
  const int kpad = (p.K + 3) & ~3;
  float* const sP = smem + lay.P + warp * kpad;

  for (int i = tid; i < kBlueM; i += kBlueThreads) s_tw[i] = p.blue_tw[i];
  for (int i = tid; i < p.L; i += kBlueThreads) s_aw[i] = p.blue_aw[i];

  for (long long tile_idx = blockIdx.x; tile_idx < p.n_tiles; tile_idx += gridDim.x) {
    const pds_tile tile = p.tiles[tile_idx];
    const int span = (tile.nframes - 1) * p.S + p.L;
    __syncthreads();  // the previous tile's samples are no longer needed (and the tables are loaded)
    stage_samples_slow<T, kBlueThreads>(s_x, p, tile, span, 0, 0);
    __syncthreads();
    if (warp >= tile.nframes) continue;
    const float* __restrict__ fx = s_x + warp * p.S;
    cplx z[32];
    float e = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int n = lane + 32 * r;
      if (n < p.L) {
        const float x = fx[n];
        const float2 a = s_aw[n];
        e = fmaf(x, x, e);
        z[r] = cmake(x * a.x, x * a.y);
      } else {
        z[r] = cmake(0.f, 0.f);
      }
    }
    fft1024_warp(z, lane, s_tw, scr);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const cplx y = cmul(z[r], __ldg(p.blue_b + r * 32 + lane));
      z[r] = cmake(cre(y), -cim(y));
    }
    fft1024_warp(z, lane, s_tw, scr);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int k = lane + 32 * r;
      if (k < p.K) {
        const float pw = cnorm(z[r]);
        sP[k] = POWER ? pw : sqrtf(pw);
      }
    }
    __syncwarp();
    float* __restrict__ dst = p.out + (tile.out_row + warp) * p.C;
    for (int f = lane; f < p.F; f += 32) {  // compute.py:416-460 as a banded dot product
      const float* __restrict__ wt = p.weights + p.band_off[f];
      const float* __restrict__ pp = sP + p.band_lo[f];
      const int n = min(p.band_n4[f] * 4, p.K - p.band_lo[f]);  // the zero padding of the taps is not read
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(pp[j], wt[j], acc);
      if (p.use_log) acc = __logf(fmaxf(acc, p.log_floor));
      dst[p.include_energy + f] = acc;
    }
    if (p.include_energy) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
      if (lane == 0) {
        float v = e * p.inv_L;
        if (!POWER) v = sqrtf(v);
        if (p.use_log) v = __logf(fmaxf(v, p.log_floor));
        dst[0] = v;
      }
    }
    __syncwarp();
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


namespace pds {

// ---- kernel pickers: each is defined in its own translation unit (stft_k_*.cu), so that the many
// template instantiations compile in parallel ---------------------------------------------------
using KernelFn = void (*)(const StftParams);
KernelFn pick_fused512(bool power, int dtype, int mode);  // scalar-bank kernel
KernelFn pick_ws512(bool power, int mode);                // round-1 software-pipelined kernel
KernelFn pick_direct(bool power, int dtype);              // O(L K) fallback
KernelFn pick_bluestein(bool power, int dtype);           // non-power-of-two DFT sizes up to 512
KernelFn pick_tc_256(bool power, int dtype, int mode);
KernelFn pick_tc_512(bool power, int dtype, int mode);
KernelFn pick_tc_1024(bool power, int dtype, int mode);
KernelFn pick_tc_2048(bool power, int dtype, int mode);
KernelFn pick_tc2_512(bool power, int dtype, int mode);
KernelFn pick_tc2_probe(int which);                       // development probes (1: no bank, 2: no transform)
KernelFn pick_w512(bool power, int mode, int nt);
KernelFn pick_umma(bool power, int dtype, int nt);           // tcgen05 transform (stft_umma.cuh), mt = tiles of 16 filters in {2, 3, 4}

// frames per sub-group and pass: 2 where the registers allow it (R1 <= 16), see fft_frames
template <int N>
constexpr int tc_frames() {
  return FftGeom<N>::R1 <= 16 ? 2 : 1;
}

template <int N, int MODE>
KernelFn pick_tc_mode(bool power, int dtype) {
  constexpr int NF = tc_frames<N>();
  if (power) return dtype == PDS_I16 ? stft_tc_kernel<N, true, short, MODE, NF> : stft_tc_kernel<N, true, float, MODE, NF>;
  return dtype == PDS_I16 ? stft_tc_kernel<N, false, short, MODE, NF> : stft_tc_kernel<N, false, float, MODE, NF>;
}

template <int N>
KernelFn pick_tc_n(bool power, int dtype, int mode) {
  switch (mode) {
    case kRows13: return pick_tc_mode<N, kRows13>(power, dtype);
    case kRows16: return pick_tc_mode<N, kRows16>(power, dtype);
    default: return pick_tc_mode<N, kRowsAny>(power, dtype);
  }
}

template <int N, int MODE>
KernelFn pick_tc2_mode(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_tc2_kernel<N, true, short, MODE> : stft_tc2_kernel<N, true, float, MODE>;
  return dtype == PDS_I16 ? stft_tc2_kernel<N, false, short, MODE> : stft_tc2_kernel<N, false, float, MODE>;
}

template <int N>
KernelFn pick_tc2_n(bool power, int dtype, int mode) {
  switch (mode) {
    case kRows13: return pick_tc2_mode<N, kRows13>(power, dtype);
    case kRows16: return pick_tc2_mode<N, kRows16>(power, dtype);
    default: return pick_tc2_mode<N, kRowsAny>(power, dtype);
  }
}

}  // namespace pds
