// stft.cu -- host side of the fused STFT frame-feature path: plan construction (constant tables,
// kernel choice), tile tables, launches, and the C ABI.  The kernels live in stft_device.cuh and are
// instantiated by the stft_k_*.cu translation units.
#include <cuda_fp16.h>

#include "stft_device.cuh"
#include "stft_umma.cuh"

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
  // PDS_STFT_KERNEL, read once when the plan is made: unset or '2' = stft_tc2_kernel where it exists
  // (else stft_tc_kernel), '1' = stft_tc_kernel, 'p' = stft_w_kernel (warp-specialised pipeline,
  // 16-frame tiles), 's' = scalar-bank kernel, 'w' = the round-1 software-pipelined kernel
  int variant = 2;
  bool want_ws = false, want_scalar = false;
  bool blue = false;  // stft_bluestein_kernel: non-power-of-two dft_size <= 512
  size_t blue_smem_bytes = 0;
  int blue_grid_limit = 0;
  bool bf16_bank = false;  // dense bank contracted with bf16 two-term splits (fragment layout differs)
  bool um = false;  // stft_umma_kernel: the transform on tcgen05 (dft_size 512, frame shift a multiple of 4, <= 64 filters)
  size_t um_smem_bytes = 0;
  int um_nt = 0;
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

KernelFn pick_tc(const pds_stft_plan* plan, int dtype) {
  const bool tc2 = plan->variant == 2;
  switch (plan->N) {
    case 256: return pick_tc_256(plan->power, dtype, plan->row_mode);  // (eight lanes per frame: tc only)
    case 512:
      if (plan->variant == 3 || plan->variant == 4) {
        KernelFn fn = pick_tc2_probe(plan->variant - 2);
        if (fn) return fn;
      }
      return tc2 ? pick_tc2_512(plan->power, dtype, plan->row_mode) : pick_tc_512(plan->power, dtype, plan->row_mode);
    case 1024: return pick_tc_1024(plan->power, dtype, plan->row_mode);
    case 2048: return pick_tc_2048(plan->power, dtype, plan->row_mode);
    default: return nullptr;
  }
}

KernelFn pick_w(const pds_stft_plan* plan) { return pick_w512(plan->power, plan->row_mode, plan->w_nt); }

// round-to-nearest-even conversion to bfloat16 (the upper 16 bits of the result's float pattern)
uint32_t bf16_rn(float w) {
  uint32_t bits;
  std::memcpy(&bits, &w, 4);
  bits += 0x7fffu + ((bits >> 16) & 1u);
  return bits >> 16;
}

// round-to-nearest split of a weight into two tf32-representable terms
void split_tf32(float w, float* hi, float* lo) {
  uint32_t bits;
  std::memcpy(&bits, &w, 4);
  bits = (bits + 0x1000u) & 0xffffe000u;
  std::memcpy(hi, &bits, 4);
  *lo = w - *hi;
}

KernelFn pick_kernel(const pds_stft_plan* plan, int dtype) {
  if (plan->fast) {
    // the scalar-bank kernel is instantiated for N = 512 only; other sizes run stft_tc_kernel
    return plan->fused ? pick_fused512(plan->power, dtype, plan->row_mode) : nullptr;
  }
  return pick_direct(plan->power, dtype);
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
    if (force[0] == 'p') plan->variant = 5;
    if (force[0] == 'u') plan->variant = 6;
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
    // Dense banks (at least as many filter groups as warps): one item per group covers both halves of a
    // 32-frame tile, so that each weight fragment is fetched once per tile.
    // Fragment type: bf16 m16n8k16 MMAs on two-term splits by default (half the MMAs and half the fragment
    // bytes of split-tf32; relative error 3 * 2^-17 instead of 2^-20, tolerance 1e-4); PDS_STFT_BANK=tf32
    // keeps split-tf32 m16n8k8, which stft_w_kernel and dft_size 2048 always use.
    const bool dense = N <= 1024 && n_ntiles >= n_warps && getenv("PDS_STFT_SPLIT_HALVES") == nullptr;
    const char* bank_env = getenv("PDS_STFT_BANK");
    const bool dense_bf16 = N <= 1024 && plan->variant != 5 && !(bank_env && bank_env[0] == 't');
    plan->bf16_bank = dense_bf16;
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
      for (int b = blk0; b < blk0 + nblk; ++b) {
        if (dense_bf16) {
          // bf16 m16n8k16 fragments of the whole 16-bin block: k = 2t, 2t+1, 2t+8, 2t+9 <-> bins 4t .. 4t+3;
          // every weight as two bf16 terms (w1 = rn(w), w2 = rn(w - w1)): {b0_w1, b1_w1, b0_w2, b1_w2}
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            const int f = 8 * j + g;
            uint32_t packed[2][2] = {{0u, 0u}, {0u, 0u}};  // [term][b0 | b1]
            for (int c = 0; c < 4; ++c) {
              const int bin = 16 * b + 4 * t + c;
              float w = 0.f;
              if (f < F && bin >= d->band_lo[f] && bin < d->band_lo[f] + d->band_len[f])
                w = d->weights[d->band_off[f] + (bin - d->band_lo[f])];
              const uint32_t w1 = bf16_rn(w);
              float w1f;
              const uint32_t w1bits = w1 << 16;
              std::memcpy(&w1f, &w1bits, 4);
              const uint32_t w2 = bf16_rn(w - w1f);
              packed[0][c >> 1] |= w1 << (16 * (c & 1));
              packed[1][c >> 1] |= w2 << (16 * (c & 1));
            }
            for (int term = 0; term < 2; ++term)
              for (int q = 0; q < 2; ++q) {
                float as_float;
                std::memcpy(&as_float, &packed[term][q], 4);
                tc_frags.push_back(as_float);
              }
          }
          continue;
        }
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
      }
      // one item per 16-frame half -- or, when there are at least as many filter groups as warps, one
      // item for both halves of a 32-frame tile (each weight fragment is then fetched once per tile)
      if (dense) {
        runs.push_back({8 * j, dense_bf16 ? kBothHalvesBf16 : kBothHalves, blk0, 2 * nblk, frag});
      } else {
        for (int m0 = 0; m0 < (N > 1024 ? 16 : kTileFrames); m0 += 16)
          runs.push_back({8 * j, dense_bf16 ? (m0 | kHalfBf16) : m0, blk0, nblk, frag});
      }
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
        tc_items.push_back(r.slot == kBothHalves || r.slot == kBothHalvesBf16 ? r.nblk / 2 : r.nblk);
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
    if (plan->tc && plan->variant == 5 && N == 512 && d->preemph == 0.f && d->dither == 0.f && F <= 8 * kWMaxNT && !plan->bf16_bank) {
      const int w_span = (kWTile - 1) * S + L;
      const size_t w_bytes = WLayout::bytes(w_span, L, (int)(tc_frags.size() / 4));
      if (w_bytes <= smem_cap) {
        plan->w = true;
        plan->w_smem_bytes = w_bytes;
        plan->w_nt = (F + 7) / 8;
        p.span_max = w_span;
      }
    }
    // tcgen05 transform: dft_size 512, frame shift a multiple of four samples (16-byte frame starts), at most
    // 64 filters, operands + two sample stages in shared memory
    if (plan->tc && plan->variant == 6 && N == 512 && S % 4 == 0 && F <= 64) {
      const int kch = 2 * ((L + 63) / 64);
      const int nt = F <= 32 ? 2 : (F <= 48 ? 3 : 4);  // tiles of sixteen filters
      const size_t um_bytes = (size_t)um_layout(kch, p.span_max, L).total;
      if (kch <= kUmMaxKch && um_bytes <= smem_cap) {
        plan->um = true;
        plan->um_smem_bytes = um_bytes;
        plan->um_nt = nt;
        p.um_kch = kch;
      }
    }
    if (!plan->fused && !plan->tc) plan->fast = false;  // huge frame shift / dft_size 2048: direct kernel
    if (plan->fast && plan->fused && d->preemph == 0.f && d->dither == 0.f) {
      const WsLayout ws = ws_layout(N, G, R1, p.span_max, p_rows, npairs, plan->C, pair_total);
      plan->ws = (size_t)ws.total <= smem_cap;
      plan->ws_smem_bytes = ws.total;
    }
  }
  if (!plan->fast && !pow2 && N >= 16 && N <= kBlueM / 2 && getenv("PDS_STFT_NO_BLUESTEIN") == nullptr) {
    const size_t blue_bytes = (size_t)blue_layout((kBlueTileFrames - 1) * S + L, L, K).total;
    if (blue_bytes <= smem_cap) {
      plan->blue = true;
      plan->blue_smem_bytes = blue_bytes;
      p.span_max = (kBlueTileFrames - 1) * S + L;
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
  plan->tile_frames = plan->fast ? (N > 1024 || plan->w ? 16 : kTileFrames) : (plan->blue ? kBlueTileFrames : kDirectTileFrames);

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

  // Bluestein tables (double precision): a[n] = window[n] conj(c[n]), c[n] = exp(i pi n^2 / N) with n^2
  // reduced mod 2 N in integers; B = DFT_1024 of the chirp kernel b[m] = c[|m|], |m| < N, scaled by
  // 1 / 1024 (the inverse transform's factor); and the twiddles of the 1024-point transform
  std::vector<float2> blue_aw, blue_b, blue_tw;
  if (plan->blue) {
    const double pi = 3.14159265358979323846264338327950288;
    auto chirp = [&](long long n, double* re, double* im) {
      const long long q = (n * n) % (2LL * N);
      const double a = pi * (double)q / (double)N;
      *re = std::cos(a), *im = std::sin(a);
    };
    blue_aw.resize(L);
    for (int n = 0; n < L; ++n) {
      double cr, ci;
      chirp(n, &cr, &ci);
      blue_aw[n] = make_float2((float)(d->window[n] * cr), (float)(-(double)d->window[n] * ci));
    }
    std::vector<double> br(kBlueM, 0.0), bi(kBlueM, 0.0);
    for (int m = -(N - 1); m <= N - 1; ++m) {
      double cr, ci;
      chirp(m < 0 ? -m : m, &cr, &ci);
      br[(m + kBlueM) % kBlueM] = cr, bi[(m + kBlueM) % kBlueM] = ci;
    }
    blue_b.resize(kBlueM);
    for (int k = 0; k < kBlueM; ++k) {  // plain O(M^2) DFT, once per plan (1 M complex multiply-adds)
      double sr = 0.0, si = 0.0;
      for (int m = 0; m < kBlueM; ++m) {
        const double a = -two_pi * (double)((long long)k * m % kBlueM) / kBlueM;
        const double wr = std::cos(a), wi = std::sin(a);
        sr += br[m] * wr - bi[m] * wi;
        si += br[m] * wi + bi[m] * wr;
      }
      blue_b[k] = make_float2((float)(sr / kBlueM), (float)(si / kBlueM));
    }
    blue_tw.resize(kBlueM);
    for (int k1 = 0; k1 < 32; ++k1)
      for (int lane = 0; lane < 32; ++lane) {
        const double a = -two_pi * (double)(lane * k1) / kBlueM;
        blue_tw[k1 * 32 + lane] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
  }

  // ---- one device blob -----------------------------------------------------------------
  // stft_umma_kernel: constant operand, window, bank fragments in the kernel's bin order (see stft_umma.cuh)
  std::vector<unsigned char> um_bmat;
  std::vector<float> um_window;
  std::vector<uint32_t> um_frags;
  std::vector<int> um_masks(4 * kUmSteps, 0), um_offs(4 * kUmSteps, 0);
  if (plan->um) {
    const int kch = p.um_kch;
    um_window.assign(N, 0.f);
    for (int i = 0; i < L; ++i) um_window[i] = d->window[i];
    // row n of the operand = column n of D: Re Y[n] (n < 64), Re Y[64] (n = 64), Im Y[n - 64] (n > 64).
    // K position 8 c + k <-> j = 8 c + ((k + c) mod 8): the order in which lane c of a builder reads its samples
    um_bmat.assign(2 * (size_t)kch * kUmLboB, 0);
    for (int row = 0; row < 128; ++row) {
      for (int pos = 0; pos < 8 * kch; ++pos) {
        const int c = pos / 8, k = pos % 8, j = 8 * c + ((k + c) & 7);
        double v;
        if (row == 64) v = (j % 2) ? -1.0 : 1.0;  // cos(pi j)
        else {
          const int k1 = row % 64;
          const double a = two_pi * (double)((j * k1) % 128) / 128.0;
          v = row > 64 ? -std::sin(a) : std::cos(a);
        }
        const __half hi = __float2half_rn((float)v);
        const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
        const size_t at = (size_t)c * kUmLboB + (size_t)(row / 8) * 128 + (size_t)(row % 8) * 16 + (size_t)k * 2;
        const unsigned short hb = __half_as_ushort(hi), lb = __half_as_ushort(lo);
        std::memcpy(um_bmat.data() + at, &hb, 2);
        std::memcpy(um_bmat.data() + (size_t)kch * kUmLboB + at, &lb, 2);
      }
    }
    // bin of k-slot kk of bin group s at k1 = 16 qq + kk (-1: none)
    auto bin_of = [](int qq, int s, int kk) {
      const int k1 = 16 * qq + kk;
      switch (s) {
        case 0: return k1;
        case 1: return 128 + k1;
        case 2: return 256 - k1;
        case 3: return k1 == 0 ? 64 : 128 - k1;   // the duplicate of bin 128 carries bin 64
        default: return (qq == 0 && kk == 0) ? 192 : -1;
      }
    };
    auto weight = [&](int f, int bin) -> float {
      if (f >= F || bin < 0) return 0.f;
      const int tt = bin - d->band_lo[f];
      return (tt >= 0 && tt < d->band_len[f]) ? d->weights[d->band_off[f] + tt] : 0.f;
    };
    for (int qq = 0; qq < 4; ++qq)
      for (int s = 0; s < kUmSteps; ++s) {
        um_offs[qq * kUmSteps + s] = (int)(um_frags.size() / 256);
        if (s == 4 && qq != 0) continue;
        for (int mt = 0; mt < plan->um_nt; ++mt) {
          bool any = false;
          for (int kk = 0; kk < 16 && !any; ++kk)
            for (int gg = 0; gg < 16 && !any; ++gg) any = weight(16 * mt + gg, bin_of(qq, s, kk)) != 0.f;
          if (!any) continue;
          um_masks[qq * kUmSteps + s] |= 1 << mt;
          // mma.m16n8k16 A fragment of lane (g, t): a0 = (row g, k 2t, 2t+1), a1 = (row g + 8, same k),
          // a2 = (row g, k 2t + 8, 2t + 9), a3 = (row g + 8, same k); high terms of 32 lanes, then low terms
          std::vector<uint32_t> hi_part(128), lo_part(128);
          for (int ln = 0; ln < 32; ++ln) {
            const int gg = ln / 4, tt = ln % 4;
            for (int a = 0; a < 4; ++a) {
              const int row = 16 * mt + gg + ((a & 1) ? 8 : 0), k0 = 2 * tt + ((a & 2) ? 8 : 0);
              uint32_t hi[2], lo[2];
              for (int i = 0; i < 2; ++i) {
                const float w = weight(row, bin_of(qq, s, k0 + i));
                hi[i] = bf16_rn(w);
                const uint32_t hb = hi[i] << 16;
                float hf;
                std::memcpy(&hf, &hb, 4);
                lo[i] = bf16_rn(w - hf);
              }
              hi_part[4 * ln + a] = hi[0] | (hi[1] << 16);
              lo_part[4 * ln + a] = lo[0] | (lo[1] << 16);
            }
          }
          um_frags.insert(um_frags.end(), hi_part.begin(), hi_part.end());
          um_frags.insert(um_frags.end(), lo_part.begin(), lo_part.end());
        }
      }
  }
  if (um_frags.empty()) um_frags.resize(4, 0u);

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
  size_t o_baw = align16(o_tcf + sizeof(float) * tc_frags.size());
  size_t o_bb = align16(o_baw + sizeof(float2) * blue_aw.size());
  size_t o_btw = align16(o_bb + sizeof(float2) * blue_b.size());
  size_t o_uma = align16(o_btw + sizeof(float2) * blue_tw.size());
  size_t o_umw = align16(o_uma + um_bmat.size());
  size_t o_umf = align16(o_umw + sizeof(float) * um_window.size());
  size_t o_umm = align16(o_umf + sizeof(uint32_t) * um_frags.size());
  size_t o_umo = align16(o_umm + sizeof(int) * um_masks.size());
  size_t blob_bytes = align16(o_umo + sizeof(int) * um_offs.size());
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
  if (plan->blue) {
    std::memcpy(blob.data() + o_baw, blue_aw.data(), sizeof(float2) * blue_aw.size());
    std::memcpy(blob.data() + o_bb, blue_b.data(), sizeof(float2) * blue_b.size());
    std::memcpy(blob.data() + o_btw, blue_tw.data(), sizeof(float2) * blue_tw.size());
  }
  if (plan->um) {
    std::memcpy(blob.data() + o_uma, um_bmat.data(), um_bmat.size());
    std::memcpy(blob.data() + o_umw, um_window.data(), sizeof(float) * um_window.size());
  }
  std::memcpy(blob.data() + o_umf, um_frags.data(), sizeof(uint32_t) * um_frags.size());
  std::memcpy(blob.data() + o_umm, um_masks.data(), sizeof(int) * um_masks.size());
  std::memcpy(blob.data() + o_umo, um_offs.data(), sizeof(int) * um_offs.size());
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
  p.blue_aw = reinterpret_cast<const float2*>(base + o_baw);
  p.blue_b = reinterpret_cast<const float2*>(base + o_bb);
  p.blue_tw = reinterpret_cast<const float2*>(base + o_btw);
  p.tc_items = reinterpret_cast<const int4*>(base + o_tci);
  p.tc_wstart = reinterpret_cast<const int*>(base + o_tcw);
  p.tc_frags = reinterpret_cast<const float4*>(base + o_tcf);
  p.um_bmat = base + o_uma;
  p.um_window = reinterpret_cast<const float*>(base + o_umw);
  p.um_frags = base + o_umf;
  p.um_masks = reinterpret_cast<const int*>(base + o_umm);
  p.um_offs = reinterpret_cast<const int*>(base + o_umo);
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
  for (int dt = 0; dt < 2 && plan->um; ++dt) {
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_umma(plan->power, dt, plan->um_nt)),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->um_smem_bytes);
    if (err != cudaSuccess) {
      cudaGetLastError();
      plan->um = false;
    }
  }
  for (int dt = 0; dt < 2 && plan->blue; ++dt) {
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_bluestein(plan->power, dt)),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->blue_smem_bytes);
    if (err != cudaSuccess) {
      cudaGetLastError();
      plan->blue = false;
      plan->tile_frames = kDirectTileFrames;
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
    err = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_ws512(plan->power, plan->row_mode)),
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
  if (plan->blue) {
    int occ_b = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, reinterpret_cast<const void*>(pick_bluestein(plan->power, PDS_F32)),
                                                  kBlueThreads, plan->blue_smem_bytes);
    plan->blue_grid_limit = prop.multiProcessorCount * std::max(1, occ_b);
  }
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

extern "C" const char* pds_stft_kernel_name(const pds_stft_plan* plan, int sig_dtype) {
  if (!plan) return "";
  if (plan->ws && sig_dtype == PDS_F32 && plan->want_ws) return "pds::stft_ws_kernel";
  if (plan->um) return "pds::stft_umma_kernel";
  if (plan->w && sig_dtype == PDS_F32) return "pds::stft_w_kernel";
  if (plan->tc && !(plan->want_scalar && plan->fused)) {
    const bool tc2 = plan->variant == 2 && plan->N == 512;
    return tc2 ? "pds::stft_tc2_kernel" : "pds::stft_tc_kernel";
  }
  if (plan->blue) return "pds::stft_bluestein_kernel";
  return plan->fast ? "pds::stft_fused_kernel" : "pds::stft_direct_kernel";
}

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
    pick_ws512(plan->power, plan->row_mode)<<<grid, kWsThreads, plan->ws_smem_bytes,
                                                static_cast<cudaStream_t>(stream)>>>(p);
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  if (plan->um) {
    PDS_REQUIRE(n_tiles < ((int64_t)1 << 30), "at most 2^30 tiles per launch (got %lld)", (long long)n_tiles);
    const int grid = (int)std::min<int64_t>(n_tiles, plan->num_sms);
    pick_umma(plan->power, sig_dtype, plan->um_nt)<<<grid, kUmThreads, plan->um_smem_bytes,
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
  if (plan->blue) {
    const int grid = (int)std::min<int64_t>(n_tiles, plan->blue_grid_limit);
    pick_bluestein(plan->power, sig_dtype)<<<grid, kBlueThreads, plan->blue_smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
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
