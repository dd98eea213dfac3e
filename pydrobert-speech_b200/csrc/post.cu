// post.cu -- pre/post-processor kernels: Deltas (K4), CMVN statistics (K5) and application (K6),
// plus stand-alone pre-emphasis / dither passes for pipelines that cannot fuse them.
//
// All three post kernels are HBM-bound streaming passes over the packed (rows x cols) float32
// feature matrix; they are written for coalesced row-major access and enough resident CTAs to
// cover HBM latency (grid sized from the SM count), not for arithmetic throughput.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace pds {

// ------------------------------------------------------------------------------------------
// Deltas: out[r, k*cols + c] = sum_j f_k[j] * in[clamp(r + j - half_k, utt_lo, utt_hi - 1), c]
// (post.py:462-491 with pad_mode='edge', concatenate=True, axis=0, target_axis=-1)
// ------------------------------------------------------------------------------------------
constexpr int kDeltaMaxTaps = 240;
constexpr int kDeltaMaxOrders = 8;
constexpr int kDeltaMaxRows = 128;  // rows of output per CTA (fewer when the columns are wide)
constexpr int kDeltaMaxChunk = 256;  // columns per CTA
constexpr int kDeltaThreads = 256;

struct DeltaParams {
  const float* in;
  float* out;
  const long long* row_off;
  long long total_rows;
  long long n_utts;
  int cols;
  int orders;
  int half_max;
  int rows_per_cta;
  int chunk;  // columns per CTA
  int filt_off[kDeltaMaxOrders];
  int filt_len[kDeltaMaxOrders];
  float taps[kDeltaMaxTaps];
  // fused CMVN (deltas25_kernel modes 1 and 2): statistics of / normalisation by the OUTPUT columns
  double* stats;           // (2, 3*cols + 1): accumulated (mode 1) or read (mode 2)
  int norm_var;
  int* zero_var;
};

// One CTA produces `rows_per_cta` consecutive rows of one column chunk: the rows plus a halo of
// half_max rows on either side are staged in shared memory once (coalesced; each input row is read
// from HBM ~1.06 times), every thread then walks (row, col) pairs with col fastest -- no integer
// division in the loops -- and clamps the time index to its utterance's range, looked up once per row.
__global__ void __launch_bounds__(kDeltaThreads) deltas_kernel(const __grid_constant__ DeltaParams p) {
  extern __shared__ __align__(16) float s_rows[];  // (R + 2*half_max) x cw | lo[R] | hi[R]
  const int tid = threadIdx.x;
  const int R = p.rows_per_cta, H = p.half_max, cols = p.cols;
  const long long r0 = (long long)blockIdx.x * R;
  const int nrows = (int)min((long long)R, p.total_rows - r0);
  const int col0 = blockIdx.y * p.chunk;
  const int cw = min(p.chunk, cols - col0);
  const int stage_rows = nrows + 2 * H;
  int* s_lo = reinterpret_cast<int*>(s_rows + (R + 2 * H) * p.chunk);
  int* s_hi = s_lo + R;

  // utterance bounds of every output row, relative to r0 - H (the first staged row)
  if (tid < nrows) {
    const long long r = r0 + tid;
    long long lo = 0, hi = p.n_utts;  // largest u with row_off[u] <= r (skips empty utterances)
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if (p.row_off[mid] <= r) lo = mid; else hi = mid;
    }
    s_lo[tid] = (int)(p.row_off[lo] - (r0 - H));
    s_hi[tid] = (int)(p.row_off[lo + 1] - 1 - (r0 - H));
  }
  // stage the rows (clamped to the matrix; per-utterance clamping happens on use)
  const int dr = kDeltaThreads / cw, dc = kDeltaThreads - dr * cw;
  {
    const long long first = r0 - H;
    const int total = stage_rows * cw;
    if (cw == cols && first >= 0 && first + stage_rows <= p.total_rows) {
      const float* __restrict__ src = p.in + first * cols;  // one contiguous block
      int i = tid;
      for (; i + 3 * kDeltaThreads < total; i += 4 * kDeltaThreads) {
        const float a = src[i], b = src[i + kDeltaThreads], c = src[i + 2 * kDeltaThreads],
                    d = src[i + 3 * kDeltaThreads];
        s_rows[i] = a, s_rows[i + kDeltaThreads] = b, s_rows[i + 2 * kDeltaThreads] = c,
        s_rows[i + 3 * kDeltaThreads] = d;
      }
      for (; i < total; i += kDeltaThreads) s_rows[i] = src[i];
    } else {
      int sr = tid / cw, c = tid - sr * cw;
      for (int i = tid; i < total; i += kDeltaThreads) {
        const long long r = max(0LL, min(p.total_rows - 1, first + sr));
        s_rows[i] = p.in[r * cols + col0 + c];
        sr += dr, c += dc;
        if (c >= cw) c -= cw, ++sr;
      }
    }
  }
  __syncthreads();

  const int out_cols = cols * (p.orders + 1);
  int lr = tid / cw, c = tid - lr * cw;
  for (int i = tid; i < nrows * cw; i += kDeltaThreads) {
    const int lo = s_lo[lr], hi = s_hi[lr];
    const int centre = lr + H;  // staged index of this row
    float* __restrict__ dst = p.out + (r0 + lr) * out_cols + col0 + c;
    dst[0] = s_rows[centre * cw + c];
    for (int k = 0; k < p.orders; ++k) {
      const int len = p.filt_len[k], half = (len - 1) / 2;
      const float* __restrict__ f = p.taps + p.filt_off[k];
      float acc = 0.f;
      for (int j = 0; j < len; ++j) {
        const int rr = max(lo, min(hi, centre + j - half));
        acc = fmaf(f[j], s_rows[rr * cw + c], acc);
      }
      dst[(k + 1) * cols] = acc;
    }
    lr += dr, c += dc;
    if (c >= cw) c -= cw, ++lr;
  }
}

// Specialisation for the default Deltas(num_deltas=2, context_window=2): filters of 5 and 9 taps.
// Same staging; then every thread owns one column of eight consecutive rows and slides the 16
// staged values it needs through registers (taps in registers too): 16 shared-memory loads and
// 112 FMAs for 24 outputs instead of one clamped load per tap.  Groups that touch an utterance
// boundary take the clamped per-tap path.
//   MODE 0: store the (rows x 3*cols) outputs                     (post.Deltas)
//   MODE 1: do not store; add the outputs' per-column sum / sum of squares (float64) to `stats`
//           (post.Standardize.accumulate on Deltas output without materialising it)
//   MODE 2: store the outputs normalised with `stats`             (Deltas -> Standardize.apply)
// Persistent grid; a thread keeps its column for the whole launch (tid = group * cols + column).
constexpr int kD25Rows = 8;
enum { kD25Store = 0, kD25Stats = 1, kD25Apply = 2 };

template <int MODE>
__global__ void __launch_bounds__(kDeltaThreads) deltas25_kernel(const __grid_constant__ DeltaParams p) {
  extern __shared__ __align__(16) float s_rows[];  // (R + 8) x cols | lo[R] | hi[R] | scale, shift [3*cols]
  const int tid = threadIdx.x;
  const int R = p.rows_per_cta, cols = p.cols;
  constexpr int H = 4;
  const int out_cols = 3 * cols;
  int* s_lo = reinterpret_cast<int*>(s_rows + (R + 2 * H) * cols);
  int* s_hi = s_lo + R;
  float* s_scale = reinterpret_cast<float*>(s_hi + R);
  float* s_shift = s_scale + out_cols;
  float f1[5], f2[9];
#pragma unroll
  for (int j = 0; j < 5; ++j) f1[j] = p.taps[p.filt_off[0] + j];
#pragma unroll
  for (int j = 0; j < 9; ++j) f2[j] = p.taps[p.filt_off[1] + j];
  if (MODE == kD25Apply) {  // scale / shift of every output column (post.py:264-294)
    const double count = p.stats[out_cols];
    for (int c = tid; c < out_cols; c += kDeltaThreads) {
      const double mean = p.stats[c] / count;
      double scale = 1.0;
      if (p.norm_var) {
        double var = p.stats[out_cols + 1 + c] / count - mean * mean;
        if (fabs(var) <= 1e-8) {  // np.isclose(var, 0)
          var = 1.0;
          if (p.zero_var) *p.zero_var = 1;
        }
        scale = 1.0 / sqrt(var);
      }
      s_scale[c] = (float)scale;
      s_shift[c] = (float)(mean * scale);
    }
  }
  const int G = kDeltaThreads / cols;            // row groups in flight per pass
  const int g0 = tid / cols, c = tid - g0 * cols;  // this thread's column never changes
  const bool worker = g0 < G;
  double sum[3] = {0.0, 0.0, 0.0}, sq[3] = {0.0, 0.0, 0.0};
  const long long nblocks = (p.total_rows + R - 1) / R;

  for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const long long r0 = blk * R;
    const int nrows = (int)min((long long)R, p.total_rows - r0);
    const int stage_rows = nrows + 2 * H;
    __syncthreads();  // the previous block's rows are no longer needed
    if (tid < nrows) {
      const long long r = r0 + tid;
      long long lo = 0, hi = p.n_utts;  // largest u with row_off[u] <= r (skips empty utterances)
      while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (p.row_off[mid] <= r) lo = mid; else hi = mid;
      }
      s_lo[tid] = (int)(p.row_off[lo] - (r0 - H));
      s_hi[tid] = (int)(p.row_off[lo + 1] - 1 - (r0 - H));
    }
    {
      const long long first = r0 - H;
      const int total = stage_rows * cols;
      if (first >= 0 && first + stage_rows <= p.total_rows) {
        const float* __restrict__ src = p.in + first * cols;  // one contiguous block
        int i = tid;
        for (; i + 3 * kDeltaThreads < total; i += 4 * kDeltaThreads) {
          const float a = src[i], b = src[i + kDeltaThreads], cc = src[i + 2 * kDeltaThreads],
                      d = src[i + 3 * kDeltaThreads];
          s_rows[i] = a, s_rows[i + kDeltaThreads] = b, s_rows[i + 2 * kDeltaThreads] = cc,
          s_rows[i + 3 * kDeltaThreads] = d;
        }
        for (; i < total; i += kDeltaThreads) s_rows[i] = src[i];
      } else {
        for (int i = tid; i < total; i += kDeltaThreads) {
          const int sr = i / cols, cc = i - sr * cols;
          const long long r = max(0LL, min(p.total_rows - 1, first + sr));
          s_rows[i] = p.in[r * cols + cc];
        }
      }
    }
    __syncthreads();
    if (!worker) continue;
    const int groups = (nrows + kD25Rows - 1) / kD25Rows;
    for (int g = g0; g < groups; g += G) {
      const int row0 = g * kD25Rows;
      const int last = min(row0 + kD25Rows, nrows) - 1;
      float* __restrict__ dst = p.out + (r0 + row0) * out_cols + c;
      // the group is interior when the first row's window starts and the last row's window ends
      // inside their (common) utterance; staged index of local row lr is lr + H
      const bool interior = s_lo[row0] <= row0 && s_hi[last] >= row0 + kD25Rows + 2 * H - 1 &&
                            s_lo[last] == s_lo[row0];
      float v[kD25Rows + 2 * H];
      if (interior) {
#pragma unroll
        for (int i = 0; i < kD25Rows + 2 * H; ++i) v[i] = s_rows[(row0 + i) * cols + c];
      }
#pragma unroll
      for (int q = 0; q < kD25Rows; ++q) {
        if (row0 + q > last) break;
        float d0, d1 = 0.f, d2 = 0.f;
        if (interior) {
          d0 = v[q + H];
#pragma unroll
          for (int j = 0; j < 5; ++j) d1 = fmaf(f1[j], v[q + 2 + j], d1);
#pragma unroll
          for (int j = 0; j < 9; ++j) d2 = fmaf(f2[j], v[q + j], d2);
        } else {
          const int lr = row0 + q, lo = s_lo[lr], hi = s_hi[lr], centre = lr + H;
          d0 = s_rows[centre * cols + c];
#pragma unroll
          for (int j = 0; j < 5; ++j) d1 = fmaf(f1[j], s_rows[max(lo, min(hi, centre + j - 2)) * cols + c], d1);
#pragma unroll
          for (int j = 0; j < 9; ++j) d2 = fmaf(f2[j], s_rows[max(lo, min(hi, centre + j - 4)) * cols + c], d2);
        }
        if (MODE == kD25Stats) {
          sum[0] += (double)d0, sq[0] += (double)d0 * d0;
          sum[1] += (double)d1, sq[1] += (double)d1 * d1;
          sum[2] += (double)d2, sq[2] += (double)d2 * d2;
        } else if (MODE == kD25Apply) {
          dst[q * out_cols] = fmaf(d0, s_scale[c], -s_shift[c]);
          dst[q * out_cols + cols] = fmaf(d1, s_scale[cols + c], -s_shift[cols + c]);
          dst[q * out_cols + 2 * cols] = fmaf(d2, s_scale[2 * cols + c], -s_shift[2 * cols + c]);
        } else {
          dst[q * out_cols] = d0;
          dst[q * out_cols + cols] = d1;
          dst[q * out_cols + 2 * cols] = d2;
        }
      }
    }
  }
  if (MODE == kD25Stats) {
    // fold the row groups of a column through shared memory, then one atomic per column and CTA
    __syncthreads();
    double* s_red = reinterpret_cast<double*>(s_rows);  // [G][6][cols]
    if (worker) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s_red[(g0 * 6 + k) * cols + c] = sum[k];
        s_red[(g0 * 6 + 3 + k) * cols + c] = sq[k];
      }
    }
    __syncthreads();
    if (worker && g0 == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double a = 0.0, b = 0.0;
        for (int g = 0; g < G; ++g) a += s_red[(g * 6 + k) * cols + c], b += s_red[(g * 6 + 3 + k) * cols + c];
        atomicAdd(p.stats + k * cols + c, a);
        atomicAdd(p.stats + (out_cols + 1) + k * cols + c, b);
      }
    }
    if (blockIdx.x == 0 && tid == 0) atomicAdd(p.stats + out_cols, (double)p.total_rows);
  }
}

// ------------------------------------------------------------------------------------------
// Streaming form of the same specialisation (the default): no shared-memory staging.  A thread owns
// one column of a run of kD25sRun consecutive rows and slides the nine samples its filters need
// through registers: per group of eight rows it issues eight independent (coalesced: neighbouring
// threads are neighbouring columns) loads, then 8 x (5 + 9) FMAs and the stores -- about 25
// instructions per (row, column) against 75 in deltas25_kernel, whose load -> barrier -> compute
// structure also left it latency bound at four CTAs per SM (ncu: issue slots 47 % busy, long
// scoreboard 6.5 stalls per issue).  A CTA's 256 / cols runs are adjacent, so it reads and writes one
// contiguous chunk; the persistent grid walks the chunks.  The utterance holding a chunk's first row
// is looked up (binary search on row_off) by an otherwise idle thread while the previous chunk is in
// work; runs then walk forward from it.  Rows within four of an utterance boundary take clamped loads.
// ------------------------------------------------------------------------------------------
// rows per run, measured on 10.99 M x 41 (tools/probe_post.py): the statistics pass likes long runs (fewer
// halo reloads: 1.01 ms at 128 rows against 1.37 at 16), the storing passes short ones (a CTA's stores stay
// close together: 1.80 ms at 16 rows against 2.31 at 128)
constexpr int kD25sRunStats = 128, kD25sRunStore = 16;

template <int MODE>
__global__ void __launch_bounds__(kDeltaThreads, MODE == kD25Stats ? 4 : 3) deltas25s_kernel(const __grid_constant__ DeltaParams p) {
  __shared__ long long s_utt[2];
  extern __shared__ __align__(16) float s_dyn[];  // statistics fold (MODE 1)
  const int tid = threadIdx.x, cols = p.cols, out_cols = 3 * cols;
  const int runs = kDeltaThreads / cols;
  const int run = tid / cols, c = tid - run * cols;
  const bool worker = run < runs;
  const int run_rows = p.rows_per_cta;  // rows per run on this path
  const long long chunk_rows = (long long)runs * run_rows;
  const long long n_chunks = (p.total_rows + chunk_rows - 1) / chunk_rows;
  float f1[5], f2[9];
#pragma unroll
  for (int j = 0; j < 5; ++j) f1[j] = p.taps[p.filt_off[0] + j];
#pragma unroll
  for (int j = 0; j < 9; ++j) f2[j] = p.taps[p.filt_off[1] + j];
  float scale[3] = {1.f, 1.f, 1.f}, shift[3] = {0.f, 0.f, 0.f};
  if (MODE == kD25Apply && worker) {  // scale / shift of this thread's three output columns (post.py:264-294)
    const double count = p.stats[out_cols];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double mean = p.stats[k * cols + c] / count;
      double sc = 1.0;
      if (p.norm_var) {
        double var = p.stats[out_cols + 1 + k * cols + c] / count - mean * mean;
        if (fabs(var) <= 1e-8) {  // np.isclose(var, 0)
          var = 1.0;
          if (p.zero_var) *p.zero_var = 1;
        }
        sc = 1.0 / sqrt(var);
      }
      scale[k] = (float)sc;
      shift[k] = (float)(mean * sc);
    }
  }
  double sum[3] = {0.0, 0.0, 0.0}, sq[3] = {0.0, 0.0, 0.0};
  const float* __restrict__ in = p.in + c;

  // utterance of a chunk's first row: largest u with row_off[u] <= r (skips empty utterances)
  auto find_utt = [&](long long r) {
    long long lo = 0, hi = p.n_utts;
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if (p.row_off[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
  };
  const int searcher = kDeltaThreads - 1;
  if (tid == searcher && blockIdx.x < n_chunks) s_utt[0] = find_utt((long long)blockIdx.x * chunk_rows);
  __syncthreads();

  int parity = 0;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, parity ^= 1) {
    const long long next = chunk + gridDim.x;
    if (tid == searcher && next < n_chunks) s_utt[parity ^ 1] = find_utt(next * chunk_rows);
    if (worker) {
      const long long r_begin = chunk * chunk_rows + (long long)run * run_rows;
      const long long r_end = min(r_begin + run_rows, p.total_rows);
      long long u = s_utt[parity];
      long long lo = 0, hi = -1;  // rows of the current utterance (inclusive)
      float w[16];                // w[j] = x[r + j - 4] for the group's first row r
      float nxt[8];               // rows r + 12 .. r + 19, requested while this group computes
      bool carried = false, have_nxt = false;
      for (long long r = r_begin; r < r_end; r += 8) {
        if (r > hi) {  // (re)locate the utterance of row r
          while (p.row_off[u + 1] <= r) ++u;
          lo = p.row_off[u], hi = p.row_off[u + 1] - 1;
        }
        const int n = (int)min(8LL, r_end - r);
        const bool interior = r - 4 >= lo && r + 11 <= hi && n == 8;
        if (interior) {
          const float* __restrict__ src = in + (r - 4) * cols;
          if (carried) {  // the previous group left rows r - 4 .. r + 3 in w[8..15]
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = w[j + 8];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = src[(long long)j * cols];
          }
          if (carried && have_nxt) {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[8 + j] = nxt[j];
          } else {
#pragma unroll
            for (int j = 8; j < 16; ++j) w[j] = src[(long long)j * cols];
          }
          // software pipeline (storing passes: three CTAs per SM at 80 registers; the statistics pass is issue
          // bound and keeps four): the next group's new rows are requested before this group's arithmetic
          have_nxt = MODE != kD25Stats && r + 4 >= lo && r + 19 <= hi && r_end - (r + 8) >= 8;
          if (have_nxt) {
#pragma unroll
            for (int j = 0; j < 8; ++j) nxt[j] = src[(long long)(16 + j) * cols];
          }
        }
        carried = interior;
        // one row's three values: into the statistics, or out (normalised or not)
        auto emit = [&](long long row, float d0, float d1, float d2) {
          if (MODE == kD25Stats) {
            sum[0] += (double)d0, sq[0] += (double)d0 * d0;
            sum[1] += (double)d1, sq[1] += (double)d1 * d1;
            sum[2] += (double)d2, sq[2] += (double)d2 * d2;
          } else {
            float* __restrict__ dst = p.out + row * out_cols + c;
            if (MODE == kD25Apply) {
              __stcs(dst, fmaf(d0, scale[0], -shift[0]));
              __stcs(dst + cols, fmaf(d1, scale[1], -shift[1]));
              __stcs(dst + 2 * cols, fmaf(d2, scale[2], -shift[2]));
            } else {
              __stcs(dst, d0);
              __stcs(dst + cols, d1);
              __stcs(dst + 2 * cols, d2);
            }
          }
        };
        if (interior) {  // eight rows straight from the window: no per-row branches
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) d1 = fmaf(f1[j], w[q + 2 + j], d1);
#pragma unroll
            for (int j = 0; j < 9; ++j) d2 = fmaf(f2[j], w[q + j], d2);
            emit(r + q, w[q + 4], d1, d2);
          }
        } else {  // within four rows of an utterance boundary, or the tail of the run: clamped loads
          for (int q = 0; q < n; ++q) {
            const long long rr = r + q;
            if (rr > hi) {
              while (p.row_off[u + 1] <= rr) ++u;
              lo = p.row_off[u], hi = p.row_off[u + 1] - 1;
            }
            float d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) d1 = fmaf(f1[j], in[max(lo, min(hi, rr + j - 2)) * cols], d1);
#pragma unroll
            for (int j = 0; j < 9; ++j) d2 = fmaf(f2[j], in[max(lo, min(hi, rr + j - 4)) * cols], d2);
            emit(rr, in[rr * cols], d1, d2);
          }
        }
      }
    }
    __syncthreads();  // the next chunk's utterance index is in place
  }
  if (MODE == kD25Stats) {
    // fold the runs of a column through shared memory, then one atomic per column and CTA
    double* s_red = reinterpret_cast<double*>(s_dyn);  // [runs][6][cols]
    if (worker) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s_red[(run * 6 + k) * cols + c] = sum[k];
        s_red[(run * 6 + 3 + k) * cols + c] = sq[k];
      }
    }
    __syncthreads();
    if (worker && run == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double a = 0.0, b = 0.0;
        for (int g = 0; g < runs; ++g) a += s_red[(g * 6 + k) * cols + c], b += s_red[(g * 6 + 3 + k) * cols + c];
        atomicAdd(p.stats + k * cols + c, a);
        atomicAdd(p.stats + (out_cols + 1) + k * cols + c, b);
      }
    }
    if (blockIdx.x == 0 && tid == 0) atomicAdd(p.stats + out_cols, (double)p.total_rows);
  }
}

// ------------------------------------------------------------------------------------------
// CMVN statistics: per-column sum and sum of squares in float64 (post.py:175-191)
// ------------------------------------------------------------------------------------------
constexpr int kStatsThreadsX = 32;
constexpr int kStatsThreadsY = 8;

__global__ void __launch_bounds__(kStatsThreadsX* kStatsThreadsY)
    cmvn_stats_kernel(const float* __restrict__ feats, long long n_rows, int cols,
                      double* __restrict__ stats) {
  // blockIdx.y picks a 32-column stripe, blockIdx.x a slab of rows; thread.x = column (coalesced),
  // thread.y strides rows.
  __shared__ double s_sum[kStatsThreadsY][kStatsThreadsX];
  __shared__ double s_sq[kStatsThreadsY][kStatsThreadsX];
  const int c = blockIdx.y * kStatsThreadsX + threadIdx.x;
  double sum = 0.0, sq = 0.0;
  if (c < cols) {
    const long long stride = (long long)gridDim.x * kStatsThreadsY;
    long long r = (long long)blockIdx.x * kStatsThreadsY + threadIdx.y;
    // four independent loads in flight per thread
    for (; r + 3 * stride < n_rows; r += 4 * stride) {
      const float a = feats[r * cols + c], b = feats[(r + stride) * cols + c];
      const float d = feats[(r + 2 * stride) * cols + c], e = feats[(r + 3 * stride) * cols + c];
      sum += (double)a + (double)b + (double)d + (double)e;
      sq += (double)a * a + (double)b * b + (double)d * d + (double)e * e;
    }
    for (; r < n_rows; r += stride) {
      const float a = feats[r * cols + c];
      sum += (double)a;
      sq += (double)a * a;
    }
  }
  s_sum[threadIdx.y][threadIdx.x] = sum;
  s_sq[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
#pragma unroll
    for (int y = 1; y < kStatsThreadsY; ++y) {
      sum += s_sum[y][threadIdx.x];
      sq += s_sq[y][threadIdx.x];
    }
    atomicAdd(stats + c, sum);
    atomicAdd(stats + (cols + 1) + c, sq);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0)
    atomicAdd(stats + cols, (double)n_rows);
}

// 16-byte variant: the matrix is read as a flat float4 stream.  With B threads per CTA such that
// 4 * B is a multiple of `cols`, and CTA slabs that are multiples of B float4s, the four
// (consecutive, wrapping) columns a thread sees never change: sums stay in registers (float64),
// are folded through shared memory once per CTA, then one atomic per column and CTA.
__global__ void __launch_bounds__(256)
    cmvn_stats4_kernel(const float4* __restrict__ feats4, long long total4, long long slab, int cols, int B,
                       long long n_rows, double* __restrict__ stats) {
  extern __shared__ __align__(16) double s_fold[];  // sum[cols] | sq[cols]
  const int t = threadIdx.x;
  for (int i = t; i < 2 * cols; i += blockDim.x) s_fold[i] = 0.0;
  __syncthreads();
  if (t < B) {
    const long long begin = (long long)blockIdx.x * slab, end = min(total4, begin + slab);
    double sum[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
    long long i = begin + t;
    for (; i + 3LL * B < end; i += 4LL * B) {  // four independent 16-byte loads in flight
      const float4 a = __ldcs(feats4 + i), b = __ldcs(feats4 + i + B), c = __ldcs(feats4 + i + 2LL * B),
                   d = __ldcs(feats4 + i + 3LL * B);
      sum[0] += ((double)a.x + (double)b.x) + ((double)c.x + (double)d.x);
      sum[1] += ((double)a.y + (double)b.y) + ((double)c.y + (double)d.y);
      sum[2] += ((double)a.z + (double)b.z) + ((double)c.z + (double)d.z);
      sum[3] += ((double)a.w + (double)b.w) + ((double)c.w + (double)d.w);
      sq[0] += ((double)a.x * a.x + (double)b.x * b.x) + ((double)c.x * c.x + (double)d.x * d.x);
      sq[1] += ((double)a.y * a.y + (double)b.y * b.y) + ((double)c.y * c.y + (double)d.y * d.y);
      sq[2] += ((double)a.z * a.z + (double)b.z * b.z) + ((double)c.z * c.z + (double)d.z * d.z);
      sq[3] += ((double)a.w * a.w + (double)b.w * b.w) + ((double)c.w * c.w + (double)d.w * d.w);
    }
    for (; i < end; i += B) {
      const float4 a = __ldcs(feats4 + i);
      sum[0] += (double)a.x, sum[1] += (double)a.y, sum[2] += (double)a.z, sum[3] += (double)a.w;
      sq[0] += (double)a.x * a.x, sq[1] += (double)a.y * a.y, sq[2] += (double)a.z * a.z, sq[3] += (double)a.w * a.w;
    }
    int c = (int)((4LL * t) % cols);  // 4 * begin is a multiple of cols
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_fold[c], sum[k]);
      atomicAdd(&s_fold[cols + c], sq[k]);
      c = c + 1 == cols ? 0 : c + 1;
    }
  }
  __syncthreads();
  for (int c = t; c < cols; c += blockDim.x) {
    atomicAdd(stats + c, s_fold[c]);
    atomicAdd(stats + (cols + 1) + c, s_fold[cols + c]);
  }
  if (blockIdx.x == 0 && t == 0) atomicAdd(stats + cols, (double)n_rows);
}

// ------------------------------------------------------------------------------------------
// CMVN apply: y = x * scale[c] - mean[c] * scale[c]  (post.py:264-294)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    cmvn_apply_kernel(const float* __restrict__ feats, float* __restrict__ out, long long n_rows,
                      int cols, const double* __restrict__ stats, int norm_var,
                      int* __restrict__ zero_var) {
  extern __shared__ __align__(16) float s_tab[];  // scale[cols] | shift[cols]
  float* s_scale = s_tab;
  float* s_shift = s_tab + cols;
  const double count = stats[cols];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const double mean = stats[c] / count;
    double scale = 1.0;
    if (norm_var) {
      double var = stats[cols + 1 + c] / count - mean * mean;
      if (fabs(var) <= 1e-8) {  // np.isclose(var, 0)
        var = 1.0;
        if (zero_var) *zero_var = 1;
      }
      scale = 1.0 / sqrt(var);
    }
    s_scale[c] = (float)scale;
    s_shift[c] = (float)(mean * scale);
  }
  __syncthreads();
  const long long total = n_rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c = (int)(i % cols);
  const int step = (int)(stride % cols);
  // 16 bytes per thread and access when the buffers allow it: a quarter of the memory requests
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const long long total4 = total >> 2;
    const int step4 = (int)((4 * stride) % cols);
    int c0 = (int)((4 * i) % cols);
    for (long long i4 = i; i4 < total4; i4 += stride) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(feats) + i4);
      int c1 = c0 + 1 == cols ? 0 : c0 + 1;
      int c2 = c1 + 1 == cols ? 0 : c1 + 1;
      int c3 = c2 + 1 == cols ? 0 : c2 + 1;
      float4 y;
      y.x = fmaf(x.x, s_scale[c0], -s_shift[c0]);
      y.y = fmaf(x.y, s_scale[c1], -s_shift[c1]);
      y.z = fmaf(x.z, s_scale[c2], -s_shift[c2]);
      y.w = fmaf(x.w, s_scale[c3], -s_shift[c3]);
      __stcs(reinterpret_cast<float4*>(out) + i4, y);
      c0 += step4;
      if (c0 >= cols) c0 -= cols;
    }
    return;
  }
  for (; i < total; i += stride) {
    out[i] = fmaf(feats[i], s_scale[c], -s_shift[c]);
    c += step;
    if (c >= cols) c -= cols;
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone pre-processors over a packed batch; blockIdx.y = utterance
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    preemph_kernel(const float* __restrict__ in, float* __restrict__ out,
                   const long long* __restrict__ sig_off, const long long* __restrict__ sig_len,
                   float coeff) {
  const long long off = sig_off[blockIdx.y], len = sig_len[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const float x = in[off + i];
    out[off + i] = i > 0 ? fmaf(-coeff, in[off + i - 1], x) : x;
  }
}

__global__ void __launch_bounds__(256)
    dither_kernel(const float* __restrict__ in, float* __restrict__ out,
                  const long long* __restrict__ sig_off, const long long* __restrict__ sig_len,
                  float coeff, unsigned long long seed) {
  const long long off = sig_off[blockIdx.y], len = sig_len[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
    out[off + i] = in[off + i] + coeff * philox_normal(seed, blockIdx.y, (unsigned long long)i);
}

}  // namespace pds

using namespace pds;

namespace {
// shared by pds_deltas / pds_deltas_cmvn_accumulate / pds_deltas_cmvn_apply
int run_deltas(int mode, const float* d_in, float* d_out, int64_t total_rows, int32_t n_cols, int64_t n_utts,
               const int64_t* d_row_off, int32_t orders, const float* h_filters, const int32_t* h_filter_len,
               double* d_stats, int32_t norm_var, int32_t* d_zero_var, void* stream) {
  PDS_REQUIRE(total_rows >= 0 && n_cols >= 1 && n_utts >= 0 && orders >= 0, "bad shape");
  if (total_rows == 0) return PDS_OK;
  PDS_REQUIRE(d_in && (d_out || mode == kD25Stats) && d_row_off && n_utts >= 1, "null buffer");
  PDS_REQUIRE(orders <= kDeltaMaxOrders, "at most %d delta orders are supported", kDeltaMaxOrders);
  PDS_REQUIRE(orders == 0 || (h_filters && h_filter_len), "null filter table");
  DeltaParams p;
  p.in = d_in, p.out = d_out, p.row_off = reinterpret_cast<const long long*>(d_row_off);
  p.total_rows = total_rows, p.n_utts = n_utts, p.cols = n_cols, p.orders = orders, p.half_max = 0;
  p.stats = d_stats, p.norm_var = norm_var, p.zero_var = d_zero_var;
  int at = 0;
  for (int k = 0; k < orders; ++k) {
    const int len = h_filter_len[k];
    PDS_REQUIRE(len >= 1 && (len % 2) == 1, "delta filter %d has even/empty length %d", k, len);
    if (at + len > kDeltaMaxTaps) {
      set_error("delta filters need %d taps (max %d)", at + len, kDeltaMaxTaps);
      return PDS_ERR_UNSUPPORTED;
    }
    p.filt_off[k] = at, p.filt_len[k] = len;
    for (int j = 0; j < len; ++j) p.taps[at + j] = h_filters[at + j];
    at += len;
    p.half_max = std::max(p.half_max, (len - 1) / 2);
  }
  // column chunks of at most 256, and as many rows per CTA as ~64 KB of shared memory hold
  p.chunk = std::min<int>(n_cols, kDeltaMaxChunk);
  const int budget_rows = (64 * 1024) / (int)(sizeof(float) * p.chunk) - 2 * p.half_max;
  p.rows_per_cta = std::max(8, std::min(kDeltaMaxRows, budget_rows));
  size_t smem = sizeof(float) * (size_t)(p.rows_per_cta + 2 * p.half_max) * p.chunk +
                2 * sizeof(int) * p.rows_per_cta;
  if (smem > 200 * 1024) {
    set_error("deltas: a context of %d rows exceeds shared memory", p.half_max);
    return PDS_ERR_UNSUPPORTED;
  }
  const long long grid_x = (total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  const bool fast = orders == 2 && p.filt_len[0] == 5 && p.filt_len[1] == 9 && n_cols <= kDeltaMaxChunk;
  if (fast) {
    // default Deltas(2, context 2): register-sliding specialisation, persistent grid
    int device = 0;
    PDS_CUDA_CHECK(cudaGetDevice(&device));
    smem += 2 * sizeof(float) * 3 * (size_t)n_cols;                                     // scale | shift
    smem = std::max(smem, sizeof(double) * 6 * (size_t)(kDeltaThreads / n_cols) * n_cols + 16);  // statistics fold
    const void* fn = mode == kD25Stats   ? reinterpret_cast<const void*>(deltas25_kernel<kD25Stats>)
                     : mode == kD25Apply ? reinterpret_cast<const void*>(deltas25_kernel<kD25Apply>)
                                         : reinterpret_cast<const void*>(deltas25_kernel<kD25Store>);
    if (smem > 48 * 1024) PDS_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the statistics mode keeps its sums in registers across row blocks (persistent grid, one
    // atomic per column and CTA); the store modes run one CTA per row block, which measured faster
    const unsigned grid = mode == kD25Stats ? (unsigned)std::min<long long>(grid_x, (long long)sm_count(device) * 8)
                                            : (unsigned)grid_x;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool staged = getenv("PDS_DELTAS_KERNEL") && getenv("PDS_DELTAS_KERNEL")[0] == 's';
    if (!staged) {
      // streaming kernel (default): persistent grid, eight CTAs per SM; dynamic shared memory only
      // for the statistics fold
      static const int run_env = getenv("PDS_D25_RUN") ? std::max(8, atoi(getenv("PDS_D25_RUN")) / 8 * 8) : 0;
      static const int grid_env = getenv("PDS_D25_GRID") ? std::max(1, atoi(getenv("PDS_D25_GRID"))) : 0;
      const int run_rows = run_env ? run_env : (mode == kD25Stats ? kD25sRunStats : kD25sRunStore);
      const int ctas_per_sm = grid_env ? grid_env : 24;  // whole waves of the 4 (statistics) or 3 (storing passes) resident CTAs
      p.rows_per_cta = run_rows;
      const long long chunk_rows = (long long)(kDeltaThreads / n_cols) * run_rows;
      const long long n_chunks = (total_rows + chunk_rows - 1) / chunk_rows;
      const unsigned sgrid = (unsigned)std::min<long long>(n_chunks, (long long)sm_count(device) * ctas_per_sm);
      const size_t fold = mode == kD25Stats ? sizeof(double) * 6 * (size_t)(kDeltaThreads / n_cols) * n_cols + 16 : 0;
      if (mode == kD25Stats) deltas25s_kernel<kD25Stats><<<sgrid, kDeltaThreads, fold, st>>>(p);
      else if (mode == kD25Apply) deltas25s_kernel<kD25Apply><<<sgrid, kDeltaThreads, fold, st>>>(p);
      else deltas25s_kernel<kD25Store><<<sgrid, kDeltaThreads, fold, st>>>(p);
      PDS_CUDA_CHECK(cudaGetLastError());
      return PDS_OK;
    }
    if (mode == kD25Stats) deltas25_kernel<kD25Stats><<<grid, kDeltaThreads, smem, st>>>(p);
    else if (mode == kD25Apply) deltas25_kernel<kD25Apply><<<grid, kDeltaThreads, smem, st>>>(p);
    else deltas25_kernel<kD25Store><<<grid, kDeltaThreads, smem, st>>>(p);
    PDS_CUDA_CHECK(cudaGetLastError());
    return PDS_OK;
  }
  if (mode != kD25Store) {
    set_error("fused Deltas + CMVN covers Deltas(num_deltas=2, context_window=2) only; materialise the deltas first");
    return PDS_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    PDS_CUDA_CHECK(cudaFuncSetAttribute(deltas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid_y = (n_cols + p.chunk - 1) / p.chunk;
  PDS_REQUIRE(grid_y <= 65535, "too many columns (%d)", n_cols);
  deltas_kernel<<<dim3((unsigned)grid_x, (unsigned)grid_y), kDeltaThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}
}  // namespace

extern "C" int pds_deltas(const float* d_in, float* d_out, int64_t total_rows, int32_t n_cols,
                          int64_t n_utts, const int64_t* d_row_off, int32_t orders,
                          const float* h_filters, const int32_t* h_filter_len, void* stream) {
  return run_deltas(kD25Store, d_in, d_out, total_rows, n_cols, n_utts, d_row_off, orders, h_filters,
                    h_filter_len, nullptr, 0, nullptr, stream);
}

extern "C" int pds_deltas_cmvn_accumulate(const float* d_in, int64_t total_rows, int32_t n_cols, int64_t n_utts,
                                          const int64_t* d_row_off, int32_t orders, const float* h_filters,
                                          const int32_t* h_filter_len, double* d_stats, void* stream) {
  PDS_REQUIRE(d_stats, "null statistics buffer");
  return run_deltas(kD25Stats, d_in, nullptr, total_rows, n_cols, n_utts, d_row_off, orders, h_filters,
                    h_filter_len, d_stats, 0, nullptr, stream);
}

extern "C" int pds_deltas_cmvn_apply(const float* d_in, float* d_out, int64_t total_rows, int32_t n_cols,
                                     int64_t n_utts, const int64_t* d_row_off, int32_t orders,
                                     const float* h_filters, const int32_t* h_filter_len, const double* d_stats,
                                     int32_t norm_var, int32_t* d_zero_var, void* stream) {
  PDS_REQUIRE(d_stats, "null statistics buffer");
  return run_deltas(kD25Apply, d_in, d_out, total_rows, n_cols, n_utts, d_row_off, orders, h_filters,
                    h_filter_len, const_cast<double*>(d_stats), norm_var, d_zero_var, stream);
}

extern "C" int pds_cmvn_accumulate(const float* d_feats, int64_t n_rows, int32_t n_cols,
                                   double* d_stats, void* stream) {
  PDS_REQUIRE(n_rows >= 0 && n_cols >= 1, "bad shape");
  if (n_rows == 0) return PDS_OK;
  PDS_REQUIRE(d_feats && d_stats, "null buffer");
  int device = 0;
  PDS_CUDA_CHECK(cudaGetDevice(&device));
  {
    // 16-byte path: needs a thread count B <= 256 with cols | 4B, an aligned buffer, whole float4s
    int g = n_cols % 4 == 0 ? 4 : (n_cols % 2 == 0 ? 2 : 1);
    const int unit = n_cols / g;  // B must be a multiple of this
    const long long total = n_rows * (long long)n_cols;
    if (unit <= 256 && (total & 3) == 0 && (reinterpret_cast<uintptr_t>(d_feats) & 15u) == 0) {
      const int B = unit * (256 / unit);
      const long long total4 = total >> 2;
      long long ctas = std::min<long long>((long long)sm_count(device) * 8, (total4 + B - 1) / B);
      long long slab = (total4 + ctas - 1) / ctas;
      slab = (slab + B - 1) / B * B;
      ctas = (total4 + slab - 1) / slab;
      cmvn_stats4_kernel<<<(unsigned)ctas, 256, 2 * sizeof(double) * n_cols, static_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const float4*>(d_feats), total4, slab, n_cols, B, n_rows, d_stats);
      PDS_CUDA_CHECK(cudaGetLastError());
      return PDS_OK;
    }
  }
  const int stripes = (n_cols + kStatsThreadsX - 1) / kStatsThreadsX;
  // ~8 resident CTAs per SM in total, but never more slabs than there are row groups
  long long slabs = std::max(1, sm_count(device) * 8 / stripes);
  slabs = std::min<long long>(slabs, (n_rows + kStatsThreadsY - 1) / kStatsThreadsY);
  dim3 grid((unsigned)slabs, (unsigned)stripes), block(kStatsThreadsX, kStatsThreadsY);
  cmvn_stats_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_feats, n_rows, n_cols, d_stats);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}

extern "C" int pds_cmvn_apply(const float* d_feats, float* d_out, int64_t n_rows, int32_t n_cols,
                              const double* d_stats, int32_t norm_var, int32_t* d_zero_var,
                              void* stream) {
  PDS_REQUIRE(n_rows >= 0 && n_cols >= 1, "bad shape");
  if (n_rows == 0) return PDS_OK;
  PDS_REQUIRE(d_feats && d_out && d_stats, "null buffer");
  int device = 0;
  PDS_CUDA_CHECK(cudaGetDevice(&device));
  const long long total = n_rows * (long long)n_cols;
  long long grid = std::min<long long>((total + 255) / 256, (long long)sm_count(device) * 8);
  cmvn_apply_kernel<<<(unsigned)grid, 256, 2 * sizeof(float) * n_cols, static_cast<cudaStream_t>(stream)>>>(
      d_feats, d_out, n_rows, n_cols, d_stats, norm_var, d_zero_var);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}

extern "C" int pds_preemphasize(const float* d_in, float* d_out, int64_t n_utts,
                                const int64_t* d_sig_off, const int64_t* d_sig_len,
                                int64_t total_samples, float coeff, void* stream) {
  PDS_REQUIRE(n_utts >= 0 && total_samples >= 0, "bad shape");
  if (n_utts == 0 || total_samples == 0) return PDS_OK;
  PDS_REQUIRE(d_in && d_out && d_in != d_out && d_sig_off && d_sig_len, "null or aliased buffer");
  PDS_REQUIRE(n_utts <= 65535, "at most 65535 utterances per call");
  const long long per_utt = std::max<long long>(1, total_samples / n_utts);
  dim3 grid((unsigned)std::min<long long>(64, (per_utt + 1023) / 1024), (unsigned)n_utts);
  preemph_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, d_out, reinterpret_cast<const long long*>(d_sig_off),
      reinterpret_cast<const long long*>(d_sig_len), coeff);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}

extern "C" int pds_dither(const float* d_in, float* d_out, int64_t n_utts, const int64_t* d_sig_off,
                          const int64_t* d_sig_len, int64_t total_samples, float coeff,
                          uint64_t seed, void* stream) {
  PDS_REQUIRE(n_utts >= 0 && total_samples >= 0, "bad shape");
  if (n_utts == 0 || total_samples == 0) return PDS_OK;
  PDS_REQUIRE(d_in && d_out && d_sig_off && d_sig_len, "null buffer");
  PDS_REQUIRE(n_utts <= 65535, "at most 65535 utterances per call");
  const long long per_utt = std::max<long long>(1, total_samples / n_utts);
  dim3 grid((unsigned)std::min<long long>(64, (per_utt + 1023) / 1024), (unsigned)n_utts);
  dither_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, d_out, reinterpret_cast<const long long*>(d_sig_off),
      reinterpret_cast<const long long*>(d_sig_len), coeff, seed);
  PDS_CUDA_CHECK(cudaGetLastError());
  return PDS_OK;
}
