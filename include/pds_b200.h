/* pds_b200.h -- C ABI of the B200-native frame-feature hot path of pydrobert-speech.
 *
 * The reference (sdrobert/pydrobert-speech) is pure Python/NumPy and has NO FFI of its own; its
 * boundary is the Python class surface.  This header is the boundary a
 * non-Python host would bind, and what the Python shells in pydrobert-speech_b200/ call through
 * ctypes.  Each entry point names the reference routine it replaces
 * (paths relative to /root/reference/src/pydrobert/speech/).
 *
 * Conventions
 *   - every function returns 0 (PDS_OK) or a negative pds_status; no exceptions cross the ABI;
 *     pds_last_error() returns a thread-local, human readable description of the last failure.
 *   - "d_" pointers are device pointers on the plan's device, "h_" pointers are host pointers.
 *   - no hidden synchronisation in the *_run/*device entry points: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) and the caller synchronises.
 *   - the caller owns every buffer; a plan owns only its device-side constant tables.
 *   - there is no CPU fallback: without a usable CUDA device plan creation fails.
 */
#ifndef PDS_B200_H_
#define PDS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pds_status {
  PDS_OK = 0,
  PDS_ERR_INVALID = -1,     /* bad argument (maps to ValueError in the Python shell)   */
  PDS_ERR_CUDA = -2,        /* a CUDA runtime call failed; see pds_last_error()         */
  PDS_ERR_UNSUPPORTED = -3, /* geometry outside what the kernels implement             */
  PDS_ERR_NOMEM = -4
} pds_status;

typedef enum pds_dtype {
  PDS_F32 = 0, /* IEEE float32 samples                                   */
  PDS_I16 = 1, /* signed 16-bit PCM samples (converted to float, unscaled) */
  PDS_F64 = 2  /* reserved: the kernels take float32 / int16; hosts convert float64 first */
} pds_dtype;

const char* pds_last_error(void);
int pds_version(void);
/* Number of CUDA devices visible (0 when there is no driver / device). */
int pds_device_count(void);

/* ------------------------------------------------------------------------------------------ *
 * STFT frame computer                                                                          *
 * replaces compute.py:388-460 (_compute_frame) and :574-607 (compute_full), with               *
 * pre.py:90-149 (Dither, Preemphasize) optionally fused into the sample-staging step.          *
 * ------------------------------------------------------------------------------------------ */

typedef struct pds_stft_desc {
  int32_t frame_length;   /* L: samples per frame                      (compute.py:319-332) */
  int32_t frame_shift;    /* S: samples between frames                 (compute.py:305)     */
  int32_t dft_size;       /* N >= L: DFT length                        (compute.py:344-347) */
  int32_t pad_left;       /* symmetric padding before frame 0          (compute.py:582-587) */
  int32_t num_filts;      /* F                                                               */
  int32_t include_energy; /* 1: column 0 holds the frame energy        (compute.py:392-398) */
  int32_t use_power;      /* 1: sum |X|^2 W, 0: sum |X| W              (compute.py:221-226) */
  int32_t use_log;        /* 1: log(max(v, log_floor))                 (compute.py:458-459) */
  float log_floor;        /* config.LOG_FLOOR_VALUE                                           */
  float preemph;          /* 0 = off; else y[i] = x[i] - c x[i-1]      (pre.py:136-149)     */
  float dither;           /* 0 = off; else x + N(0, dither^2)          (pre.py:90-104)      */
  int32_t dither_first;   /* order of the two fused pre-processors                            */
  const float* window;    /* h_ [L] analysis window                    (compute.py:343)     */
  /* Folded filter-bank weights W (F x (N/2+1)), one contiguous band per row.  Row f covers   *
   * bins [band_lo[f], band_lo[f] + band_len[f]); its taps start at weights[band_off[f]].     *
   * W is obtained by replaying the reference's segment loop (compute.py:416-457).            */
  const int32_t* band_lo;  /* h_ [F] */
  const int32_t* band_len; /* h_ [F] */
  const int64_t* band_off; /* h_ [F] */
  const float* weights;    /* h_ [sum band_len] */
} pds_stft_desc;

typedef struct pds_stft_plan pds_stft_plan;

/* One unit of work for the fused kernel: up to pds_stft_tile_frames() consecutive frames of one
 * utterance.  32 bytes, built by pds_stft_fill_tiles (host) and consumed on the device. */
typedef struct pds_tile {
  int64_t sig_off; /* index of the utterance's first sample in the packed signal buffer      */
  int32_t sig_len; /* samples in the utterance                                                */
  int32_t start;   /* first sample of the tile's first frame, relative to the utterance; <0   *
                    * or beyond sig_len means symmetric reflection                            */
  int32_t nframes; /* frames in this tile                                                     */
  int32_t utt;     /* utterance id (keys the dither stream)                                   */
  int64_t out_row; /* row of the packed (total_frames x num_coeffs) output of the first frame */
} pds_tile;

int pds_stft_plan_create(const pds_stft_desc* desc, int device, pds_stft_plan** plan);
void pds_stft_plan_destroy(pds_stft_plan* plan);
int pds_stft_num_coeffs(const pds_stft_plan* plan);
int pds_stft_tile_frames(const pds_stft_plan* plan);
/* 1 if the plan runs the shared-memory FFT kernel, 0 if it runs the generic direct-DFT kernel
 * (non power-of-two dft_size, dft_size outside [256, 2048]). */
int pds_stft_is_fast_path(const pds_stft_plan* plan);
/* Name of the kernel pds_stft_run launches for samples of `sig_dtype` (for benchmark records). */
const char* pds_stft_kernel_name(const pds_stft_plan* plan, int sig_dtype);

/* Frame count of a signal: 0 if sig_len < L/2 + 1 else (sig_len + S/2) / S  (compute.py:580-596) */
int64_t pds_stft_num_frames(const pds_stft_plan* plan, int64_t sig_len);

/* Prefix-sum the frame counts of n_utts signals into frame_off[0..n_utts] and count the tiles. */
int pds_stft_layout(const pds_stft_plan* plan, int64_t n_utts, const int64_t* h_sig_len,
                    int64_t* h_frame_off, int64_t* n_tiles);
/* Fill `tiles` (n_tiles entries as reported by pds_stft_layout). */
int pds_stft_fill_tiles(const pds_stft_plan* plan, int64_t n_utts, const int64_t* h_sig_off,
                        const int64_t* h_sig_len, const int64_t* h_frame_off, pds_tile* h_tiles);

/* Streaming support (compute.py:462-572): tiles for frames [first_frame, first_frame + nframes)
 * of ONE signal whose buffer starts `buf_origin` samples after the true start of the signal. */
int pds_stft_fill_tiles_range(const pds_stft_plan* plan, int64_t sig_off, int64_t buf_len,
                              int64_t buf_origin, int64_t first_frame, int64_t nframes,
                              int64_t out_row, pds_tile* h_tiles, int64_t* n_tiles);

/* Enqueue the fused kernel.  d_out is (total_frames x num_coeffs) float32, row-major; d_tiles
 * must be 16-byte aligned (descriptors are prefetched with 16-byte asynchronous copies). */
int pds_stft_run(pds_stft_plan* plan, const void* d_signal, int sig_dtype, const pds_tile* d_tiles,
                 int64_t n_tiles, float* d_out, uint64_t seed, void* stream);

/* Host-buffer convenience used by non-PyTorch hosts: packs nothing, copies h_signal
 * (total_samples elements of sig_dtype) to the device, runs, copies (total_frames x C) back.
 * Synchronous.  h_frame_off (n_utts + 1) receives the row offsets of each utterance. */
int pds_stft_compute_host(pds_stft_plan* plan, const void* h_signal, int sig_dtype,
                          int64_t total_samples, int64_t n_utts, const int64_t* h_sig_off,
                          const int64_t* h_sig_len, float* h_out, int64_t out_capacity_rows,
                          int64_t* h_frame_off, uint64_t seed);

/* ------------------------------------------------------------------------------------------ *
 * Pre-processors as stand-alone passes (used when they cannot be fused)                        *
 * ------------------------------------------------------------------------------------------ */
/* pre.py:136-149 ; one launch over n_utts packed signals, in float32, out of place */
int pds_preemphasize(const float* d_in, float* d_out, int64_t n_utts, const int64_t* d_sig_off,
                     const int64_t* d_sig_len, int64_t total_samples, float coeff, void* stream);
/* pre.py:90-104 ; Philox stream keyed by (seed, utterance, sample) */
int pds_dither(const float* d_in, float* d_out, int64_t n_utts, const int64_t* d_sig_off,
               const int64_t* d_sig_len, int64_t total_samples, float coeff, uint64_t seed,
               void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Post-processors                                                                              *
 * ------------------------------------------------------------------------------------------ */
/* post.Deltas.apply(axis=time) with edge padding and concatenation on the coefficient axis
 * (post.py:441-491).  d_in is (total_rows x n_cols); d_out is (total_rows x n_cols*(orders+1)).
 * d_row_off (n_utts + 1) delimits utterances: the time filter never crosses them.
 * h_filters holds the `orders` correlation filters back to back, filter i having
 * 2*i*context+1... taps given by h_filter_len[i]. */
int pds_deltas(const float* d_in, float* d_out, int64_t total_rows, int32_t n_cols,
               int64_t n_utts, const int64_t* d_row_off, int32_t orders,
               const float* h_filters, const int32_t* h_filter_len, void* stream);

/* post.Standardize.accumulate (post.py:175-191): d_stats is (2 x (n_cols+1)) float64 laid out
 * like Kaldi CMVN stats [[sum x..., count], [sum x^2..., 0]] and is ADDED to. */
int pds_cmvn_accumulate(const float* d_feats, int64_t n_rows, int32_t n_cols, double* d_stats,
                        void* stream);
/* post.Standardize.apply (post.py:250-295) given stats already reduced over ranks.
 * y = x * scale[c] - mean[c] * scale[c]; d_out may alias d_feats.  d_zero_var (int32, may be
 * NULL) is set to 1 if a variance was ~0 and replaced by 1 (the shell turns that into the
 * reference's warning). */
int pds_cmvn_apply(const float* d_feats, float* d_out, int64_t n_rows, int32_t n_cols,
                   const double* d_stats, int32_t norm_var, int32_t* d_zero_var, void* stream);

/* Deltas followed by Standardize without materialising the deltas in between (the chain of
 * BASELINE config 5: post.py:441-491 then post.py:160-305).  Both calls compute the
 * (total_rows x 3*n_cols) Deltas output on the fly from d_in (total_rows x n_cols):
 *   pds_deltas_cmvn_accumulate ADDS its column sums / sums of squares / row count to d_stats,
 *   laid out (2 x (3*n_cols + 1)) like pds_cmvn_accumulate;
 *   pds_deltas_cmvn_apply writes the normalised output to d_out (total_rows x 3*n_cols).
 * Implemented for the default Deltas(num_deltas=2, context_window=2) filters (5 and 9 taps) and
 * n_cols <= 256; anything else returns PDS_ERR_UNSUPPORTED (use pds_deltas + pds_cmvn_*). */
int pds_deltas_cmvn_accumulate(const float* d_in, int64_t total_rows, int32_t n_cols, int64_t n_utts,
                               const int64_t* d_row_off, int32_t orders, const float* h_filters,
                               const int32_t* h_filter_len, double* d_stats, void* stream);
int pds_deltas_cmvn_apply(const float* d_in, float* d_out, int64_t total_rows, int32_t n_cols,
                          int64_t n_utts, const int64_t* d_row_off, int32_t orders,
                          const float* h_filters, const int32_t* h_filter_len, const double* d_stats,
                          int32_t norm_var, int32_t* d_zero_var, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Short-integration frame computer (compute.py:613-999; spec tests/test_compute.py:129-176)   *
 * ------------------------------------------------------------------------------------------ */
typedef struct pds_si_desc {
  int32_t frame_shift;    /* S                                                               */
  int32_t num_filts;      /* number of FIR filters INCLUDING the energy Dirac if any         */
  int32_t max_support;    /* taps per filter (all filters clamped to it, compute.py:742)     */
  int32_t pad_left;       /* zeros before the signal: max(0, S - translation) centered, else 0 */
  int32_t frame_start;    /* index into the full convolution of frame 0's pooling window     */
  int32_t frames_lost;    /* causal frame-count quirk d of SURVEY A.3                        */
  int32_t use_power;
  int32_t use_log;
  int32_t is_real;        /* 1: h_imag may be NULL                                           */
  float log_floor;
  const float* h_real;    /* h_ [num_filts x max_support] impulse responses, real part      */
  const float* h_imag;    /* h_ [num_filts x max_support] imaginary part                     */
  const float* window;    /* h_ [2*S] pooling window                                         */
} pds_si_desc;

typedef struct pds_si_plan pds_si_plan;
int pds_si_plan_create(const pds_si_desc* desc, int device, pds_si_plan** plan);
void pds_si_plan_destroy(pds_si_plan* plan);
/* max(0, (sig_len + S/2) / S - frames_lost)  (compute.py:824-847) */
int64_t pds_si_num_frames(const pds_si_plan* plan, int64_t sig_len);
int pds_si_tile_frames(const pds_si_plan* plan);
/* Same tiling protocol as the STFT computer; pds_tile.start is the index, in the full linear
 * convolution of the left-padded signal, of the first sample pooled by the tile's first frame. */
int pds_si_layout(const pds_si_plan* plan, int64_t n_utts, const int64_t* h_sig_len,
                  int64_t* h_frame_off, int64_t* n_tiles);
int pds_si_fill_tiles(const pds_si_plan* plan, int64_t n_utts, const int64_t* h_sig_off,
                      const int64_t* h_sig_len, const int64_t* h_frame_off, pds_tile* h_tiles);
/* d_signal is float32; d_out is (total_frames x num_filts) float32, row-major. */
int pds_si_run(pds_si_plan* plan, const float* d_signal, const pds_tile* d_tiles, int64_t n_tiles,
               float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDS_B200_H_ */
