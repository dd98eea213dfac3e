"""Host-side constant tables consumed by the CUDA kernels (float64 in, float32 out).

* :func:`stft_geometry` -- frame count / padding rules of ``STFTFrameComputer.compute_full``
  (reference ``compute.py:578-600``).
* :func:`fold_filters` -- the folded, real, non-negative weight matrix ``W`` such that::

      feat[t, f] = sum_k |rfft(frame_t * window, N)[k]| ** p * W[f, k]

  reproduces the reference's per-filter segment walk (``compute.py:416-457``).  ``W`` is obtained
  by *replaying that walk on bin indices*, not from the mathematically exact Hermitian fold:
  the reference's conjugate branch keys its offset on ``half_len % 2`` where ``dft_size % 2``
  was meant, which shifts the negative-frequency image by one bin for every power-of-two DFT
  size; parity is defined against that behaviour.
"""

from typing import List, NamedTuple, Sequence, Tuple

import numpy as np

__all__ = ["BandedWeights", "fold_filters", "half_spectrum_len", "stft_geometry"]


def half_spectrum_len(dft_size: int) -> int:
    """``len(np.fft.rfft(x, dft_size))``"""
    return dft_size // 2 + 1


def stft_pad_left(frame_length: int, frame_shift: int, centered: bool, kaldi_shift: bool) -> int:
    if not centered:
        return 0
    if kaldi_shift:
        return frame_length // 2 - frame_shift // 2
    return (frame_length + 1) // 2 - 1


def stft_geometry(
    sig_len: int, frame_length: int, frame_shift: int, centered: bool, kaldi_shift: bool
) -> Tuple[int, int, int]:
    """``(num_frames, pad_left, pad_right)`` for a signal of ``sig_len`` samples"""
    pad_left = stft_pad_left(frame_length, frame_shift, centered, kaldi_shift)
    if sig_len < frame_length // 2 + 1:
        return 0, pad_left, 0
    num_frames = max(0, (sig_len + frame_shift // 2) // frame_shift)
    total = (num_frames - 1) * frame_shift - pad_left + frame_length
    return num_frames, pad_left, max(0, total - sig_len)


class BandedWeights(NamedTuple):
    """``W`` stored one contiguous band per row"""

    lo: np.ndarray  # (F,) int32 first bin of each band
    length: np.ndarray  # (F,) int32 band lengths
    offset: np.ndarray  # (F,) int64 start of each band in ``taps``
    taps: np.ndarray  # (sum length,) float32
    dense: np.ndarray  # (F, K) float64, kept for tests / the oracle hand-off

    @property
    def nnz(self) -> int:
        return int(self.length.sum())


def _segment_bins(start: int, filt_len: int, half_len: int) -> List[np.ndarray]:
    """Bin index touched by each tap of a truncated response that starts at ``start``

    Follows the alternating plain / conjugate-reversed segments of ``compute.py:423-455``.
    """
    odd = half_len % 2  # sic -- the reference tests the half-spectrum length, see module doc
    mirror_len = half_len - 2 + odd
    pieces, consumed, conjugate = [], 0, False
    while consumed < filt_len:
        if conjugate:
            seg = max(0, min(start + filt_len - consumed, mirror_len) - start)
            # python slice half_spect[-2 + odd - start : -2 + odd - start - seg : -1]
            pieces.append(mirror_len - start - np.arange(seg))
            start -= mirror_len
        else:
            seg = max(0, min(start + filt_len - consumed, half_len) - start)
            pieces.append(start + np.arange(seg))
            start -= half_len
        conjugate = not conjugate
        consumed += seg
        start = max(0, start)
    return pieces


def fold_filters(
    start_idxs: Sequence[int],
    truncated_filts: Sequence[np.ndarray],
    dft_size: int,
    use_power: bool,
    is_real: bool,
) -> BandedWeights:
    """Fold truncated frequency responses onto the half spectrum (see module docstring)"""
    half_len = half_spectrum_len(dft_size) if dft_size % 2 == 0 else (dft_size + 1) // 2
    num_filts = len(start_idxs)
    dense = np.zeros((num_filts, half_len), dtype=np.float64)
    for f, (start, filt) in enumerate(zip(start_idxs, truncated_filts)):
        gain = np.abs(np.asarray(filt)) ** (2 if use_power else 1)
        bins = np.concatenate(_segment_bins(int(start), len(gain), half_len) or [np.arange(0)])
        np.add.at(dense[f], bins.astype(np.int64), gain)
    if is_real:
        dense *= 2  # Hermitian twin of a real filter; doubles DC / Nyquist too, like the reference
    lo = np.zeros(num_filts, dtype=np.int32)
    length = np.zeros(num_filts, dtype=np.int32)
    for f in range(num_filts):
        nz = np.flatnonzero(dense[f])
        if len(nz):  # hull of the non-zeros; interior zeros are stored explicitly
            lo[f], length[f] = nz[0], nz[-1] - nz[0] + 1
    offset = np.zeros(num_filts, dtype=np.int64)
    offset[1:] = np.cumsum(length[:-1])
    taps = np.concatenate(
        [dense[f, lo[f] : lo[f] + length[f]] for f in range(num_filts)] + [np.zeros(0)]
    ).astype(np.float32)
    return BandedWeights(lo, length, offset, taps, dense)
