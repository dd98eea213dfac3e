"""Oracle for ``STFTFrameComputer.compute_full`` (test infrastructure, see package docstring).

Follows ``/root/reference/src/pydrobert/speech/compute.py``:

* framing / symmetric padding ........ ``compute_full``   lines 574-607
* energy, window, rFFT ............... ``_compute_frame`` lines 388-413
* per-filter segment walk ............ ``_compute_frame`` lines 416-460
* ``_power`` / ``_mag`` .............. lines 221-226

The tables (``window``, ``start_idxs``, ``truncated_filts``) are inputs: the tests feed either
the golden tables dumped from the reference or the product's own host tables, which are
themselves compared with the golden ones.  The FFT is ``numpy.fft.rfft`` (pocketfft), i.e. the
third-party routine behind the reference's ``USE_FFTPACK=False`` branch; the reference's own
test-suite pins both branches to each other (``tests/test_compute.py:113-126``).
"""

import numpy as np

LOG_FLOOR_VALUE = 1e-5  # config.py:53


def pad_left_for(frame_length, frame_shift, centered, kaldi_shift):
    """compute.py:582-587"""
    if not centered:
        return 0
    if kaldi_shift:
        return frame_length // 2 - frame_shift // 2
    return (frame_length + 1) // 2 - 1


def frame_signal(signal, frame_length, frame_shift, pad_left):
    """(num_frames, frame_length) float64 matrix of padded frames; compute.py:578-606"""
    signal = np.asarray(signal, dtype=np.float64)
    if len(signal) < frame_length // 2 + 1:
        return np.empty((0, frame_length))
    num_frames = max(0, (len(signal) + frame_shift // 2) // frame_shift)
    total_len = (num_frames - 1) * frame_shift - pad_left + frame_length
    pad_right = max(0, total_len - len(signal))
    if pad_left or pad_right:
        signal = np.pad(signal, (pad_left, pad_right), "symmetric")
    idx = frame_shift * np.arange(num_frames)[:, None] + np.arange(frame_length)[None, :]
    return signal[idx]


def _walk_filter(half_spect, start_idx, truncated_filt, use_power):
    """One filter of compute.py:416-455, vectorised over the leading (frame) axis only"""
    half_len = half_spect.shape[-1]
    trunc_len = len(truncated_filt)
    odd = half_len % 2
    consumed, conjugate = 0, False
    val = np.zeros(half_spect.shape[:-1])
    while consumed < trunc_len:
        if conjugate:
            seg_len = max(0, min(start_idx + trunc_len - consumed, half_len - 2 + odd) - start_idx)
            if seg_len:
                first = -2 + odd - start_idx
                seg = half_spect[..., first : first - seg_len : -1].conj()
                prod = seg * truncated_filt[consumed : consumed + seg_len]
                val = val + (np.sum(np.abs(prod) ** 2, -1) if use_power else np.sum(np.abs(prod), -1))
            start_idx -= half_len - 2 + odd
        else:
            seg_len = max(0, min(start_idx + trunc_len - consumed, half_len) - start_idx)
            if seg_len:
                seg = half_spect[..., start_idx : start_idx + seg_len]
                prod = seg * truncated_filt[consumed : consumed + seg_len]
                val = val + (np.sum(np.abs(prod) ** 2, -1) if use_power else np.sum(np.abs(prod), -1))
            start_idx -= half_len
        conjugate = not conjugate
        consumed += seg_len
        start_idx = max(0, start_idx)
    return val


def stft_features(
    signal,
    window,
    dft_size,
    start_idxs,
    truncated_filts,
    frame_shift,
    pad_left,
    use_power,
    use_log,
    include_energy,
    is_real,
    linear=False,
):
    """Float64 features ``(num_frames, num_filts + include_energy)``

    ``linear=True`` returns the values before the log regardless of ``use_log`` (used for the
    relative tolerance on linear power / magnitude).
    """
    frame_length = len(window)
    frames = frame_signal(signal, frame_length, frame_shift, pad_left)
    num_coeffs = len(start_idxs) + int(bool(include_energy))
    coeffs = np.zeros((frames.shape[0], num_coeffs))
    if not frames.shape[0]:
        return coeffs
    col = 0
    if include_energy:  # compute.py:392-398, on the raw frame
        energy = np.einsum("tl,tl->t", frames, frames) / frame_length
        coeffs[:, 0] = energy if use_power else energy ** 0.5
        col = 1
    half_spect = np.fft.rfft(frames * np.asarray(window, dtype=np.float64), n=dft_size, axis=-1)
    for f, (start_idx, filt) in enumerate(zip(start_idxs, truncated_filts)):
        val = _walk_filter(half_spect, int(start_idx), np.asarray(filt), use_power)
        coeffs[:, col + f] = 2 * val if is_real else val  # compute.py:456-457
    if use_log and not linear:
        coeffs = np.log(np.maximum(coeffs, LOG_FLOOR_VALUE))  # compute.py:458-459
    return coeffs


def stft_features_looped(
    signal, window, dft_size, start_idxs, truncated_filts, frame_shift, pad_left,
    use_power, use_log, include_energy, is_real,
):
    """The same computation frame by frame, filter by filter -- the reference's own loop
    structure (one ``_walk_filter`` call per (frame, filter)); used on small cases to pin the
    vectorised form and as the "loop-faithful" CPU timing in bench.py."""
    frame_length = len(window)
    frames = frame_signal(signal, frame_length, frame_shift, pad_left)
    num_coeffs = len(start_idxs) + int(bool(include_energy))
    coeffs = np.zeros((frames.shape[0], num_coeffs))
    window = np.asarray(window, dtype=np.float64)
    for t in range(frames.shape[0]):
        frame, out = frames[t], coeffs[t]
        col = 0
        if include_energy:
            out[0] = np.inner(frame, frame) / frame_length
            if not use_power:
                out[0] **= 0.5
            if use_log:
                out[0] = np.log(max(out[0], LOG_FLOOR_VALUE))
            col = 1
        half_spect = np.fft.rfft(frame * window, n=dft_size)
        for f, (start_idx, filt) in enumerate(zip(start_idxs, truncated_filts)):
            val = float(_walk_filter(half_spect, int(start_idx), np.asarray(filt), use_power))
            if is_real:
                val *= 2
            if use_log:
                val = np.log(max(val, LOG_FLOOR_VALUE))
            out[col + f] = val
    return coeffs
