"""Oracle for ``SIFrameComputer.compute_full`` (test infrastructure).

The reference computes short-integration features by overlap-save FFT convolution
(``compute.py:774-996``).  Its own test-suite states the equivalent direct form
(``tests/test_compute.py:129-176``, class ``SIFrameComputerNpConvolve``); this module restates
that direct form, extended to the causal frame-count rule of ``compute.py:824-847`` (see
SURVEY.md appendix A.3).
"""

import numpy as np

LOG_FLOOR_VALUE = 1e-5


def si_num_frames(sig_len, frame_shift, frames_lost):
    return max(0, (sig_len + frame_shift // 2) // frame_shift - frames_lost)


def si_features(
    signal, impulse_responses, window, frame_shift, pad_left, frame_start, frames_lost,
    use_power, use_log, linear=False,
):
    """``impulse_responses`` is (num_coeffs, max_support) complex or real; ``window`` is (2S,)"""
    signal = np.asarray(signal, dtype=np.float64)
    num_frames = si_num_frames(len(signal), frame_shift, frames_lost)
    num_coeffs = len(impulse_responses)
    coeffs = np.zeros((num_frames, num_coeffs))
    if not num_frames:
        return coeffs
    max_support = impulse_responses.shape[1]
    need = frame_start + (num_frames + 1) * frame_shift + max_support
    padded = np.zeros(max(need, pad_left + len(signal)) + max_support)
    padded[pad_left : pad_left + len(signal)] = signal
    window = np.asarray(window, dtype=np.float64)
    for c, filt in enumerate(impulse_responses):
        y = np.convolve(padded, filt)
        u = (y * y.conj()).real if use_power else np.abs(y)
        for t in range(num_frames):
            begin = frame_start + t * frame_shift
            coeffs[t, c] = np.sum(u[begin : begin + 2 * frame_shift] * window)
    if use_log and not linear:
        coeffs = np.log(np.maximum(coeffs, LOG_FLOOR_VALUE))
    return coeffs
