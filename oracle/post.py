"""Oracle for ``post.Deltas`` and ``post.Standardize`` (test infrastructure).

``deltas`` restates ``post.py:441-491`` in the form of the reference's own executable spec
``KaldiDeltas`` (``tests/test_post.py:136-176``): an explicit clamped-index sum.
``cmvn_*`` restate ``post.py:160-191`` and ``post.py:250-295``.
"""

import numpy as np


def delta_filters(num_deltas, context_window=2):
    """post.py:455-460: f_0 = [1], f_{i+1} = f_i (*) ramp, ramp = [-W..W] / sum j^2"""
    ramp = np.arange(-context_window, context_window + 1, dtype=np.float64)
    ramp /= np.sum(ramp ** 2)
    filts = [np.ones(1)]
    for _ in range(num_deltas):
        filts.append(np.convolve(filts[-1], ramp))
    return filts


def deltas(features, num_deltas, context_window=2):
    """Deltas along axis 0 (time) with edge padding, concatenated on axis 1"""
    features = np.asarray(features, dtype=np.float64)
    num_frames = features.shape[0]
    out = [features]
    for filt in delta_filters(num_deltas, context_window)[1:]:
        half = (len(filt) - 1) // 2
        acc = np.zeros_like(features)
        for j, coeff in enumerate(filt):  # correlation: out[t] = sum_j f[j] x[clamp(t + j - half)]
            idx = np.clip(np.arange(num_frames) + j - half, 0, num_frames - 1)
            acc += coeff * features[idx]
        out.append(acc)
    return np.concatenate(out, axis=1)


def cmvn_accumulate(features, stats=None):
    """Kaldi-layout sufficient statistics (2, F+1) float64; post.py:175-191"""
    features = np.asarray(features)
    if stats is None:
        stats = np.zeros((2, features.shape[1] + 1))
    stats[0, -1] += features.shape[0]
    stats[0, :-1] += features.sum(axis=0, dtype=np.float64)
    stats[1, :-1] += np.square(features, dtype=np.float64).sum(axis=0)
    return stats


def cmvn_apply(features, stats, norm_var=True):
    """post.py:250-295 with accumulated stats; returns float64"""
    features = np.asarray(features, dtype=np.float64)
    count = stats[0, -1]
    means = stats[0, :-1] / count
    if norm_var:
        varss = stats[1, :-1] / count - means ** 2
        varss = np.where(np.isclose(varss, 0), 1.0, varss)
        scales = 1 / varss ** 0.5
    else:
        scales = np.ones_like(means)
    return features * scales - means * scales
