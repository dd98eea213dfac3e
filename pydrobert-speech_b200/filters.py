"""Filter banks and analysis windows: the host-side table builders of the frame computers.

Everything in this module runs once, in float64, when a computer is constructed.  It produces
the constant tables the CUDA kernels consume (window taps, truncated frequency responses,
impulse responses, supports).  The *values* follow the reference definitions
(``pydrobert/speech/filters.py``); the code is organised around NumPy array expressions rather
than the reference's per-sample Python loops, because a bank is rebuilt for every plan and the
gammatone / Gabor loops are the slow part of start-up.

Reference map (file:line in ``/root/reference/src/pydrobert/speech/filters.py``):

* ``LinearFilterBank`` interface ............ 49-237
* ``TriangularOverlappingFilterBank`` ....... 240-440
* ``Fbank`` ................................. 443-626
* ``GaborFilterBank`` ....................... 629-900
* ``ComplexGammatoneFilterBank`` ............ 903-1211
* windows ................................... 1217-1349
"""

import abc
import math

from typing import Mapping, Optional, Tuple, Union

import numpy as np

from . import config
from .alias import AliasedFactory, alias_factory_subclass_from_arg
from .scales import MelScaling, ScalingFunction
from .util import angular_to_hertz, hertz_to_angular

__all__ = [
    "BartlettWindow",
    "BlackmanWindow",
    "ComplexGammatoneFilterBank",
    "Fbank",
    "GaborFilterBank",
    "GammaWindow",
    "HammingWindow",
    "HannWindow",
    "LinearFilterBank",
    "TriangularOverlappingFilterBank",
    "WindowFunction",
]

_TWO_PI = 2 * np.pi


def _half_width(width: int) -> int:
    """Number of bins of a real signal's one-sided spectrum"""
    return (width + 1) // 2 if width % 2 else width // 2 + 1


def _check_hz_range(low_hz, high_hz, sampling_rate):
    # shared by fbank / gabor / gammatone (filters.py:488-493, 702-707, 983-988)
    if low_hz < 0 or (high_hz and (high_hz <= low_hz or high_hz > sampling_rate // 2)):
        # (the reference formats high_hz=None here and dies with a TypeError instead)
        top = float("nan") if high_hz is None else high_hz
        raise ValueError("Invalid frequency range: ({:.2f},{:.2f}".format(low_hz, top))


def _uniform_scale_points(scaling_function, low_hz, high_hz, num_filts, half_step):
    """Hertz values of points spaced uniformly on ``scaling_function`` between the two edges

    ``half_step=False`` gives the ``num_filts + 2`` triangle vertices; ``half_step=True`` gives
    the ``num_filts + 1`` mid-way intersection points used by the Gabor / gammatone banks.
    """
    lo = scaling_function.hertz_to_scale(low_hz)
    hi = scaling_function.hertz_to_scale(high_hz)
    delta = (hi - lo) / (num_filts + 1)
    if half_step:
        return tuple(
            scaling_function.scale_to_hertz(lo + delta * (i + 0.5))
            for i in range(num_filts + 1)
        )
    return tuple(
        scaling_function.scale_to_hertz(lo + delta * i) for i in range(num_filts + 2)
    )


class LinearFilterBank(AliasedFactory):
    """A fixed set of LTI filters, lowest centre frequency first

    Sub-classes provide each filter in the time domain (:func:`get_impulse_response`), in the
    frequency domain (:func:`get_frequency_response`) and as the non-negligible slice of the
    frequency response (:func:`get_truncated_response`), plus effective supports in samples
    (:obj:`supports`) and Hertz (:obj:`supports_hz`).
    """

    @abc.abstractproperty
    def is_real(self) -> bool:
        ...

    @abc.abstractproperty
    def is_analytic(self) -> bool:
        ...

    @abc.abstractproperty
    def is_zero_phase(self) -> bool:
        ...

    @abc.abstractproperty
    def num_filts(self) -> int:
        ...

    @abc.abstractproperty
    def sampling_rate(self) -> float:
        ...

    @abc.abstractproperty
    def supports_hz(self) -> Tuple[Tuple[float, float], ...]:
        ...

    @abc.abstractproperty
    def supports(self) -> Tuple[Tuple[float, float], ...]:
        ...

    @property
    def supports_ms(self) -> Tuple[Tuple[float, float], ...]:
        rate = self.sampling_rate
        return tuple((lo * 1000 / rate, hi * 1000 / rate) for lo, hi in self.supports)

    @abc.abstractmethod
    def get_impulse_response(self, filt_idx: int, width: int) -> np.ndarray:
        ...

    @abc.abstractmethod
    def get_frequency_response(
        self, filt_idx: int, width: int, half: bool = False
    ) -> np.ndarray:
        ...

    @abc.abstractmethod
    def get_truncated_response(
        self, filt_idx: int, width: int
    ) -> Tuple[int, np.ndarray]:
        ...


class _VertexBank(LinearFilterBank):
    """Shared plumbing of the two triangular banks: ``num_filts + 2`` vertices in Hertz"""

    _vertices: Tuple[float, ...]
    _rate: float
    _analytic: bool

    @property
    def is_real(self) -> bool:
        return not self._analytic

    @property
    def is_analytic(self) -> bool:
        return self._analytic

    @property
    def is_zero_phase(self) -> bool:
        return True

    @property
    def num_filts(self) -> int:
        return len(self._vertices) - 2

    @property
    def sampling_rate(self) -> float:
        return self._rate

    @property
    def centers_hz(self) -> Tuple[float, ...]:
        """Frequency of maximum gain of each filter"""
        return self._vertices[1:-1]

    @property
    def supports_hz(self) -> Tuple[Tuple[float, float], ...]:
        return tuple(zip(self._vertices[:-2], self._vertices[2:]))

    def _angular_vertices(self, filt_idx):
        return tuple(
            hertz_to_angular(self._vertices[filt_idx + i], self._rate) for i in range(3)
        )

    def _bin_range(self, filt_idx, width):
        # first / last DFT bin inside the triangle (filters.py:401-404, 581-584)
        left, right = self._vertices[filt_idx], self._vertices[filt_idx + 2]
        left_idx = int(np.ceil(width * left / self._rate))
        right_idx = int(width * right / self._rate)
        assert self._rate * (left_idx - 1) / width <= left
        assert self._rate * (right_idx + 1) / width >= right, width
        return left_idx, right_idx

    def _gain(self, filt_idx, bins, width):
        """Gain of filter ``filt_idx`` at integer DFT bins ``bins`` (array) of a ``width`` DFT"""
        raise NotImplementedError

    def get_frequency_response(
        self, filt_idx: int, width: int, half: bool = False
    ) -> np.ndarray:
        left_idx, right_idx = self._bin_range(filt_idx, width)
        size = _half_width(width) if half else width
        res = np.zeros(size, dtype=np.float64)
        bins = np.arange(left_idx, min(size, right_idx + 1))
        if len(bins):
            vals = self._gain(filt_idx, bins, width)
            res[bins] = vals
            if not half and not self._analytic:
                res[-bins] = vals  # Hermitian image (bin 0 maps onto itself)
        return res


class TriangularOverlappingFilterBank(_VertexBank):
    """Triangles (in frequency) whose vertices are uniform on a scale

    Reference: ``filters.py:240-440``.
    """

    aliases = {"tri", "triangular"}

    def __init__(
        self,
        scaling_function: Union[ScalingFunction, Mapping, str],
        num_filts: int = 40,
        high_hz: Optional[float] = None,
        low_hz: float = 20.0,
        sampling_rate: float = 16000,
        analytic: bool = False,
    ):
        scaling_function = alias_factory_subclass_from_arg(
            ScalingFunction, scaling_function
        )
        nyquist = sampling_rate / 2
        if high_hz is None:
            high_hz = nyquist
        # 1 Hz of slack for serialisation round-off (filters.py:292-296)
        if not (0 <= low_hz < high_hz <= nyquist + 1):
            raise ValueError(
                "Invalid frequency range: ({:.2f},{:.2f}".format(low_hz, high_hz)
            )
        high_hz = min(high_hz, nyquist)
        self._rate = sampling_rate
        self._vertices = _uniform_scale_points(
            scaling_function, low_hz, high_hz, num_filts, False
        )
        self._analytic = analytic

    @property
    def supports(self) -> Tuple[Tuple[float, float], ...]:
        # envelope bound 2(w_r - w_l) / ((w_c - w_l)(w_r - w_c) t^2 pi)   (filters.py:344-358)
        out = []
        for idx in range(self.num_filts):
            left, mid, right = self._angular_vertices(idx)
            K = np.sqrt(8 * (right - left) / np.pi)
            K /= np.sqrt(config.EFFECTIVE_SUPPORT_THRESHOLD)
            K /= np.sqrt(mid - left) * np.sqrt(right - mid)
            K = int(np.ceil(K))
            out.append((-K // 2 - 1, K // 2 + 1))
        return tuple(out)

    def get_impulse_response(self, filt_idx: int, width: int) -> np.ndarray:
        # closed-form inverse transform of a triangle, periodised by folding t onto the buffer
        # (filters.py:360-393)
        left, mid, right = self._angular_vertices(filt_idx)
        if right - mid > mid - left:
            denom, div = right - mid, mid - left
        else:
            denom, div = mid - left, right - mid
        denom *= (int(self._analytic) + 1) * np.pi
        t = np.arange(1, width + 1, dtype=np.float64)
        if self._analytic:
            basis = lambda w: np.exp(1j * w * t)  # noqa: E731
        else:
            basis = lambda w: np.cos(w * t)  # noqa: E731
        numer = (right - left) / div * basis(mid)
        numer -= (right - mid) / div * basis(left)
        numer -= (mid - left) / div * basis(right)
        vals = numer / t ** 2
        res = np.zeros(width, dtype=np.complex128 if self._analytic else np.float64)
        inner = vals[: width - 1]  # t = 1 .. width-1 land on res[t] and res[-t]
        res[1:] += inner
        res[1:] += np.conj(inner[::-1])
        res[0] += vals[width - 1]  # t == width wraps onto sample 0
        dc = mid / div * (right ** 2 - left ** 2)
        dc += right / div * (left ** 2 - mid ** 2)
        dc += left / div * (mid ** 2 - right ** 2)
        res[0] += dc / 2
        res /= denom
        return res

    def _gain(self, filt_idx, bins, width):
        left, mid, right = self._vertices[filt_idx : filt_idx + 3]
        hz = self._rate * bins / width
        rising = (hz - left) / (mid - left)
        falling = (right - hz) / (right - mid)
        return np.where(hz <= mid, rising, falling)

    def get_truncated_response(
        self, filt_idx: int, width: int
    ) -> Tuple[int, np.ndarray]:
        # note: buffer is always 1 + right - left long, even if clipped at `width`
        # (filters.py:433-440)
        left_idx, right_idx = self._bin_range(filt_idx, width)
        res = np.zeros(1 + right_idx - left_idx, dtype=np.float64)
        bins = np.arange(left_idx, min(width, right_idx + 1))
        if len(bins):
            res[bins - left_idx] = self._gain(filt_idx, bins, width)
        return left_idx, res


class Fbank(_VertexBank):
    """Kaldi/HTK style bank: triangular *in mel*, square-rooted so that power-after-filtering
    reproduces filtering-after-power

    Reference: ``filters.py:443-626``.
    """

    aliases = {"fbank"}

    def __init__(
        self,
        num_filts: int = 40,
        high_hz: Optional[float] = None,
        low_hz: float = 20.0,
        sampling_rate: float = 16000,
        analytic: bool = False,
    ):
        _check_hz_range(low_hz, high_hz, sampling_rate)
        self._rate = sampling_rate
        if high_hz is None:
            high_hz = sampling_rate // 2
        self._vertices = _uniform_scale_points(
            MelScaling(), low_hz, high_hz, num_filts, False
        )
        self._analytic = analytic

    @property
    def supports(self) -> Tuple[Tuple[float, float], ...]:
        # decay bound of the square-rooted triangle (filters.py:542-560)
        eps = config.EFFECTIVE_SUPPORT_THRESHOLD
        out = []
        for idx in range(self.num_filts):
            left, mid, right = self._angular_vertices(idx)
            K = right - left + 2 * ((right - mid) * (mid - left)) ** 2
            K /= eps ** 2 * np.pi
            K /= (right - mid) * (mid - left)
            K /= np.sqrt(eps)
            K /= np.sqrt(mid - left) * np.sqrt(right - mid)
            K **= 0.3333
            K = int(np.ceil(K))
            out.append((-K // 2 - 1, K // 2 + 1))
        return tuple(out)

    def get_impulse_response(self, filt_idx: int, width: int) -> np.ndarray:
        # numerically inverted, like the reference (filters.py:562-569)
        if self.is_analytic:
            return np.fft.ifft(self.get_frequency_response(filt_idx, width, half=False))
        half = self.get_frequency_response(filt_idx, width, half=True)
        return np.fft.irfft(half, n=width)

    def _mel_triangle(self, filt_idx, bins, width):
        mel = MelScaling()
        left, mid, right = (
            mel.hertz_to_scale(v) for v in self._vertices[filt_idx : filt_idx + 3]
        )
        pos = 1127.0 * np.log(1 + (self._rate * bins / width) / 700.0)
        rising = (pos - left) / (mid - left)
        falling = (right - pos) / (right - mid)
        return np.where(pos <= mid, rising, falling)

    def _gain(self, filt_idx, bins, width):
        return self._mel_triangle(filt_idx, bins, width) ** 0.5

    def get_truncated_response(
        self, filt_idx: int, width: int
    ) -> Tuple[int, np.ndarray]:
        left_idx, right_idx = self._bin_range(filt_idx, width)
        stop = min(width, right_idx + 1)
        res = np.zeros(stop - left_idx, dtype=np.float64)
        bins = np.arange(left_idx, stop)
        if len(bins):
            res[:] = self._mel_triangle(filt_idx, bins, width)
        return left_idx, res ** 0.5


class _IntersectBank(LinearFilterBank):
    """Plumbing shared by the Gabor and gammatone banks (complex, laid out by intersections)"""

    _rate: float
    _centers_hz: Tuple[float, ...]
    _supports: Tuple[Tuple[int, int], ...]
    _supports_ang: Tuple[Tuple[float, float], ...]
    _wrap_below: bool
    _scale_l2_norm: bool
    _erb: bool

    @property
    def is_real(self) -> bool:
        return False

    @property
    def is_analytic(self) -> bool:
        return not self._wrap_below

    @property
    def num_filts(self) -> int:
        return len(self._centers_hz)

    @property
    def sampling_rate(self) -> float:
        return self._rate

    @property
    def centers_hz(self) -> Tuple[float, ...]:
        """Frequency of maximum gain of each filter"""
        return self._centers_hz

    @property
    def supports_hz(self) -> Tuple[Tuple[float, float], ...]:
        return tuple(
            (angular_to_hertz(lo, self._rate), angular_to_hertz(hi, self._rate))
            for lo, hi in self._supports_ang
        )

    @property
    def supports(self) -> Tuple[Tuple[float, float], ...]:
        return self._supports

    @property
    def scaled_l2_norm(self) -> bool:
        return self._scale_l2_norm

    @property
    def erb(self) -> bool:
        return self._erb


class GaborFilterBank(_IntersectBank):
    r"""Complex Gaussian-envelope filters

    .. math:: \widehat{f}(\omega) = C \sqrt{2\sigma} \pi^{1/4} e^{-\sigma^2(\xi-\omega)^2/2}

    Adjacent filters intersect at their ERB (``erb=True``) or 3 dB (``erb=False``) bandwidth.
    Reference: ``filters.py:629-900``.
    """

    aliases = {"gabor"}

    def __init__(
        self,
        scaling_function: Union[ScalingFunction, Mapping, str],
        num_filts: int = 40,
        high_hz: Optional[float] = None,
        low_hz: float = 20.0,
        sampling_rate: float = 16000,
        scale_l2_norm: bool = False,
        erb: bool = False,
    ):
        scaling_function = alias_factory_subclass_from_arg(
            ScalingFunction, scaling_function
        )
        self._scale_l2_norm = scale_l2_norm
        self._erb = erb
        _check_hz_range(low_hz, high_hz, sampling_rate)
        self._rate = sampling_rate
        if high_hz is None:
            high_hz = sampling_rate // 2
        edges = _uniform_scale_points(
            scaling_function, low_hz, high_hz, num_filts, True
        )
        log_2, log_pi = np.log(2), np.log(np.pi)
        t_const = -2 * np.log(config.EFFECTIVE_SUPPORT_THRESHOLD)
        f_const = t_const
        if scale_l2_norm:
            f_const += log_2 + 0.5 * log_pi
            t_const -= 0.5 * log_pi
        else:
            t_const -= log_2 + log_pi
        bw_const = np.sqrt(np.pi) / 2 if erb else np.sqrt(3 / 10 * np.log(10))
        centers_hz, centers_ang, stds = [], [], []
        supports, supports_ang, wrap_supports_ang = [], [], []
        self._wrap_below = False
        for lo_edge, hi_edge in zip(edges[:-1], edges[1:]):
            center_hz = (lo_edge + hi_edge) / 2
            center_ang = hertz_to_angular(center_hz, sampling_rate)
            std = bw_const / hertz_to_angular(center_hz - lo_edge, sampling_rate)
            log_std = np.log(std)
            if scale_l2_norm:
                diff_ang = np.sqrt(log_std + f_const) / std
                wrap_diff_ang = np.sqrt(log_std + f_const + log_2) / std
                diff_samps = int(np.ceil(std * np.sqrt(t_const - log_std)))
            else:
                diff_ang = np.sqrt(f_const) / std
                wrap_diff_ang = np.sqrt(f_const + log_2) / std
                diff_samps = int(np.ceil(std * np.sqrt(t_const - 2 * log_std)))
            if center_ang - diff_ang < 0:
                self._wrap_below = True
            centers_hz.append(center_hz)
            centers_ang.append(center_ang)
            stds.append(std)
            supports_ang.append((center_ang - diff_ang, center_ang + diff_ang))
            wrap_supports_ang.append(2 * wrap_diff_ang)
            supports.append((-diff_samps, diff_samps))
        self._centers_hz = tuple(centers_hz)
        self._centers_ang = tuple(centers_ang)
        self._stds = tuple(stds)
        self._supports_ang = tuple(supports_ang)
        self._wrap_supports_ang = tuple(wrap_supports_ang)
        self._supports = tuple(supports)

    @property
    def is_zero_phase(self) -> bool:
        return True

    def get_impulse_response(self, filt_idx: int, width: int) -> np.ndarray:
        # one period folded onto the buffer from both sides (filters.py:823-839)
        xi, std = self._centers_ang[filt_idx], self._stds[filt_idx]
        if self._scale_l2_norm:
            const = -0.5 * np.log(std) - 0.25 * np.log(np.pi)
        else:
            const = -0.5 * np.log(2 * np.pi) - np.log(std)
        t = np.arange(width + 1, dtype=np.float64)
        vals = np.exp(-(t ** 2) / (2 * std ** 2) + const + 1j * xi * t)
        res = np.zeros(width, dtype=np.complex128)
        res += vals[:width]  # t = 0 .. width-1 at res[t]
        res[1:] += np.conj(vals[1:width][::-1])  # t = 1 .. width-1 mirrored at res[-t]
        res[0] += np.conj(vals[width])  # t = width mirrored onto sample 0
        return res

    def _gaussian(self, filt_idx, bins, width, periods):
        xi, std = self._centers_ang[filt_idx], self._stds[filt_idx]
        const = 0.5 * np.log(2 * std) + 0.25 * np.log(np.pi) if self._scale_l2_norm else 0
        scale = -(std ** 2) / 2
        res = np.zeros(len(bins), dtype=np.float64)
        for period in periods:  # few terms; keep the reference's summation order
            omega = (bins / width + period) * 2 * np.pi
            res += np.exp(scale * (xi - omega) ** 2 + const)
        return res

    def get_frequency_response(
        self, filt_idx: int, width: int, half: bool = False
    ) -> np.ndarray:
        lowest, highest = self._supports_ang[filt_idx]
        size = _half_width(width) if half else width
        periods = range(
            -1 - int(max(-lowest, 0) / _TWO_PI), 2 + int(highest / _TWO_PI)
        )
        return self._gaussian(filt_idx, np.arange(size), width, periods)

    def get_truncated_response(
        self, filt_idx: int, width: int
    ) -> Tuple[int, np.ndarray]:
        # if even the half-threshold support spans the period, aliasing can lift any bin above
        # the threshold: the whole period is "support" (filters.py:873-879)
        if self._wrap_supports_ang[filt_idx] >= _TWO_PI:
            return 0, self.get_frequency_response(filt_idx, width)
        lowest, highest = self._supports_ang[filt_idx]
        left_idx = int(np.ceil(width * lowest / _TWO_PI))
        right_idx = int(width * highest / _TWO_PI)
        periods = range(-int(max(-lowest, 0) / _TWO_PI), 1 + int(highest / _TWO_PI))
        bins = np.arange(left_idx, right_idx + 1)
        return left_idx % width, self._gaussian(filt_idx, bins, width, periods)


class ComplexGammatoneFilterBank(_IntersectBank):
    r"""Gammatone envelopes on complex carriers

    .. math:: h(t) = c t^{n-1} e^{-\alpha t + i\xi t} u(t), \quad
              H(\omega) = \frac{c (n-1)!}{(\alpha + i(\omega - \xi))^n}

    Reference: ``filters.py:903-1211``.
    """

    aliases = {"gammatone", "tonebank"}

    def __init__(
        self,
        scaling_function: Union[ScalingFunction, Mapping, str],
        num_filts: int = 40,
        high_hz: Optional[float] = None,
        low_hz: float = 20.0,
        sampling_rate: float = 16000,
        order: int = 4,
        max_centered: bool = False,
        scale_l2_norm: bool = False,
        erb: bool = False,
    ):
        scaling_function = alias_factory_subclass_from_arg(
            ScalingFunction, scaling_function
        )
        self._scale_l2_norm = scale_l2_norm
        self._erb = erb
        _check_hz_range(low_hz, high_hz, sampling_rate)
        if not isinstance(order, int) or order <= 0:
            raise ValueError("order must be a positive integer")
        self._order = order
        self._rate = sampling_rate
        if high_hz is None:
            high_hz = sampling_rate // 2
        edges = _uniform_scale_points(
            scaling_function, low_hz, high_hz, num_filts, True
        )
        log_eps = np.log(config.EFFECTIVE_SUPPORT_THRESHOLD)
        log_fact2 = np.log(math.factorial(2 * order - 2))
        log_fact = np.log(math.factorial(order - 1))
        log_2 = np.log(2)
        if erb:
            alpha_const = log_2 * (2 * order - 1) + 2 * log_fact - log_fact2
        else:
            alpha_const = -0.5 * np.log(4 * (2 ** (1 / order)) - 4)
        centers_hz, xis, alphas, cs, offsets = [], [], [], [], []
        supports, supports_ang, wrap_supports_ang = [], [], []
        self._wrap_below = False
        for lo_edge, hi_edge in zip(edges[:-1], edges[1:]):
            center_hz = (lo_edge + hi_edge) / 2
            xi = hertz_to_angular(center_hz, sampling_rate)
            log_alpha = alpha_const + np.log(
                hertz_to_angular(hi_edge - lo_edge, sampling_rate)
            )
            alpha = np.exp(log_alpha)
            if scale_l2_norm:
                log_c = 0.5 * (log_2 + log_alpha + log_fact2)
                log_c -= order * (log_alpha + log_2)
            else:
                log_c = order * log_alpha - log_fact
            c = np.exp(log_c)
            offset = -(order - 1) / alpha if max_centered else 0
            supp_a = (2 / order) * (log_c + log_fact - log_eps)
            wrap_supp_a = supp_a + (2 / order) * log_2
            supp_b = np.exp(2 * log_alpha)
            diff_ang = (np.exp(supp_a) - supp_b) ** 0.5
            wrap_diff_ang = (np.exp(wrap_supp_a) - supp_b) ** 0.5
            centers_hz.append(center_hz)
            xis.append(xi)
            alphas.append(alpha)
            cs.append(c)
            offsets.append(offset)
            supports.append(self._temporal_support(alpha, c, xi, offset))
            supports_ang.append((xi - diff_ang, xi + diff_ang))
            if xi - diff_ang < 0:
                self._wrap_below = True
            wrap_supports_ang.append(2 * wrap_diff_ang)
        self._centers_hz = tuple(centers_hz)
        self._xis = tuple(xis)
        self._alphas = tuple(alphas)
        self._cs = tuple(cs)
        self._offsets = tuple(offsets)
        self._supports = tuple(supports)
        self._supports_ang = tuple(supports_ang)
        self._wrap_supports_ang = tuple(wrap_supports_ang)

    @property
    def order(self) -> int:
        return self._order

    @property
    def is_zero_phase(self) -> bool:
        return False

    def _envelope_at(self, t, alpha, c, xi, offset):
        # |h| and h itself at (possibly fractional) time(s) t; zero at or before the onset
        t = np.asarray(t, dtype=np.float64)
        shifted = t - offset
        live = shifted > 0
        safe = np.where(live, shifted, 1.0)
        r = np.log(c) + (self._order - 1) * np.log(safe) + (-alpha + 1j * xi) * safe
        return np.where(live, np.exp(r), 0j)

    def _temporal_support(self, alpha, c, xi, offset):
        # Newton descent along the decaying side of the envelope until it drops below the
        # threshold (filters.py:1187-1211)
        n, eps = self._order, config.EFFECTIVE_SUPPORT_THRESHOLD
        if n == 1:
            # sic: the reference parenthesises this as log(c) - log(eps) / alpha
            right = int(np.ceil((np.log(c) - np.log(eps) / alpha)))
        else:
            right = (n - 1 + np.sqrt((n - 1) / 2)) / alpha
            mag = float(np.abs(self._envelope_at(right, alpha, c, xi, offset)))
            while mag > eps:
                slope = c * np.exp(-alpha * right) * right ** (n - 2)
                slope *= (n - 1) - alpha * right
                right -= mag / slope
                mag = float(np.abs(self._envelope_at(right, alpha, c, xi, offset)))
        return (int(np.floor(offset)), int(np.ceil(right) + offset))

    def _params(self, filt_idx):
        return (
            self._alphas[filt_idx],
            self._cs[filt_idx],
            self._xis[filt_idx],
            self._offsets[filt_idx],
        )

    def get_impulse_response(self, filt_idx: int, width: int) -> np.ndarray:
        # sum of all periods that intersect the temporal support (filters.py:1116-1125)
        left_sup, right_sup = self._supports[filt_idx]
        first = int(np.floor(left_sup / width))
        last = int(np.ceil(right_sup / width))
        res = np.zeros(width, dtype=np.complex128)
        idx = np.arange(width)
        for period in range(first, last + 1):
            res += self._envelope_at(period * width + idx, *self._params(filt_idx))
        return res

    def _H(self, omega, filt_idx):
        alpha, c, xi, offset = self._params(filt_idx)
        n = self._order
        numer = np.exp(-1j * omega * offset) * c * math.factorial(n - 1)
        return numer / (alpha + 1j * (omega - xi)) ** n

    def get_frequency_response(
        self, filt_idx: int, width: int, half: bool = False
    ) -> np.ndarray:
        left_sup, right_sup = self._supports_ang[filt_idx]
        first = int(np.floor(left_sup / 2 / np.pi))
        last = int(np.ceil(right_sup / 2 / np.pi))
        size = _half_width(width) if half else width
        omega = np.arange(size, dtype=np.float64) * 2 * np.pi / width
        res = np.zeros(size, dtype=np.complex128)
        for period in range(first, last + 1):
            res += self._H(omega + 2 * np.pi * period, filt_idx)
        return res

    def get_truncated_response(
        self, filt_idx: int, width: int
    ) -> Tuple[int, np.ndarray]:
        left_sup, right_sup = self._supports_ang[filt_idx]
        # support plus the extra needed to reach half the threshold covers the period ->
        # periodisation may push any bin over the threshold (filters.py:1150-1156)
        if right_sup - left_sup + self._wrap_supports_ang[filt_idx] >= _TWO_PI:
            return 0, self.get_frequency_response(filt_idx, width)
        left_idx = int(np.ceil(width * left_sup / _TWO_PI))
        right_idx = int(width * right_sup / _TWO_PI)
        omega = np.arange(left_idx, right_idx + 1, dtype=np.float64)
        omega *= 2 * np.pi / width
        return left_idx % width, self._H(omega, filt_idx)


# --------------------------------------------------------------------------------------
# windows
# --------------------------------------------------------------------------------------


class WindowFunction(AliasedFactory):
    """A real low-pass taper applied to each frame (or to the SI pooling region)"""

    @abc.abstractmethod
    def get_impulse_response(self, width: int) -> np.ndarray:
        ...


class _CosineSumWindow(WindowFunction):
    """NumPy's classic windows, rescaled so that their taps sum to (about) one

    Reference: ``filters.py:1237-1298``.
    """

    _numpy_window = None
    _dc_gain = 1.0  # window mean for long windows; the normaliser is dc_gain * (width - 1)

    def get_impulse_response(self, width: int) -> np.ndarray:
        window = type(self)._numpy_window(width)
        window /= self._dc_gain * max(1, width - 1)
        return window


class BartlettWindow(_CosineSumWindow):
    aliases = {"bartlett", "triangular", "tri"}
    _numpy_window = staticmethod(np.bartlett)
    _dc_gain = 0.5


class BlackmanWindow(_CosineSumWindow):
    aliases = {"blackman", "black"}
    _numpy_window = staticmethod(np.blackman)
    _dc_gain = 0.42


class HammingWindow(_CosineSumWindow):
    aliases = {"hamming"}
    _numpy_window = staticmethod(np.hamming)
    _dc_gain = 0.54


class HannWindow(_CosineSumWindow):
    aliases = {"hanning", "hann"}
    _numpy_window = staticmethod(np.hanning)
    _dc_gain = 0.5


class GammaWindow(WindowFunction):
    r"""Time-reversed Gamma density :math:`t^{n-1} e^{-\alpha t}`, peaking at ``peak * width``

    The default window of causal computers.  Reference: ``filters.py:1301-1349``.
    """

    aliases = {"gamma"}

    def __init__(self, order: int = 4, peak: float = 0.75):
        self.order = order
        self.peak = peak

    def get_impulse_response(self, width: int) -> np.ndarray:
        if width <= 0:
            return np.array([], dtype=float)
        if width == 1:
            return np.array([1], dtype=float)
        ret = np.arange(width - 1, -1, -1, dtype=float)  # reflected time axis
        if self.order > 1:
            alpha = (self.order - 1) / (width - self.peak * width)
            live = width - 1  # the last tap (t = 0) stays exactly zero
        else:
            alpha = 5 / width  # roughly confine the exponential's support to the window
            live = width
        ln_c = self.order * np.log(alpha) - np.log(math.factorial(self.order - 1))
        t = ret[:live]
        ret[:live] = t ** (self.order - 1) * np.exp(-alpha * t + ln_c)
        return ret
