"""Frequency scales used to lay out filter banks (host side, float64).

Reference: ``pydrobert/speech/scales.py:39-171``.  Scalar in, scalar out; only ever called while
a bank is being constructed, never on the per-frame path.
"""

import abc
import math

from .alias import AliasedFactory

__all__ = [
    "BarkScaling",
    "LinearScaling",
    "MelScaling",
    "OctaveScaling",
    "ScalingFunction",
]


class ScalingFunction(AliasedFactory):
    """An invertible map between Hertz and some perceptual / geometric scale"""

    @abc.abstractmethod
    def scale_to_hertz(self, scale: float) -> float:
        ...

    @abc.abstractmethod
    def hertz_to_scale(self, hertz: float) -> float:
        ...


class LinearScaling(ScalingFunction):
    """``scale = (hz - low_hz) * slope_hz``"""

    aliases = {"linear", "uniform"}

    def __init__(self, low_hz: float, slope_hz: float = 1.0):
        self.low_hz = low_hz
        self.slope_hz = slope_hz

    def scale_to_hertz(self, scale: float) -> float:
        return scale / self.slope_hz + self.low_hz

    def hertz_to_scale(self, hertz: float) -> float:
        return (hertz - self.low_hz) * self.slope_hz


class OctaveScaling(ScalingFunction):
    """``scale = log2(hz / low_hz)``"""

    aliases = {"octave"}

    def __init__(self, low_hz: float):
        if low_hz <= 0:
            raise ValueError("low_hz must be positive")
        self.low_hz = low_hz

    def scale_to_hertz(self, scale: float) -> float:
        return (2 ** scale) * max(1e-10, self.low_hz)

    def hertz_to_scale(self, hertz: float) -> float:
        return math.log2(hertz / max(1e-10, self.low_hz))


class MelScaling(ScalingFunction):
    """O'Shaughnessy's mel formula, ``scale = 1127 ln(1 + hz / 700)``"""

    aliases = {"mel"}

    def scale_to_hertz(self, scale: float) -> float:
        return 700.0 * (math.exp(scale / 1127.0) - 1.0)

    def hertz_to_scale(self, hertz: float) -> float:
        return 1127.0 * math.log(1 + hertz / 700.0)


class BarkScaling(ScalingFunction):
    """Traunmueller's Bark approximation with its low/high corrections"""

    aliases = {"bark"}

    _LOW, _HIGH = 2.0, 20.1

    def scale_to_hertz(self, scale: float) -> float:
        if scale < self._LOW:
            z = (20.0 * scale - 6.0) / 17.0
        elif scale > self._HIGH:
            z = (50.0 * scale + 221.1) / 61.0
        else:
            z = scale
        return 1960.0 * (z + 0.53) / (26.28 - z)

    def hertz_to_scale(self, hertz: float) -> float:
        z = 26.81 * hertz / (1960.0 + hertz) - 0.53
        if z < self._LOW:
            return z + 0.15 * (2.0 - z)
        if z > self._HIGH:
            return z + 0.22 * (z - 20.1)
        return z
