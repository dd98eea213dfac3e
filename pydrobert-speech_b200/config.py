"""Package-level constants (reference: ``pydrobert/speech/config.py:27-53``).

These are read at *call time* by the table builders, exactly like the reference, so tests that
monkey-patch them keep working.
"""

from typing import Set

__all__ = [
    "EFFECTIVE_SUPPORT_THRESHOLD",
    "LOG_FLOOR_VALUE",
    "SOUNDFILE_SUPPORTED_FILE_TYPES",
    "USE_FFTPACK",
]

#: Kept for config compatibility only.  The FFT runs inside the CUDA kernels; the flag is a no-op.
USE_FFTPACK: bool = False

#: Magnitude below which a filter response is treated as zero when computing supports.
EFFECTIVE_SUPPORT_THRESHOLD: float = 5e-4

#: Floor applied before every logarithm.
LOG_FLOOR_VALUE: float = 1e-5

#: soundfile is not part of the hot path; the set is empty unless soundfile is importable.
SOUNDFILE_SUPPORTED_FILE_TYPES: Set[str] = set()

try:  # pragma: no cover - optional dependency, absent in the build image
    import soundfile as _sf

    SOUNDFILE_SUPPORTED_FILE_TYPES = {"wav", "ogg", "flac", "aiff"} & {
        x.lower() for x in _sf.available_formats()
    }
except ImportError:
    pass
