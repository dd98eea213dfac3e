"""Small host-side helpers: unit conversions and the signal reader used by the CLI.

Reading files is *not* part of the accelerated path (SURVEY.md section 8 marks the reference's
``util.read_signal`` out of scope); the subset below keeps the reference's call signature and
format-sniffing rules (``pydrobert/speech/util.py:339-510``) for the formats that need no
optional dependency (wav, npy, npz, pt, raw binary) and defers to ``h5py`` / ``soundfile`` when
they happen to be installed.  Kaldi tables and NIST SPHERE need packages that are not part of
this build and raise ``IOError``.
"""

import re

from typing import Any, BinaryIO, Optional, Union

import numpy as np

from . import config

__all__ = ["angular_to_hertz", "hertz_to_angular", "read_signal"]


def hertz_to_angular(hertz: float, samp_rate: float) -> float:
    """Cycles/second -> radians/sample (reference ``util.py:108-110``)"""
    return hertz * 2 * np.pi / samp_rate


def angular_to_hertz(angle: float, samp_rate: float) -> float:
    """Radians/sample -> cycles/second (reference ``util.py:113-115``)"""
    return angle * samp_rate / (2 * np.pi)


def _cast(data, dtype):
    return data.astype(dtype) if dtype else data


def _read_wav(rfilename, dtype, key, **kwargs):
    try:
        from scipy.io import wavfile

        _, data = wavfile.read(rfilename, **kwargs)
    except ImportError:
        import wave

        with wave.open(rfilename, **kwargs) as handle:
            raw = handle.readframes(handle.getnframes())
            data = np.frombuffer(raw, dtype="<i{}".format(handle.getsampwidth()))
            channels = handle.getnchannels()
        if len(data) % channels:
            raise IOError("Number of channels do not evenly divide wave samples")
        if channels > 1:
            data = data.reshape((len(data) // channels, channels))
    return _cast(data, dtype)


def _read_npy(rfilename, dtype, key, **kwargs):
    return _cast(np.load(rfilename, **kwargs), dtype)


def _read_npz(rfilename, dtype, key, **kwargs):
    archive = np.load(rfilename, **kwargs)
    return _cast(archive[key if key else "arr_0"], dtype)


def _read_pt(rfilename, dtype, key, **kwargs):
    import torch

    return _cast(torch.load(rfilename, map_location="cpu", **kwargs).numpy(), dtype)


def _read_raw(rfilename, dtype, key, **kwargs):
    if dtype:
        return np.fromfile(rfilename, dtype=dtype, **kwargs)
    return np.fromfile(rfilename, **kwargs)


def _read_hdf5(rfilename, dtype, key, **kwargs):
    import h5py

    with h5py.File(rfilename, "r", **kwargs) as handle:
        node = handle[key] if key else None
        pending = [] if key else [handle]
        while pending:  # depth-first, alphabetical: first dataset wins
            cur = pending.pop()
            if isinstance(cur, h5py.Dataset):
                node = cur
                break
            pending.extend(cur[name] for name in sorted(cur.keys(), reverse=True))
        if node is None:
            raise IOError("Could not find any dataset")
        return np.array(node, dtype=dtype) if dtype else np.array(node)


def _read_soundfile(rfilename, dtype, key, **kwargs):
    import soundfile

    native = {
        "FLOAT": np.float32,
        "DOUBLE": np.float64,
        "PCM_S8": np.int8,
        "PCM_32": np.int32,
        "PCM_24": np.int32,
    }
    with soundfile.SoundFile(rfilename, **kwargs) as handle:
        # two stages (native read, then cast) so floats are not rescaled to +/- 1
        data = handle.read(dtype=native.get(handle.subtype, np.int16))
    return _cast(data, dtype)


def _unavailable(what):
    def reader(rfilename, dtype, key, **kwargs):
        raise IOError(
            f"Reading {what} needs a package that is not part of the B200 build "
            "(pydrobert-kaldi / sph2pipe tables); convert the data to wav/npy/pt first"
        )

    return reader


_READERS = {
    "wav": _read_wav,
    "npy": _read_npy,
    "npz": _read_npz,
    "pt": _read_pt,
    "file": _read_raw,
    "hdf5": _read_hdf5,
    "soundfile": _read_soundfile,
    "table": _unavailable("Kaldi tables"),
    "kaldi": _unavailable("Kaldi objects"),
    "sph": _unavailable("NIST SPHERE files"),
}

_SUFFIXES = (
    (".wav", "wav"),
    (".hdf5", "hdf5"),
    (".npy", "npy"),
    (".npz", "npz"),
    (".pt", "pt"),
    (".sph", "sph"),
    ("|", "kaldi"),
)


def _sniff(rfilename: str) -> str:
    if re.match(r"^(ark|scp)(,\w+)*:", rfilename):
        return "table"
    ext = rfilename.rsplit(".", maxsplit=1)[-1]
    if ext in config.SOUNDFILE_SUPPORTED_FILE_TYPES:
        return ext
    for suffix, kind in _SUFFIXES:
        if rfilename.endswith(suffix):
            return kind
    raise IOError(f"Unable to infer file type from {rfilename}. Set force_as.")


def read_signal(
    rfilename: Union[str, BinaryIO],
    dtype: Optional[np.dtype] = None,
    key: Any = None,
    force_as: Optional[str] = None,
    **kwargs,
) -> np.ndarray:
    """Read an array from a file whose format is sniffed from its name (or ``force_as``)

    Same contract as the reference's ``read_signal``: ``dtype`` casts the result, ``key``
    selects an entry of keyed containers (npz / hdf5), ``force_as`` bypasses sniffing and is
    mandatory for open file objects.
    """
    if not isinstance(rfilename, str):
        if force_as is None:
            raise ValueError("cannot infer type from IO stream. Set force_as")
        if force_as in {"kaldi", "table"}:
            raise ValueError("kaldi types can't be inferred without a string rspecifier")
    elif force_as is None:
        force_as = _sniff(rfilename)
    if force_as in config.SOUNDFILE_SUPPORTED_FILE_TYPES:
        force_as = "soundfile"
    if force_as not in _READERS:
        raise ValueError(
            f"force_as ('{force_as}') is not one of "
            f"{set(_READERS) | config.SOUNDFILE_SUPPORTED_FILE_TYPES}."
        )
    return _READERS[force_as](rfilename, dtype, key, **kwargs)
