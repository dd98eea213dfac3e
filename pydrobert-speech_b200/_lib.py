"""ctypes binding of ``libpds_b200.so`` (the C ABI declared in ``include/pds_b200.h``).

The shared library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There
is no CPU fallback: if the library is missing, or the machine has no CUDA device, the first
call that needs the device raises -- it never silently computes on the host.
"""

import ctypes
import os

from typing import Optional

__all__ = [
    "PdsError",
    "PdsStftDesc",
    "PdsSiDesc",
    "PdsTile",
    "check",
    "get_lib",
    "lib_path",
    "PDS_F32",
    "PDS_I16",
    "PDS_F64",
]

PDS_F32, PDS_I16, PDS_F64 = 0, 1, 2

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)


class PdsError(RuntimeError):
    """A CUDA-side failure reported through the C ABI (status ``PDS_ERR_CUDA`` and friends)"""


class PdsStftDesc(ctypes.Structure):
    _fields_ = [
        ("frame_length", ctypes.c_int32),
        ("frame_shift", ctypes.c_int32),
        ("dft_size", ctypes.c_int32),
        ("pad_left", ctypes.c_int32),
        ("num_filts", ctypes.c_int32),
        ("include_energy", ctypes.c_int32),
        ("use_power", ctypes.c_int32),
        ("use_log", ctypes.c_int32),
        ("log_floor", ctypes.c_float),
        ("preemph", ctypes.c_float),
        ("dither", ctypes.c_float),
        ("dither_first", ctypes.c_int32),
        ("window", _c_float_p),
        ("band_lo", _c_int32_p),
        ("band_len", _c_int32_p),
        ("band_off", _c_int64_p),
        ("weights", _c_float_p),
    ]


class PdsTile(ctypes.Structure):
    _fields_ = [
        ("sig_off", ctypes.c_int64),
        ("sig_len", ctypes.c_int32),
        ("start", ctypes.c_int32),
        ("nframes", ctypes.c_int32),
        ("utt", ctypes.c_int32),
        ("out_row", ctypes.c_int64),
    ]


class PdsSiDesc(ctypes.Structure):
    _fields_ = [
        ("frame_shift", ctypes.c_int32),
        ("num_filts", ctypes.c_int32),
        ("max_support", ctypes.c_int32),
        ("pad_left", ctypes.c_int32),
        ("frame_start", ctypes.c_int32),
        ("frames_lost", ctypes.c_int32),
        ("use_power", ctypes.c_int32),
        ("use_log", ctypes.c_int32),
        ("is_real", ctypes.c_int32),
        ("log_floor", ctypes.c_float),
        ("h_real", _c_float_p),
        ("h_imag", _c_float_p),
        ("window", _c_float_p),
    ]


assert ctypes.sizeof(PdsTile) == 32

# name -> (restype, argtypes); mirrors include/pds_b200.h one to one
_vp = ctypes.c_void_p
_SIGNATURES = {
    "pds_last_error": (ctypes.c_char_p, []),
    "pds_version": (ctypes.c_int, []),
    "pds_device_count": (ctypes.c_int, []),
    "pds_stft_plan_create": (
        ctypes.c_int,
        [ctypes.POINTER(PdsStftDesc), ctypes.c_int, ctypes.POINTER(_vp)],
    ),
    "pds_stft_plan_destroy": (None, [_vp]),
    "pds_stft_num_coeffs": (ctypes.c_int, [_vp]),
    "pds_stft_tile_frames": (ctypes.c_int, [_vp]),
    "pds_stft_is_fast_path": (ctypes.c_int, [_vp]),
    "pds_stft_kernel_name": (ctypes.c_char_p, [_vp, ctypes.c_int]),
    "pds_stft_num_frames": (ctypes.c_int64, [_vp, ctypes.c_int64]),
    "pds_stft_layout": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, _c_int64_p, _c_int64_p, _c_int64_p],
    ),
    "pds_stft_fill_tiles": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, _c_int64_p, _c_int64_p, _c_int64_p, _vp],
    ),
    "pds_stft_fill_tiles_range": (
        ctypes.c_int,
        [_vp] + [ctypes.c_int64] * 6 + [_vp, _c_int64_p],
    ),
    "pds_stft_run": (
        ctypes.c_int,
        [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_uint64, _vp],
    ),
    "pds_stft_compute_host": (
        ctypes.c_int,
        [
            _vp,
            _vp,
            ctypes.c_int,
            ctypes.c_int64,
            ctypes.c_int64,
            _c_int64_p,
            _c_int64_p,
            _vp,
            ctypes.c_int64,
            _c_int64_p,
            ctypes.c_uint64,
        ],
    ),
    "pds_preemphasize": (
        ctypes.c_int,
        [_vp, _vp, ctypes.c_int64, _vp, _vp, ctypes.c_int64, ctypes.c_float, _vp],
    ),
    "pds_dither": (
        ctypes.c_int,
        [
            _vp,
            _vp,
            ctypes.c_int64,
            _vp,
            _vp,
            ctypes.c_int64,
            ctypes.c_float,
            ctypes.c_uint64,
            _vp,
        ],
    ),
    "pds_deltas": (
        ctypes.c_int,
        [
            _vp,
            _vp,
            ctypes.c_int64,
            ctypes.c_int32,
            ctypes.c_int64,
            _vp,
            ctypes.c_int32,
            _c_float_p,
            _c_int32_p,
            _vp,
        ],
    ),
    "pds_deltas_cmvn_accumulate": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _vp, ctypes.c_int32, _c_float_p, _c_int32_p,
         _vp, _vp],
    ),
    "pds_deltas_cmvn_apply": (
        ctypes.c_int,
        [_vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _vp, ctypes.c_int32, _c_float_p,
         _c_int32_p, _vp, ctypes.c_int32, _vp, _vp],
    ),
    "pds_cmvn_accumulate": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp],
    ),
    "pds_cmvn_apply": (
        ctypes.c_int,
        [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp],
    ),
    "pds_si_plan_create": (
        ctypes.c_int,
        [ctypes.POINTER(PdsSiDesc), ctypes.c_int, ctypes.POINTER(_vp)],
    ),
    "pds_si_plan_destroy": (None, [_vp]),
    "pds_si_num_frames": (ctypes.c_int64, [_vp, ctypes.c_int64]),
    "pds_si_tile_frames": (ctypes.c_int, [_vp]),
    "pds_si_layout": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, _c_int64_p, _c_int64_p, _c_int64_p],
    ),
    "pds_si_fill_tiles": (
        ctypes.c_int,
        [_vp, ctypes.c_int64, _c_int64_p, _c_int64_p, _c_int64_p, _vp],
    ),
    "pds_si_run": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

_LIB: Optional[ctypes.CDLL] = None


def lib_path() -> str:
    """Where the in-tree build puts the shared library"""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpds_b200.so")


def get_lib() -> ctypes.CDLL:
    """Load (once) and return the C-ABI library; raises if it has not been built"""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise PdsError(
                f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C pydrobert-speech_b200/csrc`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = lib
    return _LIB


def check(status: int, what: str = "") -> None:
    """Map a ``pds_status`` onto the exception types the reference's API raises"""
    if status == 0:
        return
    msg = get_lib().pds_last_error().decode("utf-8", "replace")
    if what:
        msg = f"{what}: {msg}"
    if status == -1:
        raise ValueError(msg)
    if status == -3:
        raise NotImplementedError(msg)
    if status == -4:
        raise MemoryError(msg)
    raise PdsError(msg)
