"""Device plumbing: PyTorch is used for device memory, streams and (elsewhere) NCCL only."""

import ctypes
import os

from typing import Optional

import numpy as np
import torch

from ._lib import PDS_F32, PDS_F64, PDS_I16, PdsError, PdsTile, check, get_lib

__all__ = [
    "current_device",
    "dtype_code",
    "require_cuda",
    "stream_ptr",
    "to_device",
    "TILE_DTYPE",
]

# numpy view of struct pds_tile (include/pds_b200.h)
TILE_DTYPE = np.dtype(
    [
        ("sig_off", "<i8"),
        ("sig_len", "<i4"),
        ("start", "<i4"),
        ("nframes", "<i4"),
        ("utt", "<i4"),
        ("out_row", "<i8"),
    ]
)
assert TILE_DTYPE.itemsize == ctypes.sizeof(PdsTile)


def require_cuda() -> None:
    """Fail loudly when there is nothing to run the kernels on (no CPU fallback exists)"""
    if not torch.cuda.is_available():
        raise PdsError(
            "pydrobert-speech_b200 needs a CUDA device (B200, sm_100a); none is visible and "
            "there is no CPU fallback"
        )
    get_lib()


def current_device() -> torch.device:
    require_cuda()
    idx = os.environ.get("PDS_DEVICE")
    if idx is not None:
        return torch.device("cuda", int(idx))
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_NP_TO_CODE = {
    np.dtype(np.float32): PDS_F32,
    np.dtype(np.int16): PDS_I16,
    np.dtype(np.float64): PDS_F64,
}


def dtype_code(dtype) -> Optional[int]:
    return _NP_TO_CODE.get(np.dtype(dtype))


def to_device(array: np.ndarray, device: torch.device, pin: bool = False) -> torch.Tensor:
    """Copy a contiguous host array to ``device`` (as a flat byte-compatible tensor)"""
    host = torch.from_numpy(np.ascontiguousarray(array))
    if pin:
        host = host.pin_memory()
    return host.to(device, non_blocking=pin)
