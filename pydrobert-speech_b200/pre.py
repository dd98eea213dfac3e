"""Signal pre-processors (reference: ``pydrobert/speech/pre.py:39-149``).

``Dither`` and ``Preemphasize`` keep the reference's constructor, ``coeff`` attribute, aliases
and ``apply(signal, axis=None, in_place=False)`` signature.  In a feature pipeline they are not
run as separate passes at all: :class:`..pipeline.FeaturePipeline` folds them into the sample
staging step of the fused STFT kernel.  Called on their own, ``apply`` runs the stand-alone
CUDA passes ``pds_preemphasize`` / ``pds_dither``.

Dither draws from a counter-based Philox stream keyed by ``(seed, row, sample)`` instead of
NumPy's global Mersenne twister: results are reproducible and independent of batching, but the
noise *values* differ from the reference (which only pins the distribution,
``tests/test_pre.py:32-38``).  The seed comes from :func:`seed` or, failing that, from
``numpy.random`` so that ``numpy.random.seed`` still makes runs repeatable.
"""

import abc
import warnings

from typing import Optional

import numpy as np

from .alias import AliasedFactory

__all__ = ["PreProcessor", "Dither", "Preemphasize"]

_AXIS_DEP_MSG = (
    "Specifying axis in preprocessor.apply is deprecated. "
    "Preprocessors should be applied to 1D signals only."
)


class PreProcessor(AliasedFactory):
    """A transform applied to the raw signal before framing"""

    @abc.abstractmethod
    def apply(
        self, signal: np.ndarray, axis: Optional[int] = None, in_place: bool = False
    ) -> np.ndarray:
        ...


def _rows_on_device(signal: np.ndarray, axis: Optional[int]):
    """View `signal` as independent rows along `axis` (last axis if None) packed for the GPU"""
    import torch

    from ._gpu import current_device

    device = current_device()
    if signal.ndim <= 1:
        rows = signal.reshape(1, -1)
    else:
        rows = np.moveaxis(signal, -1 if axis is None else axis, -1)
    shape = rows.shape
    rows = np.ascontiguousarray(rows.reshape(-1, shape[-1]), dtype=np.float32)
    n_rows, width = rows.shape
    d_in = torch.from_numpy(rows).to(device).reshape(-1)
    offsets = torch.arange(n_rows, dtype=torch.int64, device=device) * width
    lengths = torch.full((n_rows,), width, dtype=torch.int64, device=device)
    return device, d_in, offsets, lengths, shape


def _rows_to_host(d_out, shape, signal, axis, dtype):
    out = d_out.reshape(shape).cpu().numpy()
    if signal.ndim <= 1:
        out = out.reshape(signal.shape)
    else:
        out = np.moveaxis(out, -1, -1 if axis is None else axis)
    return out.astype(dtype, copy=False)


def _launch_rows(entry, d_in, offsets, lengths, device, *args):
    import torch

    from ._gpu import stream_ptr
    from ._lib import check, get_lib

    lib = get_lib()
    d_out = torch.empty_like(d_in)
    n_rows = len(lengths)
    with torch.cuda.device(device):
        for begin in range(0, n_rows, 65535):  # grid.y limit of the row-parallel kernels
            end = min(n_rows, begin + 65535)
            check(
                getattr(lib, entry)(
                    d_in.data_ptr(),
                    d_out.data_ptr(),
                    end - begin,
                    offsets[begin:end].data_ptr(),
                    lengths[begin:end].data_ptr(),
                    int(lengths[begin:end].sum().item()),
                    *args,
                    stream_ptr(device),
                )
            )
    return d_out


class Dither(PreProcessor):
    """Add zero-mean Gaussian noise of standard deviation `coeff` to every sample"""

    aliases = {"dither", "dithering"}

    def __init__(self, coeff: float = 1.0):
        super().__init__()
        self.coeff = coeff

    def apply(
        self, signal: np.ndarray, axis: Optional[int] = None, in_place: bool = False
    ) -> np.ndarray:
        if axis is not None:
            warnings.warn(_AXIS_DEP_MSG, DeprecationWarning)
        signal = np.asarray(signal)
        if signal.size == 0:
            return signal if in_place else signal.copy()
        seed = int(np.random.randint(0, np.iinfo(np.int64).max, dtype=np.int64))
        if axis is None or signal.ndim <= 1:
            # independent noise for every coefficient
            device, d_in, offsets, lengths, shape = _rows_on_device(signal.reshape(-1), None)
            d_out = _launch_rows("pds_dither", d_in, offsets, lengths, device, float(self.coeff), seed)
            out = d_out.cpu().numpy().reshape(signal.shape).astype(signal.dtype, copy=False)
        else:
            # one noise vector along `axis`, shared by all 1-D slices (pre.py:100-103)
            noise_src = np.zeros(signal.shape[axis], dtype=np.float32)
            device, d_in, offsets, lengths, shape = _rows_on_device(noise_src, None)
            noise = _launch_rows("pds_dither", d_in, offsets, lengths, device, float(self.coeff), seed)
            bshape = [1] * signal.ndim
            bshape[axis] = signal.shape[axis]
            out = (signal.astype(np.float64) + noise.cpu().numpy().reshape(bshape)).astype(
                signal.dtype, copy=False
            )
        if in_place:
            signal[...] = out
            return signal
        return out


class Preemphasize(PreProcessor):
    """``new[i] = old[i] - coeff * old[i - 1]``, ``new[0] = old[0]`` along the signal"""

    aliases = {"preemphasize", "preemphasis", "preemph"}

    def __init__(self, coeff: float = 0.97):
        super().__init__()
        self.coeff = coeff

    def apply(
        self, signal: np.ndarray, axis: Optional[int] = None, in_place: bool = False
    ) -> np.ndarray:
        if axis is not None:
            warnings.warn(_AXIS_DEP_MSG, DeprecationWarning)
        signal = np.asarray(signal)
        if signal.size == 0:
            return signal if in_place else signal.copy()
        device, d_in, offsets, lengths, shape = _rows_on_device(signal, axis)
        d_out = _launch_rows("pds_preemphasize", d_in, offsets, lengths, device, float(self.coeff))
        out = _rows_to_host(d_out, shape, signal, axis, signal.dtype)
        if in_place:
            signal[...] = out
            return signal
        return out
