"""Feature post-processors (reference: ``pydrobert/speech/post.py:38-563``).

* :class:`Standardize` (aliases ``standardize``, ``normalize``, ``unit``, ``cmvn``): sufficient
  statistics ``[[sum x..., count], [sum x^2..., 0]]`` in float64 (Kaldi CMVN layout) are
  accumulated on the GPU (``pds_cmvn_accumulate``), summed across ranks with one
  ``all_reduce`` (:func:`Standardize.allreduce`), and applied with ``pds_cmvn_apply``.
* :class:`Deltas`: clamp-edge FIR along the filtered axis (``pds_deltas``).
* :class:`Stack`: a pure re-indexing, done with array views on the host.

The NumPy-in / NumPy-out ``apply`` signatures are the reference's; ``*_device`` variants take
and return CUDA tensors so that a pipeline never leaves HBM.
"""

import abc
import warnings

from itertools import count
from typing import Callable, Optional, Union

import numpy as np

from .alias import AliasedFactory
from .util import read_signal

__all__ = ["PostProcessor", "Standardize", "CMVN", "Deltas", "Stack"]


class PostProcessor(AliasedFactory):
    """A transform applied to the feature tensor"""

    @abc.abstractmethod
    def apply(
        self, features: np.ndarray, axis: int = -1, in_place: bool = False
    ) -> np.ndarray:
        ...


def _as_rows(tensor: np.ndarray, axis: int):
    """(rows, C) float32 contiguous view of `tensor` with `axis` last, plus the moved shape"""
    moved = np.moveaxis(tensor, axis, -1)
    rows = np.ascontiguousarray(moved.reshape(-1, moved.shape[-1]), dtype=np.float32)
    return rows, moved.shape


class Standardize(PostProcessor):
    """Zero mean (and unit variance if `norm_var`) per coefficient

    With no statistics (``rfilename`` unset and nothing accumulated) each call to :func:`apply`
    standardises the tensor against itself.  The return type of :func:`apply` is float64, as in
    the reference (``post.py:66-364``).
    """

    aliases = {"standardize", "normalize", "unit", "cmvn"}

    def __init__(self, rfilename: Optional[str] = None, norm_var: bool = True, **kwargs):
        # The statistics live on the GPU while they are being accumulated / reduced / applied
        # (`_d_stats`, float64 (2, C+1)); `_h_stats` is the host copy, refreshed on demand.  Exactly
        # one of the two is authoritative at any time (`_host_fresh`).
        self._h_stats = None
        self._d_stats = None
        self._d_flag = None      # int32 device flag: a ~0 variance was replaced by 1
        self._host_fresh = True
        self._rows_seen = 0      # rows accumulated through this object (host-side bookkeeping, no sync)
        self._norm_var = bool(norm_var)
        if rfilename is not None:
            if "dtype" in kwargs:
                self._h_stats = read_signal(rfilename, **kwargs)
            else:
                for dtype in (np.float64, np.float32, "dm", "fm"):
                    try:
                        self._h_stats = read_signal(rfilename, dtype=dtype, **kwargs)
                        break
                    except (IOError, ValueError, ImportError, TypeError):
                        pass
                if self._h_stats is None:
                    raise IOError("Unable to load stats from {}".format(rfilename))
                if self._h_stats.ndim == 1:
                    self._h_stats = self._sanitized(self._h_stats)
        elif kwargs:
            raise TypeError("Invalid keyword arguments: {}".format(tuple(kwargs)))
        super().__init__()

    @staticmethod
    def _sanitized(flat: np.ndarray) -> np.ndarray:
        """Recover (2, F+1) float64 stats from a raw binary that may have been float32"""

        def plausible(stats):
            try:
                stats = stats.reshape((2, -1))
            except ValueError:
                return None
            ok = np.isclose(np.round(stats[0, -1]), stats[0, -1]) and np.all(stats >= 0)
            return stats if ok else None

        first = plausible(flat)
        if first is not None:
            return first
        if flat.dtype not in (np.float32, np.float64):
            raise ValueError(
                "Statistics were loaded with a weird data type ({}) and are invalid. Make sure "
                "the arguments you passed to the init are correct".format(flat.dtype)
            )
        other = np.float64 if flat.dtype == np.float32 else np.float32
        second = plausible(np.frombuffer(flat.tobytes(), dtype=other).astype(np.float64))
        if second is None:
            raise IOError(
                "Could not properly load statistics. Try specifying additional parameters in "
                "init (see docstring)"
            )
        return second

    # ---- where the statistics live -------------------------------------------------------
    @property
    def _stats(self) -> Optional[np.ndarray]:
        """Host view of the statistics; copies them back from the GPU (one sync) if they moved on"""
        if not self._host_fresh:
            self._h_stats = self._d_stats.cpu().numpy()
            self._host_fresh = True
        return self._h_stats

    @_stats.setter
    def _stats(self, value) -> None:
        self._h_stats = value
        self._d_stats = None
        self._host_fresh = True

    def _width(self) -> Optional[int]:
        if self._d_stats is not None:
            return self._d_stats.shape[1]
        return None if self._h_stats is None else self._h_stats.shape[1]

    def device_stats(self, device, num_coeffs: Optional[int] = None):
        """The float64 ``(2, C + 1)`` CUDA tensor the kernels accumulate into and read from

        Created on first use (zeros, or a copy of statistics loaded from a file); afterwards every
        ``*_device`` call and :func:`allreduce` works on it in stream order, without host syncs.
        """
        import torch

        if self._d_stats is not None and self._d_stats.device != device:
            self._stats = self._stats  # pull back, forget the copy on the other device
        if self._d_stats is None:
            if self._h_stats is not None:
                self._d_stats = torch.from_numpy(np.ascontiguousarray(self._h_stats, np.float64)).to(device)
            else:
                if num_coeffs is None:
                    raise ValueError("No stats have been accumulated")
                self._d_stats = torch.zeros((2, num_coeffs + 1), dtype=torch.float64, device=device)
            self._d_flag = torch.zeros(1, dtype=torch.int32, device=device)
        return self._d_stats

    @property
    def have_stats(self) -> bool:
        if self._rows_seen:
            return True
        return bool(self._h_stats is not None and self._h_stats[0, -1])

    @property
    def stats(self) -> Optional[np.ndarray]:
        """The ``(2, num_coeffs + 1)`` float64 statistics accumulated so far (or None)"""
        return self._stats

    def _check_width(self, num_coeffs: int) -> None:
        width = self._width()
        if width is not None and width != num_coeffs + 1:
            raise ValueError(
                "Expected feature vector of length {}; got {}".format(width - 1, num_coeffs)
            )

    def check_zero_variance(self) -> bool:
        """Raise the reference's "0 variance" warning if a device-side apply replaced a variance by
        1 since the last check.  The ``*_device`` methods never read the flag themselves (that would
        be a host sync per call); the NumPy ``apply`` and the batch pipelines call this at the end."""
        if self._d_flag is None:
            return False
        hit = bool(int(self._d_flag.item()))
        if hit:
            self._d_flag.zero_()
            warnings.warn("0 variance encountered. Replacing with 1")
        return hit

    # ---- device-resident variants ------------------------------------------------------
    def accumulate_device(self, feats) -> None:
        """Add the rows of a ``(rows, C)`` float32 CUDA tensor to the statistics (no host sync)"""
        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        rows, cols = feats.shape
        if rows == 0:
            raise ValueError("Cannot accumulate from empty array")
        self._check_width(cols)
        d_stats = self.device_stats(feats.device, cols)
        if isinstance(feats, LazyDeltas):
            # Deltas -> Standardize: statistics straight from the static features, the deltas are
            # recomputed on the fly instead of being written and read back
            if not feats.fused_call("pds_deltas_cmvn_accumulate", d_stats.data_ptr()):
                feats = feats.materialize()
        if not isinstance(feats, LazyDeltas):
            feats = feats.contiguous()
            with torch.cuda.device(feats.device):
                check(get_lib().pds_cmvn_accumulate(feats.data_ptr(), rows, cols, d_stats.data_ptr(),
                                                    stream_ptr(feats.device)))
        self._host_fresh = False
        self._rows_seen += rows

    def apply_device(self, feats, out=None):
        """Standardise a ``(rows, C)`` float32 CUDA tensor with the accumulated statistics

        Stream-ordered after whatever produced the statistics (``accumulate_device``,
        :func:`allreduce`); the zero-variance flag is left on the device, see
        :func:`check_zero_variance`."""
        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        rows, cols = feats.shape
        self._check_width(cols)
        if not self.have_stats:
            raise ValueError("No stats have been accumulated")
        d_stats = self.device_stats(feats.device, cols)
        if isinstance(feats, LazyDeltas):
            if out is None:
                out = torch.empty((rows, cols), dtype=torch.float32, device=feats.device)
            if feats.fused_call("pds_deltas_cmvn_apply", d_stats.data_ptr(), int(self._norm_var),
                                self._d_flag.data_ptr(), out=out):
                return out
            feats = feats.materialize()
        feats = feats.contiguous()
        out = torch.empty_like(feats) if out is None else out
        with torch.cuda.device(feats.device):
            check(get_lib().pds_cmvn_apply(feats.data_ptr(), out.data_ptr(), rows, cols,
                                           d_stats.data_ptr(), int(self._norm_var), self._d_flag.data_ptr(),
                                           stream_ptr(feats.device)))
        return out

    def allreduce(self, group=None) -> None:
        """Sum the statistics over all ranks of a ``torch.distributed`` process group

        The only collective of the whole feature pipeline: ``2 * (C + 1)`` doubles.  With the
        NCCL backend the device-resident statistics are reduced in place, ordered after the
        accumulation kernels and before the apply kernels on the current stream (no host
        round trip); with gloo (CPU tests) the host copy is reduced.
        """
        import torch
        import torch.distributed as dist

        if self._width() is None:
            raise ValueError("No stats have been accumulated to reduce")
        if dist.get_backend(group) == "nccl":
            if self._d_stats is None:
                self.device_stats(torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(self._d_stats, op=dist.ReduceOp.SUM, group=group)
            self._host_fresh = False
        else:
            buf = torch.from_numpy(np.ascontiguousarray(self._stats))
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            self._stats = buf.numpy()
        self._rows_seen = max(self._rows_seen, 1)

    # ---- reference API -------------------------------------------------------------------
    def accumulate(self, features: np.ndarray, axis: int = -1) -> None:
        import torch

        from ._gpu import current_device

        features = np.asarray(features)
        if (features.shape and not np.prod(features.shape)) or not len(features):
            raise ValueError("Cannot accumulate from empty array")
        if features.ndim > 1:
            rows, _ = _as_rows(features, axis)
        else:
            rows = np.ascontiguousarray(features.reshape(1, -1), dtype=np.float32)
        self._check_width(rows.shape[1])
        self.accumulate_device(torch.from_numpy(rows).to(current_device()))

    def apply(self, features: np.ndarray, axis: int = -1, in_place: bool = False) -> np.ndarray:
        """``(x - mean) / std`` per coefficient along `axis`, returned as float64

        The arithmetic runs in float32 on the GPU (the statistics are float64).  With
        ``in_place=True`` and a float64 `features` the result is also written back into the
        argument, as in the reference (``post.py:263-264``)."""
        import torch

        from ._gpu import current_device

        features = np.asarray(features)
        if (features.shape and not np.prod(features.shape)) or not len(features):
            raise ValueError("Cannot apply to empty array")
        is_vector = features.ndim <= 1
        if is_vector:
            rows, moved_shape = np.ascontiguousarray(features.reshape(1, -1), np.float32), None
        else:
            rows, moved_shape = _as_rows(features, axis)
        self._check_width(rows.shape[1])
        single = rows.shape[0] == 1
        if not self.have_stats and single:
            if self._norm_var:
                raise ValueError(
                    "Unable to standardize the variance of a vector with no global statistics"
                )
            warnings.warn("Standardizing a single vector to 0")
            if in_place and features.dtype == np.float64:
                features[...] = 0
                return features
            return np.zeros(features.shape, dtype=np.float64)
        d_rows = torch.from_numpy(rows).to(current_device())
        if self.have_stats:
            d_out = self.apply_device(d_rows)
            self.check_zero_variance()
        else:  # local statistics: accumulate on this tensor only, then forget them
            local = Standardize(norm_var=self._norm_var)
            local.accumulate_device(d_rows)
            d_out = local.apply_device(d_rows)
            local.check_zero_variance()
        out = d_out.cpu().numpy().astype(np.float64)
        if is_vector:
            out = out.reshape(features.shape)
        else:
            out = np.moveaxis(out.reshape(moved_shape), -1, axis)
        if in_place and features.dtype == np.float64:
            features[...] = out
            return features
        return out

    def save(self, wfilename: str, key: Optional[str] = None, compress: bool = False,
             overwrite: bool = True) -> None:
        """Write the statistics (``.npy``, ``.npz`` or raw float64), reference ``post.py:307-361``"""
        if not self.have_stats:
            raise ValueError("No stats have been accumulated to save")
        if wfilename.endswith(".npy"):
            np.save(wfilename, self._stats)
        elif wfilename.endswith(".npz"):
            array = dict()
            if overwrite:
                try:
                    array = dict(np.load(wfilename))
                except IOError:
                    pass
            if key is None:
                key = next(k for k in ("arr_{}".format(v) for v in count(0)) if k not in array)
            array[key] = self._stats
            (np.savez_compressed if compress else np.savez)(wfilename, **array)
        else:
            self._stats.tofile(wfilename)


CMVN = Standardize


class Deltas(PostProcessor):
    """Append (or stack) ``num_deltas`` orders of delta features

    Order ``i + 1`` correlates the features along `axis` with the ``i``-fold self-convolution of
    the ramp ``[-W..W] / sum j^2`` (``W = context_window``), the sequence being edge-padded
    (reference ``post.py:367-491``).
    """

    aliases = {"deltas"}

    def __init__(self, num_deltas: int, target_axis: int = -1, concatenate: bool = True,
                 context_window: int = 2, pad_mode: Union[str, Callable] = "edge", **kwargs):
        self._target_axis = target_axis
        self._pad_mode = pad_mode
        self._pad_kwargs = kwargs
        self.concatenate = bool(concatenate)
        self.num_deltas = num_deltas
        ramp = np.arange(-context_window, context_window + 1, dtype=np.float64)
        ramp /= np.sum(ramp ** 2)
        self._filts = [np.ones(1, dtype=np.float64)]
        for _ in range(num_deltas):
            self._filts.append(np.convolve(self._filts[-1], ramp))

    def lazy_device(self, feats, row_off=None) -> "LazyDeltas":
        """Like :meth:`apply_device`, but the output is only described, not computed: passing the
        result to ``Standardize.accumulate_device`` / ``apply_device`` runs the fused
        Deltas + CMVN kernels (``pds_deltas_cmvn_*``), which never write the un-normalised deltas.
        ``materialize()`` gives the plain tensor."""
        return LazyDeltas(self, feats, row_off)

    def apply_device(self, feats, row_off=None, out=None):
        """``(rows, C)`` float32 CUDA tensor -> ``(rows, C * (num_deltas + 1))``

        `row_off` (int64 CUDA tensor, ``n_utts + 1``) delimits utterances packed along the rows;
        the filter never reaches across a boundary.  Default: one utterance.  `out`: a contiguous
        float32 CUDA tensor of the result's shape to write into instead of a new one.
        """
        import ctypes

        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        rows, cols = feats.shape
        feats = feats.contiguous()
        if row_off is None:
            row_off = torch.tensor([0, rows], dtype=torch.int64, device=feats.device)
        shape = (rows, cols * (self.num_deltas + 1))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=feats.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {shape}")
        taps = np.concatenate(self._filts[1:] + [np.zeros(0)]).astype(np.float32)
        lens = np.array([len(f) for f in self._filts[1:]], dtype=np.int32)
        with torch.cuda.device(feats.device):
            check(get_lib().pds_deltas(
                feats.data_ptr(), out.data_ptr(), rows, cols, len(row_off) - 1, row_off.data_ptr(),
                self.num_deltas, taps.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                lens.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), stream_ptr(feats.device)))
        return out

    def apply(self, features: np.ndarray, axis: int = -1, in_place: bool = False) -> np.ndarray:
        import torch

        from ._gpu import current_device

        features = np.asarray(features)
        if self._pad_mode != "edge" or self._pad_kwargs:
            raise NotImplementedError("the CUDA delta kernel implements pad_mode='edge' only")
        if features.size == 0 or self.num_deltas == 0:
            pieces = [features] * (self.num_deltas + 1)
        else:
            # filtered axis first, everything else flattened into columns
            moved = np.moveaxis(features, axis, 0)
            flat = np.ascontiguousarray(moved.reshape(moved.shape[0], -1), dtype=np.float32)
            d_out = self.apply_device(torch.from_numpy(flat).to(current_device()))
            out = d_out.cpu().numpy().reshape((moved.shape[0], self.num_deltas + 1) + moved.shape[1:])
            pieces = [features] + [
                np.moveaxis(out[:, k], 0, axis).astype(features.dtype, copy=False)
                for k in range(1, self.num_deltas + 1)
            ]
        if self.concatenate:
            return np.concatenate(pieces, self._target_axis)
        return np.stack(pieces, self._target_axis)


class Stack(PostProcessor):
    """Group `num_vectors` consecutive frames into one (reference ``post.py:494-563``)

    No arithmetic is involved: the result is a re-indexing of the input.
    """

    aliases = {"stack"}

    def __init__(self, num_vectors: int, time_axis: int = 0,
                 pad_mode: Optional[Union[str, Callable]] = None, **kwargs) -> None:
        if num_vectors < 1:
            raise ValueError(f"Expected num_vectors to be positive, got {num_vectors}")
        self.num_vectors = num_vectors
        self.time_axis = time_axis
        self._pad_mode = pad_mode
        self._pad_kwargs = kwargs

    def apply(self, features: np.ndarray, axis: int = -1, in_place: bool = False) -> np.ndarray:
        features = np.asarray(features)
        axis %= features.ndim
        time_axis = self.time_axis % features.ndim
        if axis == time_axis:
            raise RuntimeError(f"feature and time axes are the same ({axis})")
        num_frames = features.shape[time_axis]
        if self._pad_mode is not None and num_frames % self.num_vectors:
            padding = [(0, 0)] * features.ndim
            padding[time_axis] = (0, self.num_vectors - num_frames % self.num_vectors)
            features = np.pad(features, padding, self._pad_mode, **self._pad_kwargs)
            num_frames = features.shape[time_axis]
        kept = num_frames // self.num_vectors * self.num_vectors
        index = [slice(None)] * features.ndim
        groups = []
        for i in range(self.num_vectors):
            index[time_axis] = slice(i, kept, self.num_vectors)
            groups.append(features[tuple(index)])
        return np.concatenate(groups, axis)


class LazyDeltas:
    """The not-yet-computed output of ``Deltas.apply_device(feats, row_off)`` (see
    :meth:`Deltas.lazy_device`)"""

    def __init__(self, deltas: "Deltas", feats, row_off=None):
        import torch

        self.deltas = deltas
        self.feats = feats.contiguous()
        rows = feats.shape[0]
        self.row_off = (torch.tensor([0, rows], dtype=torch.int64, device=feats.device)
                        if row_off is None else row_off)
        self.device = feats.device
        self.shape = (rows, feats.shape[1] * (deltas.num_deltas + 1))

    def materialize(self):
        return self.deltas.apply_device(self.feats, self.row_off)

    def fused_call(self, name: str, *tail, out=None) -> bool:
        """Run ``pds_deltas_cmvn_accumulate`` / ``_apply``; False if the filters are not covered"""
        import ctypes

        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        d = self.deltas
        taps = np.concatenate(d._filts[1:] + [np.zeros(0)]).astype(np.float32)
        lens = np.array([len(f) for f in d._filts[1:]], dtype=np.int32)
        rows, cols = self.feats.shape
        head = [self.feats.data_ptr()] + ([out.data_ptr()] if out is not None else [])
        args = head + [rows, cols, len(self.row_off) - 1, self.row_off.data_ptr(), d.num_deltas,
                       taps.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                       lens.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))] + list(tail) + [stream_ptr(self.device)]
        with torch.cuda.device(self.device):
            rc = getattr(get_lib(), name)(*args)
        if rc == -3:  # PDS_ERR_UNSUPPORTED: other filter lengths
            return False
        check(rc, name)
        return True
