"""Frame computers: features from fixed-length, possibly overlapping, frames of a signal.

The classes, constructor arguments, properties, aliases and error behaviour are those of the
reference (``pydrobert/speech/compute.py``); the arithmetic is not.  Construction builds
float64 tables on the host (:mod:`.filters`, :mod:`._tables`); every ``compute_*`` call then
runs hand-written sm_100a kernels through the C ABI (``include/pds_b200.h``):

* :class:`ShortTimeFourierTransformFrameComputer` (alias ``stft``) -> ``pds_stft_run``: one
  fused kernel doing framing + symmetric padding, optional dither / pre-emphasis, window,
  zero padding, a shared-memory real FFT, ``|X|`` or ``|X|^2``, the banded filter-bank
  contraction, the energy coefficient and the log.
* :class:`ShortIntegrationFrameComputer` (alias ``si``) -> ``pds_si_run``.

Besides the reference's one-signal-at-a-time ``compute_full`` / ``compute_chunk`` /
``finalize``, both computers expose :func:`compute_batch`, which processes a whole packed
batch of utterances in a single launch (the fast path used by the CLI and the benchmark).

There is no CPU implementation in this package.  A CPU restatement used for testing lives in
``oracle/`` and is never imported from here.
"""

import abc
import os

from typing import List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np

from . import config
from ._tables import BandedWeights, fold_filters, stft_pad_left
from .alias import AliasedFactory, alias_factory_subclass_from_arg
from .filters import GammaWindow, HannWindow, LinearFilterBank, WindowFunction

__all__ = [
    "frame_by_frame_calculation",
    "FrameComputer",
    "LinearFilterBankFrameComputer",
    "BatchLayout",
    "PackedSignals",
    "ShortIntegrationFrameComputer",
    "ShortTimeFourierTransformFrameComputer",
    "SIFrameComputer",
    "STFTFrameComputer",
]


class FrameComputer(AliasedFactory):
    """A signal in, a ``(num_frames, num_coeffs)`` feature matrix out

    Interface of the reference's ``FrameComputer`` (``compute.py:48-178``).
    """

    @abc.abstractproperty
    def frame_style(self) -> str:
        """``'causal'`` or ``'centered'``"""

    @abc.abstractproperty
    def sampling_rate(self) -> float:
        ...

    @abc.abstractproperty
    def frame_length(self) -> int:
        ...

    @property
    def frame_length_ms(self) -> float:
        return self.frame_length * 1000 / self.sampling_rate

    @abc.abstractproperty
    def frame_shift(self) -> int:
        ...

    @property
    def frame_shift_ms(self) -> float:
        return self.frame_shift * 1000 / self.sampling_rate

    @abc.abstractproperty
    def num_coeffs(self) -> int:
        ...

    @abc.abstractproperty
    def started(self) -> bool:
        """True between the first :func:`compute_chunk` and the next :func:`finalize`"""

    @abc.abstractmethod
    def compute_chunk(self, chunk: np.ndarray) -> np.ndarray:
        ...

    @abc.abstractmethod
    def finalize(self) -> np.ndarray:
        ...

    def compute_full(self, signal: np.ndarray) -> np.ndarray:
        return frame_by_frame_calculation(self, signal)


class LinearFilterBankFrameComputer(FrameComputer):
    """Computers with one coefficient per filter of a bank (+ optional energy at index 0)

    Reference: ``compute.py:181-218``.
    """

    def __init__(
        self, bank: Union[LinearFilterBank, Mapping, str], include_energy: bool = False
    ):
        self._bank = alias_factory_subclass_from_arg(LinearFilterBank, bank)
        self._include_energy = bool(include_energy)

    @property
    def bank(self) -> LinearFilterBank:
        return self._bank

    @property
    def includes_energy(self) -> bool:
        return self._include_energy

    @property
    def num_coeffs(self) -> int:
        return self._bank.num_filts + int(self._include_energy)


class PackedSignals:
    """A batch of utterances laid out the way the kernels want to read them from HBM

    All signals live in one contiguous buffer.  Utterance ``u`` starts at ``offsets[u]`` and
    each start is placed so that ``offsets[u] - lead`` is a multiple of four samples:
    with ``lead = pad_left % 4`` every 32-frame tile of every utterance then begins on a 16-byte
    boundary and is staged with 128-bit loads.

    Attributes
    ----------
    data : np.ndarray
        1D host buffer (float32, int16 or float64)
    offsets, lengths : np.ndarray
        int64 arrays of length ``len(self)``
    """

    def __init__(self, data: np.ndarray, offsets: np.ndarray, lengths: np.ndarray):
        self.data = data
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)

    def __len__(self) -> int:
        return len(self.lengths)

    @property
    def total_samples(self) -> int:
        return int(self.lengths.sum())

    @staticmethod
    def layout(lengths: Sequence[int], lead: int = 0) -> Tuple[np.ndarray, int]:
        """Start offsets (and total buffer size) for utterances of the given lengths"""
        lengths = np.asarray(lengths, dtype=np.int64)
        padded = (lengths + 3) // 4 * 4
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths):
            offsets[1:] = np.cumsum(padded[:-1])
        offsets += lead
        total = int(padded.sum()) + lead + 4
        return offsets, total

    @classmethod
    def pack(
        cls, signals: Sequence[np.ndarray], dtype=np.float32, lead: int = 0, pin: bool = False
    ) -> "PackedSignals":
        lengths = np.array([len(s) for s in signals], dtype=np.int64)
        offsets, total = cls.layout(lengths, lead)
        if pin:
            import torch

            data = torch.empty(total, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory().numpy()
            data[:] = 0
        else:
            data = np.zeros(total, dtype=dtype)
        for sig, off, n in zip(signals, offsets, lengths):
            data[off : off + n] = sig
        return cls(data, offsets, lengths)


class ShortTimeFourierTransformFrameComputer(LinearFilterBankFrameComputer):
    """Filter-bank features from windowed, DFT'd frames (alias ``stft``)

    Per frame: multiply by the window, take an ``N``-point DFT, and for each filter sum
    ``|X[k] H_f[k]|^p`` over the full spectrum (``p = 2`` if `use_power` else ``1``; doubled for
    real banks), optionally followed by ``log(max(., 1e-5))``.  If `include_energy`, coefficient
    0 is the mean square (or root mean square) of the *un-windowed* frame.

    Parameters are those of the reference (``compute.py:229-362``).  Frame bounds
    (``compute.py:574-600``): the signal is symmetrically padded by ``pad_left`` samples
    (0 if causal, ``L//2 - S//2`` with `kaldi_shift`, else ``(L+1)//2 - 1``), there are
    ``(len + S//2) // S`` frames, and signals shorter than ``L//2 + 1`` produce none.
    """

    aliases = {"stft"}

    def __init__(
        self,
        bank: Union[LinearFilterBank, Mapping, str],
        frame_length_ms: Optional[float] = None,
        frame_shift_ms: Optional[float] = 10,
        frame_style: Optional[str] = None,
        include_energy: bool = False,
        pad_to_nearest_power_of_two: bool = True,
        window_function: Optional[Union[WindowFunction, Mapping, str]] = None,
        use_log: bool = True,
        use_power: bool = False,
        kaldi_shift: bool = False,
    ):
        bank = alias_factory_subclass_from_arg(LinearFilterBank, bank)
        self._rate = bank.sampling_rate
        self._frame_shift = int(0.001 * frame_shift_ms * self._rate)
        self._log = use_log
        self._power = use_power
        self._real = bank.is_real
        self._kaldi_shift = kaldi_shift
        if frame_style is None:
            frame_style = "centered" if bank.is_zero_phase else "causal"
        elif frame_style not in ("centered", "causal"):
            raise ValueError('Invalid frame style: "{}"'.format(frame_style))
        self._frame_style = frame_style
        if frame_length_ms is None:
            # widest temporal support, but at least one DFT bin inside the narrowest band
            widest = max(right - left for left, right in bank.supports)
            narrowest_hz = min(right - left for left, right in bank.supports_hz)
            self._frame_length = max(widest, int(np.ceil(2 * self._rate / narrowest_hz)))
        else:
            self._frame_length = int(0.001 * frame_length_ms * bank.sampling_rate)
        if window_function is None:
            window_function = GammaWindow() if frame_style == "causal" else HannWindow()
        else:
            window_function = alias_factory_subclass_from_arg(WindowFunction, window_function)
        self._window = window_function.get_impulse_response(self._frame_length)
        if pad_to_nearest_power_of_two:
            self._dft_size = int(2 ** np.ceil(np.log2(self._frame_length)))
        else:
            self._dft_size = self._frame_length
        # same private tables as the reference: `from_stft_frame_computer`-style consumers read them
        self._filt_start_idxs: List[int] = []
        self._truncated_filts: List[np.ndarray] = []
        for filt_idx in range(bank.num_filts):
            start_idx, truncated = bank.get_truncated_response(filt_idx, self._dft_size)
            self._filt_start_idxs.append(start_idx)
            self._truncated_filts.append(truncated)
        self._weights: BandedWeights = fold_filters(
            self._filt_start_idxs, self._truncated_filts, self._dft_size, self._power, self._real
        )
        self._plans = dict()  # (device index, preemph, dither, dither_first) -> plan handle
        self._reset_stream()
        super().__init__(bank, include_energy=include_energy)

    @classmethod
    def from_tables(
        cls,
        offsets: Sequence[int],
        truncated_filts: Sequence[np.ndarray],
        frame_length: int,
        frame_shift: int,
        frame_style: str = "centered",
        window: Optional[np.ndarray] = None,
        dft_size: Optional[int] = None,
        use_log: bool = True,
        use_power: bool = False,
        include_energy: bool = False,
        kaldi_shift: bool = False,
        is_real: bool = False,
        sampling_rate: float = 16000,
    ) -> "ShortTimeFourierTransformFrameComputer":
        """Build a computer straight from ``(offset, truncated response)`` tables and geometry in
        samples, without a bank object (what the reference's PyTorch module is constructed from,
        ``torch.py:318-366``)"""
        if frame_style not in ("centered", "causal"):
            raise ValueError('Invalid frame style: "{}"'.format(frame_style))
        self = cls.__new__(cls)
        self._bank = None
        self._num_filts = len(offsets)
        self._include_energy = bool(include_energy)
        self._rate = sampling_rate
        self._frame_shift = int(frame_shift)
        self._frame_length = int(frame_length)
        self._log, self._power, self._real = bool(use_log), bool(use_power), bool(is_real)
        self._kaldi_shift = bool(kaldi_shift)
        self._frame_style = frame_style
        self._window = (
            np.ones(self._frame_length) if window is None else np.asarray(window, dtype=np.float64)
        )
        if dft_size is None:
            dft_size = int(2 ** np.ceil(np.log2(self._frame_length)))
        self._dft_size = int(dft_size)
        self._filt_start_idxs = [int(o) for o in offsets]
        self._truncated_filts = [np.asarray(f) for f in truncated_filts]
        self._weights = fold_filters(
            self._filt_start_idxs, self._truncated_filts, self._dft_size, self._power, self._real
        )
        self._plans = dict()
        self._reset_stream()
        return self

    @property
    def num_coeffs(self) -> int:
        filts = self._bank.num_filts if self._bank is not None else self._num_filts
        return filts + int(self._include_energy)

    # ---- reference properties ---------------------------------------------------------
    @property
    def frame_style(self) -> str:
        return self._frame_style

    @property
    def sampling_rate(self) -> float:
        return self._rate

    @property
    def frame_length(self) -> int:
        return self._frame_length

    @property
    def frame_shift(self) -> int:
        return self._frame_shift

    @property
    def started(self) -> bool:
        return self._started

    @property
    def kaldi_shift(self) -> bool:
        return self._kaldi_shift

    @property
    def dft_size(self) -> int:
        return self._dft_size

    @property
    def pad_left(self) -> int:
        """Samples of symmetric padding in front of frame 0"""
        return stft_pad_left(
            self._frame_length, self._frame_shift, self._frame_style == "centered", self._kaldi_shift
        )

    @property
    def folded_weights(self) -> BandedWeights:
        """The banded ``(num_filts, N/2+1)`` weight matrix the kernel contracts with"""
        return self._weights

    def num_frames(self, sig_len: int) -> int:
        if sig_len < self._frame_length // 2 + 1:
            return 0
        return max(0, (sig_len + self._frame_shift // 2) // self._frame_shift)

    # ---- device side --------------------------------------------------------------------
    def _plan(self, device, preemph: float = 0.0, dither: float = 0.0, dither_first: bool = True):
        import ctypes

        from ._lib import PdsStftDesc, check, get_lib

        # the kernel switch (read by the library once per plan) and the log floor are baked into a
        # plan: both are part of the key, so changing either takes effect on the next call
        key = (device.index, float(preemph), float(dither), bool(dither_first),
               os.environ.get("PDS_STFT_KERNEL", ""), os.environ.get("PDS_STFT_BANK", ""),
               float(config.LOG_FLOOR_VALUE))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        lib = get_lib()
        w = self._weights
        window = np.ascontiguousarray(self._window, dtype=np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        desc = PdsStftDesc(
            frame_length=self._frame_length,
            frame_shift=self._frame_shift,
            dft_size=self._dft_size,
            pad_left=self.pad_left,
            num_filts=len(self._filt_start_idxs),
            include_energy=int(self._include_energy),
            use_power=int(bool(self._power)),
            use_log=int(bool(self._log)),
            log_floor=config.LOG_FLOOR_VALUE,
            preemph=preemph,
            dither=dither,
            dither_first=int(dither_first),
            window=window.ctypes.data_as(fp),
            band_lo=w.lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            band_len=w.length.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            band_off=w.offset.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            weights=(w.taps if len(w.taps) else np.zeros(1, np.float32)).ctypes.data_as(fp),
        )
        handle = ctypes.c_void_p()
        check(lib.pds_stft_plan_create(ctypes.byref(desc), device.index, ctypes.byref(handle)),
              "creating the STFT plan")
        plan = _PlanHandle(handle, lib.pds_stft_plan_destroy)
        self._plans[key] = plan
        return plan

    def kernel_name(self, device=None, dtype=np.float32) -> str:
        """Which kernel ``run_batch`` launches on `device` for samples of `dtype` (benchmark records)"""
        from ._gpu import current_device
        from ._lib import get_lib

        plan = self._plan(current_device() if device is None else device)
        code = 1 if np.dtype(dtype) == np.int16 else 0
        return get_lib().pds_stft_kernel_name(plan.handle, code).decode()

    def plan_batch(self, offsets: np.ndarray, lengths: np.ndarray, device=None, utt_base: int = 0) -> "BatchLayout":
        """Work list for a packed batch: frame offsets on the host, tile table in HBM

        The layout depends only on the utterance offsets / lengths, so it can be built once and
        reused for every batch with the same packing (``utt_base`` is added to the utterance ids
        that key the dither stream, e.g. the global index of the shard's first utterance).
        """
        import ctypes

        import torch

        from ._gpu import TILE_DTYPE, current_device
        from ._lib import check, get_lib

        device = current_device() if device is None else device
        frame_off, tiles = self.plan_tiles(offsets, lengths, device, utt_base)
        d_tiles = torch.from_numpy(tiles.view(np.uint8)).to(device, non_blocking=True)
        return BatchLayout(frame_off, d_tiles, len(tiles), self.num_coeffs, device)

    def plan_tiles(self, offsets: np.ndarray, lengths: np.ndarray, device=None, utt_base: int = 0,
                   row_base: int = 0):
        """Host half of ``plan_batch``: ``(frame_off, tiles)`` with the tile table still on the host
        (structured array of ``pds_tile``); ``row_base`` is added to every tile's output row."""
        import ctypes

        from ._gpu import TILE_DTYPE, current_device
        from ._lib import check, get_lib

        lib = get_lib()
        device = current_device() if device is None else device
        plan = self._plan(device)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        n_utts = len(lengths)
        i64p = ctypes.POINTER(ctypes.c_int64)
        frame_off = np.zeros(n_utts + 1, dtype=np.int64)
        n_tiles = ctypes.c_int64(0)
        check(lib.pds_stft_layout(plan.handle, n_utts, lengths.ctypes.data_as(i64p),
                                  frame_off.ctypes.data_as(i64p), ctypes.byref(n_tiles)))
        tiles = np.empty(n_tiles.value, dtype=TILE_DTYPE)
        if n_tiles.value:
            check(lib.pds_stft_fill_tiles(plan.handle, n_utts, offsets.ctypes.data_as(i64p),
                                          lengths.ctypes.data_as(i64p), frame_off.ctypes.data_as(i64p),
                                          tiles.ctypes.data))
            if utt_base:
                tiles["utt"] += utt_base
            if row_base:
                tiles["out_row"] += row_base
        return frame_off, tiles

    def run_batch(self, layout: "BatchLayout", d_signal, out=None, preemph: float = 0.0,
                  dither: float = 0.0, dither_first: bool = True, seed: int = 0):
        """Enqueue the fused kernel for a planned batch on the current stream; no host sync"""
        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        device = d_signal.device
        if d_signal.dtype == torch.float64:
            d_signal = d_signal.float()  # the kernels compute in float32 anyway
        code = {torch.float32: 0, torch.int16: 1}.get(d_signal.dtype)
        if code is None:
            raise ValueError(f"unsupported sample dtype {d_signal.dtype}")
        plan = self._plan(device, preemph, dither, dither_first)
        if out is None:
            out = torch.empty((layout.rows, self.num_coeffs), dtype=torch.float32, device=device)
        if layout.n_tiles:
            with torch.cuda.device(device):
                check(get_lib().pds_stft_run(plan.handle, d_signal.data_ptr(), code,
                                             layout.d_tiles.data_ptr(), layout.n_tiles, out.data_ptr(),
                                             int(seed) & (2 ** 64 - 1), stream_ptr(device)))
        return out

    def compute_packed_device(self, d_signal, offsets: np.ndarray, lengths: np.ndarray, **pre):
        """Run the fused kernel on a packed batch that is already resident in HBM

        Parameters
        ----------
        d_signal : torch.Tensor
            1D CUDA tensor (float32, int16 or float64) holding all utterances
        offsets, lengths : np.ndarray
            int64 host arrays: start and length of each utterance inside `d_signal`
        **pre
            ``preemph``, ``dither``, ``dither_first``, ``seed``: pre-processing fused into the load

        Returns
        -------
        feats : torch.Tensor
            ``(total_frames, num_coeffs)`` float32 CUDA tensor; utterance ``u`` owns rows
            ``frame_off[u]:frame_off[u+1]``
        frame_off : np.ndarray
            int64, length ``len(lengths) + 1``
        """
        layout = self.plan_batch(offsets, lengths, d_signal.device)
        return self.run_batch(layout, d_signal, **pre), layout.frame_off

    def compute_batch(self, signals: Union[PackedSignals, Sequence[np.ndarray]], **pre) -> List[np.ndarray]:
        """Features of many signals with one launch; returns one float32 array per signal"""
        import torch

        from ._gpu import current_device, dtype_code

        if self.started:
            raise ValueError("Already started computing frames")
        device = current_device()
        if not isinstance(signals, PackedSignals):
            dtypes = {np.asarray(s).dtype for s in signals}
            dtype = dtypes.pop() if len(dtypes) == 1 else np.dtype(np.float32)
            if dtype_code(dtype) is None or dtype == np.float64:
                dtype = np.dtype(np.float32)
            signals = PackedSignals.pack([np.asarray(s) for s in signals], dtype, self.pad_left % 4)
        d_signal = torch.from_numpy(signals.data).to(device)
        feats, frame_off = self.compute_packed_device(d_signal, signals.offsets, signals.lengths, **pre)
        host = feats.cpu().numpy()
        return [host[frame_off[u] : frame_off[u + 1]] for u in range(len(signals))]

    def compute_full(self, signal: np.ndarray) -> np.ndarray:
        if self.started:
            raise ValueError("Already started computing frames")
        signal = np.asarray(signal)
        if self.num_frames(len(signal)) == 0:
            return np.empty((0, self.num_coeffs), dtype=signal.dtype)
        return self._run_host(signal, 0, 0, self.num_frames(len(signal))).astype(signal.dtype, copy=False)

    def _run_host(self, buf: np.ndarray, origin: int, first_frame: int, nframes: int) -> np.ndarray:
        """Frames ``[first_frame, first_frame + nframes)`` of a signal whose samples from
        absolute index ``origin`` on are in ``buf``; float32 result on the host"""
        import ctypes

        import torch

        from ._gpu import TILE_DTYPE, current_device, dtype_code, stream_ptr
        from ._lib import check, get_lib

        lib = get_lib()
        device = current_device()
        plan = self._plan(device)
        if dtype_code(buf.dtype) is None or buf.dtype == np.float64:
            buf = buf.astype(np.float32)
        # place the first tile on a 16-byte boundary so that it takes the vectorised staging path
        lead = (self.pad_left + origin - first_frame * self._frame_shift) % 4
        packed = np.zeros(len(buf) + lead + 4, dtype=buf.dtype)
        packed[lead : lead + len(buf)] = buf
        d_signal = torch.from_numpy(packed).to(device)
        n_tiles = ctypes.c_int64(0)
        lib.pds_stft_fill_tiles_range(plan.handle, lead, len(buf), origin, first_frame, nframes, 0,
                                      None, ctypes.byref(n_tiles))
        tiles = np.empty(n_tiles.value, dtype=TILE_DTYPE)
        check(lib.pds_stft_fill_tiles_range(plan.handle, lead, len(buf), origin, first_frame, nframes,
                                            0, tiles.ctypes.data, ctypes.byref(n_tiles)))
        d_tiles = torch.from_numpy(tiles.view(np.uint8)).to(device)
        feats = torch.empty((nframes, self.num_coeffs), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib.pds_stft_run(plan.handle, d_signal.data_ptr(), dtype_code(packed.dtype),
                                   d_tiles.data_ptr(), n_tiles.value, feats.data_ptr(), 0,
                                   stream_ptr(device)))
        return feats.cpu().numpy()

    # ---- streaming (compute.py:462-572), host-side carry buffer --------------------------
    def _reset_stream(self):
        self._started = False
        self._carry = np.zeros(0, dtype=np.float64)
        self._carry_origin = 0  # absolute sample index of self._carry[0]
        self._seen = 0  # samples received so far
        self._emitted = 0  # frames returned so far
        self._chunk_dtype = np.float64

    def compute_chunk(self, chunk: np.ndarray) -> np.ndarray:
        chunk = np.asarray(chunk)
        self._chunk_dtype = chunk.dtype
        self._started = True
        if len(chunk):
            self._carry = np.concatenate([self._carry, chunk.astype(np.float64, copy=False)])
            self._seen += len(chunk)
        L, S, pad_left = self._frame_length, self._frame_shift, self.pad_left
        # frames whose whole support (bar the left reflection) has arrived ...
        ready = (self._seen + pad_left - L) // S + 1 if self._seen + pad_left >= L else 0
        # ... never more than the final count could be
        ready = min(ready, self.num_frames(self._seen))
        out = self._emit(ready)
        return out

    def _emit(self, upto: int) -> np.ndarray:
        n_new = upto - self._emitted
        if n_new <= 0:
            return np.empty((0, self.num_coeffs), dtype=self._chunk_dtype)
        feats = self._run_host(self._carry, self._carry_origin, self._emitted, n_new)
        self._emitted = upto
        # drop what no later frame (nor the final right-hand reflection) can touch
        next_start = upto * self._frame_shift - self.pad_left
        keep_from = max(0, min(next_start, self._seen - self._frame_length))
        if keep_from > self._carry_origin:
            self._carry = self._carry[keep_from - self._carry_origin :]
            self._carry_origin = keep_from
        return feats.astype(self._chunk_dtype, copy=False)

    def finalize(self) -> np.ndarray:
        out = self._emit(self.num_frames(self._seen)) if self._started else np.empty(
            (0, self.num_coeffs), dtype=self._chunk_dtype
        )
        dtype = self._chunk_dtype
        self._reset_stream()
        self._chunk_dtype = dtype
        return out


STFTFrameComputer = ShortTimeFourierTransformFrameComputer


class ShortIntegrationFrameComputer(LinearFilterBankFrameComputer):
    """Filter, rectify, integrate (alias ``si``)

    Every filter of the bank is convolved with the whole signal, the result is squared
    (`use_power`) or its modulus taken, and ``2 * frame_shift`` samples are pooled with the
    integration window every ``frame_shift`` samples; optionally logged.  All impulse responses
    are clamped to the support of the widest filter (reference ``compute.py:613-752``).

    Geometry, in the terms of the reference's executable spec
    (``tests/test_compute.py:129-176``): with ``translation`` the common delay of the filters,
    the signal is zero-padded on the left by ``max(0, S - translation)`` samples when centered,
    frame ``t`` pools the full convolution over ``[frame_start + t S, frame_start + (t + 2) S)``
    with ``frame_start = max(0, translation - S)`` (centered) or ``translation`` (causal), and
    there are ``(len + S // 2) // S`` frames, minus one for causal banks whose post-translation
    tail is no longer than a frame shift (``compute.py:824-847``).
    """

    aliases = {"si"}

    def __init__(
        self,
        bank: Union[LinearFilterBank, Mapping, str],
        frame_shift_ms: float = 10,
        frame_style: Optional[str] = None,
        include_energy: bool = False,
        pad_to_nearest_power_of_two: bool = True,
        window_function: Optional[Union[WindowFunction, Mapping, str]] = None,
        use_power: bool = False,
        use_log: bool = True,
    ):
        bank = alias_factory_subclass_from_arg(LinearFilterBank, bank)
        self._rate = bank.sampling_rate
        self._frame_shift = int(0.001 * frame_shift_ms * self._rate)
        self._log = bool(use_log)
        self._power = bool(use_power)
        self._real = bank.is_real
        if frame_style is None:
            frame_style = "centered" if bank.is_zero_phase else "causal"
        elif frame_style not in ("centered", "causal"):
            raise ValueError('Invalid frame style: "{}"'.format(frame_style))
        self._frame_style = frame_style
        if window_function is None:
            window_function = GammaWindow() if frame_style == "causal" else HannWindow()
        else:
            window_function = alias_factory_subclass_from_arg(WindowFunction, window_function)
        shift = self._frame_shift
        self._window = window_function.get_impulse_response(2 * shift).reshape(2, shift)
        supports = bank.supports
        if frame_style == "centered":
            # every filter is re-centred on max_support // 2
            self._max_support = max(right - left for left, right in supports)
            self._translation = self._max_support // 2
        else:
            # delay everything by the largest anticipation so that all filters become causal
            self._translation = max([0] + [-left for left, _ in supports])
            self._max_support = max([0] + [right for _, right in supports]) + self._translation
        narrowest_hz = min(right - left for left, right in bank.supports_hz)
        self._frame_length = self._max_support + shift - 1
        self._dft_size = max(self._frame_length, int(np.ceil(2 * self._rate / narrowest_hz)))
        if pad_to_nearest_power_of_two:
            self._dft_size = int(2 ** np.ceil(np.log2(self._dft_size)))
        # impulse responses, delayed and clamped exactly as the reference does before its DFT
        responses = []
        if include_energy:
            dirac = np.zeros(self._max_support, dtype=np.complex128)
            dirac[self._translation] = 1  # a pure delay: "filtered" signal == signal
            responses.append(dirac)
        for filt_idx in range(bank.num_filts):
            response = bank.get_impulse_response(filt_idx, self._dft_size)
            if frame_style == "centered":
                left, right = supports[filt_idx]
                response = np.roll(response, self._translation - (left + right) // 2 + 1)
            else:
                response = np.roll(response, self._translation)
            responses.append(np.asarray(response[: self._max_support], dtype=np.complex128))
        self._impulse_responses = np.array(responses)
        if frame_style == "centered":
            self._zero_pad = max(0, shift - self._translation)
            self._pool_start = max(0, self._translation - shift)
            self._frames_lost = 0
        else:
            self._zero_pad = 0
            self._pool_start = self._translation
            self._frames_lost = int(self._max_support - self._translation <= shift)
        self._plans = dict()
        self._reset_stream()
        super().__init__(bank, include_energy=include_energy)

    @property
    def frame_style(self) -> str:
        return self._frame_style

    @property
    def sampling_rate(self) -> float:
        return self._rate

    @property
    def frame_length(self) -> int:
        return self._frame_length

    @property
    def frame_shift(self) -> int:
        return self._frame_shift

    @property
    def started(self) -> bool:
        return self._started

    def num_frames(self, sig_len: int) -> int:
        return max(0, (sig_len + self._frame_shift // 2) // self._frame_shift - self._frames_lost)

    def _plan(self, device):
        import ctypes

        from ._lib import PdsSiDesc, check, get_lib

        key = (device.index, os.environ.get("PDS_SI_KERNEL", ""), os.environ.get("PDS_SI_PAIRS", ""),
               float(config.LOG_FLOOR_VALUE))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        lib = get_lib()
        fp = ctypes.POINTER(ctypes.c_float)
        h_re = np.ascontiguousarray(self._impulse_responses.real, dtype=np.float32)
        h_im = np.ascontiguousarray(self._impulse_responses.imag, dtype=np.float32)
        window = np.ascontiguousarray(self._window.reshape(-1), dtype=np.float32)
        desc = PdsSiDesc(
            frame_shift=self._frame_shift,
            num_filts=len(self._impulse_responses),
            max_support=self._max_support,
            pad_left=self._zero_pad,
            frame_start=self._pool_start,
            frames_lost=self._frames_lost,
            use_power=int(self._power),
            use_log=int(self._log),
            is_real=int(bool(self._real)),
            log_floor=config.LOG_FLOOR_VALUE,
            h_real=h_re.ctypes.data_as(fp),
            h_imag=h_im.ctypes.data_as(fp),
            window=window.ctypes.data_as(fp),
        )
        handle = ctypes.c_void_p()
        check(lib.pds_si_plan_create(ctypes.byref(desc), device.index, ctypes.byref(handle)),
              "creating the SI plan")
        plan = _PlanHandle(handle, lib.pds_si_plan_destroy)
        self._plans[key] = plan
        return plan

    def compute_packed_device(self, d_signal, offsets: np.ndarray, lengths: np.ndarray):
        """Packed float32 CUDA batch in, ``(total_frames, num_coeffs)`` CUDA tensor + row offsets out"""
        import ctypes

        import torch

        from ._gpu import TILE_DTYPE, stream_ptr
        from ._lib import check, get_lib

        lib = get_lib()
        device = d_signal.device
        if d_signal.dtype != torch.float32:
            d_signal = d_signal.float()
        plan = self._plan(device)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        n_utts = len(lengths)
        i64p = ctypes.POINTER(ctypes.c_int64)
        frame_off = np.zeros(n_utts + 1, dtype=np.int64)
        n_tiles = ctypes.c_int64(0)
        check(lib.pds_si_layout(plan.handle, n_utts, lengths.ctypes.data_as(i64p),
                                frame_off.ctypes.data_as(i64p), ctypes.byref(n_tiles)))
        rows = int(frame_off[-1])
        feats = torch.empty((rows, self.num_coeffs), dtype=torch.float32, device=device)
        if rows == 0:
            return feats, frame_off
        tiles = np.empty(n_tiles.value, dtype=TILE_DTYPE)
        check(lib.pds_si_fill_tiles(plan.handle, n_utts, offsets.ctypes.data_as(i64p),
                                    lengths.ctypes.data_as(i64p), frame_off.ctypes.data_as(i64p),
                                    tiles.ctypes.data))
        return self._launch(plan, d_signal, tiles, feats), frame_off

    def _launch(self, plan, d_signal, tiles, feats):
        import torch

        from ._gpu import stream_ptr
        from ._lib import check, get_lib

        device = d_signal.device
        d_tiles = torch.from_numpy(tiles.view(np.uint8)).to(device, non_blocking=True)
        with torch.cuda.device(device):
            check(get_lib().pds_si_run(plan.handle, d_signal.data_ptr(), d_tiles.data_ptr(), len(tiles),
                                       feats.data_ptr(), stream_ptr(device)))
        d_tiles.record_stream(torch.cuda.current_stream(device))
        return feats

    def compute_batch(self, signals: Union[PackedSignals, Sequence[np.ndarray]]) -> List[np.ndarray]:
        import torch

        from ._gpu import current_device

        if self.started:
            raise ValueError("Already started computing frames")
        if not isinstance(signals, PackedSignals):
            signals = PackedSignals.pack([np.asarray(s) for s in signals], np.float32)
        d_signal = torch.from_numpy(signals.data).to(current_device())
        feats, frame_off = self.compute_packed_device(d_signal, signals.offsets, signals.lengths)
        host = feats.cpu().numpy()
        return [host[frame_off[u] : frame_off[u + 1]] for u in range(len(signals))]

    def compute_full(self, signal: np.ndarray) -> np.ndarray:
        if self._started:
            raise ValueError("Already started computing frames")
        signal = np.asarray(signal)
        if not np.issubdtype(signal.dtype, np.floating):
            raise ValueError("Chunk must be a float type")
        if self.num_frames(len(signal)) == 0:
            return np.empty((0, self.num_coeffs), dtype=signal.dtype)
        return self._run_host(signal, 0, 0, self.num_frames(len(signal))).astype(signal.dtype, copy=False)

    def _run_host(self, buf: np.ndarray, origin: int, first_frame: int, nframes: int) -> np.ndarray:
        import torch

        from ._gpu import TILE_DTYPE, current_device

        device = current_device()
        plan = self._plan(device)
        d_signal = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.float32)).to(device)
        from ._lib import get_lib

        tile_frames = int(get_lib().pds_si_tile_frames(plan.handle))
        starts = np.arange(first_frame, first_frame + nframes, tile_frames)
        tiles = np.zeros(len(starts), dtype=TILE_DTYPE)
        tiles["sig_off"] = 0
        tiles["sig_len"] = len(buf)
        # the kernel indexes the convolution of the left-padded *buffer*: shift by what was dropped
        tiles["start"] = self._pool_start + starts * self._frame_shift - origin
        tiles["nframes"] = np.minimum(tile_frames, first_frame + nframes - starts)
        tiles["out_row"] = starts - first_frame
        feats = torch.empty((nframes, self.num_coeffs), dtype=torch.float32, device=device)
        return self._launch(plan, d_signal, tiles, feats).cpu().numpy()

    # ---- streaming -----------------------------------------------------------------------
    def _reset_stream(self):
        self._started = False
        self._carry = np.zeros(0, dtype=np.float64)
        self._carry_origin = 0
        self._seen = 0
        self._emitted = 0
        self._ret_dtype = np.float64

    def compute_chunk(self, chunk: np.ndarray) -> np.ndarray:
        chunk = np.asarray(chunk)
        if self._started:
            if chunk.dtype != self._ret_dtype:
                raise ValueError("Chunk does not share a type with previous chunks")
        else:
            if not np.issubdtype(chunk.dtype, np.floating):
                raise ValueError("Chunk must be a float type")
            self._ret_dtype = chunk.dtype
            self._started = True
        if len(chunk):
            self._carry = np.concatenate([self._carry, chunk.astype(np.float64, copy=False)])
            self._seen += len(chunk)
        # frame t pools convolution samples < pool_start + (t + 2) S, which depend on signal
        # samples < that - zero_pad: ready once they have all arrived
        shift = self._frame_shift
        ready = (self._seen + self._zero_pad - self._pool_start) // shift - 1
        ready = max(0, min(ready, self.num_frames(self._seen)))
        return self._emit(ready)

    def _emit(self, upto: int) -> np.ndarray:
        n_new = upto - self._emitted
        if n_new <= 0:
            return np.empty((0, self.num_coeffs), dtype=self._ret_dtype)
        # dropping `origin` leading samples only shifts the convolution index (see _run_host)
        feats = self._run_host(self._carry, self._carry_origin, self._emitted, n_new)
        self._emitted = upto
        # oldest signal sample the next frame's pooling region still depends on
        next_y = self._pool_start + upto * self._frame_shift
        keep_from = max(0, next_y - (self._max_support - 1) - self._zero_pad)
        if keep_from > self._carry_origin:
            self._carry = self._carry[keep_from - self._carry_origin :]
            self._carry_origin = keep_from
        return feats.astype(self._ret_dtype, copy=False)

    def finalize(self) -> np.ndarray:
        out = self._emit(self.num_frames(self._seen)) if self._started else np.empty(
            (0, self.num_coeffs), dtype=self._ret_dtype
        )
        dtype = self._ret_dtype
        self._reset_stream()
        self._ret_dtype = dtype
        return out


SIFrameComputer = ShortIntegrationFrameComputer


class BatchLayout:
    """Result of ``plan_batch``: where each utterance's frames go and the device tile table"""

    def __init__(self, frame_off, d_tiles, n_tiles, num_coeffs, device):
        self.frame_off = frame_off
        self.d_tiles = d_tiles
        self.n_tiles = n_tiles
        self.num_coeffs = num_coeffs
        self.device = device

    @property
    def rows(self) -> int:
        return int(self.frame_off[-1])


class _PlanHandle:
    """Owns a C-ABI plan; destroys it with the computer"""

    def __init__(self, handle, destroy):
        self.handle = handle
        self._destroy = destroy

    def __del__(self):
        try:
            if self.handle:
                self._destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass


def frame_by_frame_calculation(computer: FrameComputer, signal: np.ndarray, chunk_size: int = 2 ** 10):
    """Features of `signal` through repeated ``compute_chunk`` calls (``compute.py:1002-1039``)"""
    if computer.started:
        raise ValueError("Already started computing frames")
    pieces = []
    for begin in range(0, len(signal), chunk_size):
        pieces.append(computer.compute_chunk(signal[begin : begin + chunk_size]))
    pieces.append(computer.finalize())
    return np.concatenate(pieces)
