"""PyTorch front-end: ``torch.nn.Module`` wrappers with the reference's names and constructors.

Reference: ``pydrobert/speech/torch.py:73-522``.  There, the STFT module re-implements the frame
computer with ``as_strided`` + ``torch.fft.rfft`` + a Python loop over filters, and the SI module
and the post-processors round-trip through NumPy on the CPU.  Here every ``forward`` enqueues the
same hand-written CUDA kernels as the NumPy-facing classes; tensors that already live on a CUDA
device never leave it, CPU tensors are copied over and the result copied back (so the modules can
stand in for the reference's inside ``signals-to-torch-feat-dir``).

Differences worth knowing:

* the forward pass of the STFT module is the fused CUDA kernel; its window and filters are learnable
  parameters like the reference's (``torch.py:362-366``) and gradients with respect to them and to the
  signal are produced in the backward pass by a plain-torch restatement of the same arithmetic
  (:func:`_stft_math`), so training code written against the reference runs unchanged.  The other
  modules (pre-processing, short integration, post-processors) are inference only;
* complex banks follow the *NumPy* path of the reference, which is what ``BASELINE.json`` names as
  the oracle; the reference's own torch path disagrees with its NumPy path there
  (``torch.py:213-214`` vs ``compute.py:436-438``);
* arithmetic is float32 whatever the input dtype; the output takes the input's floating dtype.
"""

import math

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import config
from ._tables import _segment_bins, half_spectrum_len, stft_geometry
from .compute import SIFrameComputer, STFTFrameComputer
from .post import PostProcessor
from .pre import Dither, Preemphasize

__all__ = [
    "pytorch_dither",
    "pytorch_preemphasize",
    "pytorch_stft_frame_computer",
    "PyTorchDither",
    "PyTorchPostProcessorWrapper",
    "PyTorchPreemphasize",
    "PyTorchShortIntegrationFrameComputer",
    "PyTorchShortTimeFourierTransformFrameComputer",
    "PyTorchSIFrameComputer",
    "PyTorchSTFTFrameComputer",
]


def _check_positive(name: str, val, nonnegative=False):
    kind = "non-negative" if nonnegative else "positive"
    if val < 0 or (val == 0 and not nonnegative):
        raise ValueError(f"Expected {name} to be {kind}; got {val}")


def _on_device(sig: torch.Tensor) -> Tuple[torch.Tensor, torch.device]:
    """The signal as a contiguous CUDA tensor plus where the result should go back to"""
    from ._gpu import current_device

    home = sig.device
    if sig.ndim != 1:
        raise RuntimeError(f"Expected x to be 1-dimensional; got {sig.ndim}")
    if home.type != "cuda":
        sig = sig.to(current_device())
    return sig.contiguous(), home


def _result(feats: torch.Tensor, sig_dtype: torch.dtype, home: torch.device) -> torch.Tensor:
    if sig_dtype.is_floating_point and sig_dtype != torch.float32:
        feats = feats.to(sig_dtype)
    return feats.to(home)


def _row_launch(entry: str, sig: torch.Tensor, *args) -> torch.Tensor:
    """Run one of the stand-alone pre-processing kernels over a single 1-D CUDA signal"""
    from ._gpu import stream_ptr
    from ._lib import check, get_lib

    src = sig.float()
    dst = torch.empty_like(src)
    meta = torch.tensor([0, src.numel()], dtype=torch.int64, device=src.device)
    with torch.cuda.device(src.device):
        check(getattr(get_lib(), entry)(src.data_ptr(), dst.data_ptr(), 1, meta[:1].data_ptr(),
                                        meta[1:].data_ptr(), src.numel(), *args, stream_ptr(src.device)))
    return dst


def pytorch_preemphasize(sig: torch.Tensor, coeff: float = 0.97) -> torch.Tensor:
    """``out[0] = sig[0]``, ``out[i] = sig[i] - coeff * sig[i - 1]`` (kernel ``pds_preemphasize``)"""
    if sig.numel() == 0:
        return sig.clone()
    d_sig, home = _on_device(sig)
    return _result(_row_launch("pds_preemphasize", d_sig, float(coeff)), sig.dtype, home)


class PyTorchPreemphasize(torch.nn.Module):
    """Module form of :func:`pytorch_preemphasize` (reference ``torch.py:79-100``)"""

    __constants__ = ("coeff",)

    def __init__(self, coeff: float = 0.97) -> None:
        super().__init__()
        self.coeff = coeff

    @classmethod
    def from_preemphasize(cls, preemphasize: Preemphasize):
        return cls(preemphasize.coeff)

    def forward(self, sig: torch.Tensor) -> torch.Tensor:
        return pytorch_preemphasize(sig, self.coeff)


def pytorch_dither(sig: torch.Tensor, coeff: float = 1.0) -> torch.Tensor:
    """``sig + N(0, coeff^2)``; the Philox stream is seeded from torch's generator, so
    ``torch.manual_seed`` makes it reproducible like the reference's ``randn_like``"""
    if sig.numel() == 0:
        return sig.clone()
    d_sig, home = _on_device(sig)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _result(_row_launch("pds_dither", d_sig, float(coeff), seed), sig.dtype, home)


class PyTorchDither(torch.nn.Module):
    """Module form of :func:`pytorch_dither` (reference ``torch.py:108-139``)"""

    __constants__ = ("coeff",)

    def __init__(self, coeff: float = 1.0):
        _check_positive("coeff", coeff, True)
        super().__init__()
        self.coeff = coeff

    @classmethod
    def from_dither(cls, dither: Dither):
        return cls(dither.coeff)

    def forward(self, sig: torch.Tensor) -> torch.Tensor:
        return pytorch_dither(sig, self.coeff)


def _stft_math(
    sig: torch.Tensor,
    filters: Sequence[torch.Tensor],
    offsets: Sequence[int],
    frame_length: int,
    frame_shift: int,
    centered: bool,
    window: Optional[torch.Tensor],
    dft_size: int,
    use_log: bool,
    use_power: bool,
    include_energy: bool,
    kaldi_shift: bool,
    is_real: bool,
    eps: float,
) -> torch.Tensor:
    """What the fused kernel computes, restated in differentiable torch ops (BACKWARD pass only)

    Same quantities as ``compute.py:388-460`` / ``torch.py:142-235`` of the reference, arranged the
    way the kernel arranges them: symmetric padding, strided frames, ``|rfft|^p`` once, then one
    matrix product with the folded weights ``W[f, bin] = sum_taps |h_f[tap]|^p`` (the tap -> bin map
    replays the reference's segment walk, :func:`_tables._segment_bins`), doubled for real banks.
    Nothing on the forward path calls this.
    """
    n = sig.size(0)
    num_frames, pad_left, pad_right = stft_geometry(n, frame_length, frame_shift, centered, kaldi_shift)
    if pad_left or pad_right:
        sig = torch.cat([sig[:pad_left].flip(0), sig, sig[n - pad_right:].flip(0)])
    frames = sig.unfold(0, frame_length, frame_shift)[:num_frames]
    columns = []
    if include_energy:
        energy = frames.square().mean(1)
        columns.append((energy if use_power else energy.sqrt()).unsqueeze(1))
    if window is not None:
        frames = frames * window.to(frames.dtype)
    spect = torch.fft.rfft(frames, dft_size, dim=1)
    spect = spect.real.square() + spect.imag.square() if use_power else spect.abs()
    half_len = half_spectrum_len(dft_size) if dft_size % 2 == 0 else (dft_size + 1) // 2
    rows = []
    for offset, filt in zip(offsets, filters):
        gain = filt.abs().to(spect.dtype)
        if use_power:
            gain = gain.square()
        pieces = _segment_bins(int(offset), int(filt.numel()), half_len)
        bins = torch.as_tensor(np.concatenate(pieces or [np.arange(0)]).astype(np.int64), device=gain.device)
        rows.append(torch.zeros(half_len, dtype=spect.dtype, device=gain.device).index_add(0, bins, gain))
    weights = torch.stack(rows)
    if is_real:
        weights = weights * 2
    columns.append(spect @ weights.t())
    feats = torch.cat(columns, 1)
    return feats.clamp_min(eps).log() if use_log else feats


class _STFTFunction(torch.autograd.Function):
    """Forward: the fused CUDA kernel.  Backward: autograd through :func:`_stft_math`"""

    @staticmethod
    def forward(ctx, module, signal, window, *filters):
        ctx.module = module
        ctx.has_window = window is not None
        ctx.save_for_backward(signal, *([window] if window is not None else []), *filters)
        return module._kernel_forward(signal)

    @staticmethod
    def backward(ctx, grad_out):
        from ._gpu import current_device

        module = ctx.module
        saved = ctx.saved_tensors
        signal, rest = saved[0], list(saved[1:])
        window = rest.pop(0) if ctx.has_window else None
        filters = rest
        device = signal.device if signal.device.type == "cuda" else current_device()
        needs = ctx.needs_input_grad  # (module, signal, window, *filters)
        work = signal.dtype if signal.dtype in (torch.float32, torch.float64) else torch.float32
        with torch.enable_grad():
            sig = signal.detach().to(device=device, dtype=work).requires_grad_(needs[1])
            win = None if window is None else window.detach().to(device).requires_grad_(needs[2])
            filts = [f.detach().to(device).requires_grad_(needs[3 + i]) for i, f in enumerate(filters)]
            out = _stft_math(
                sig, filts, module.offsets, module.frame_length, module.frame_shift, module.centered, win,
                module.dft_size, module.use_log, module.use_power, module.include_energy, module.kaldi_shift,
                module.is_real, float(config.LOG_FLOOR_VALUE))
            wanted = [t for t in [sig, win] + filts if t is not None and t.requires_grad]
            grads = list(torch.autograd.grad(out, wanted, grad_out.to(device=device, dtype=out.dtype))) if wanted else []
        result = [None]
        for original, leaf in zip([signal, window] + list(filters), [sig, win] + filts):
            if leaf is None or not leaf.requires_grad:
                result.append(None)
            else:
                result.append(grads.pop(0).to(device=original.device, dtype=original.dtype))
        return tuple(result)


class PyTorchShortTimeFourierTransformFrameComputer(torch.nn.Module):
    """Fused-kernel STFT features as a module (reference ``torch.py:238-429``)

    Parameters are the reference's: ``offsets_and_truncated_filters`` is a sequence of
    ``(offset, truncated frequency response)`` pairs, the rest is the frame geometry in samples.
    The easiest constructor is :func:`from_stft_frame_computer`.
    """

    def __init__(
        self,
        offsets_and_truncated_filters: Sequence[Tuple[int, torch.Tensor]],
        frame_length: int,
        frame_shift: int,
        frame_style: str = "centered",
        window: Optional[torch.Tensor] = None,
        dft_size: Optional[int] = None,
        use_log: bool = True,
        use_power: bool = False,
        include_energy: bool = False,
        kaldi_shift: bool = False,
        is_real: bool = False,
    ) -> None:
        offsets, filters = [], []
        for i, (offset, filt) in enumerate(offsets_and_truncated_filters):
            filt = torch.as_tensor(filt)
            if filt.ndim != 1:
                raise ValueError(f"filter {i} is not a vector")
            if not filt.size(0):
                raise ValueError(f"filter {i} is empty")
            _check_positive(f"filter {i} offset", offset, True)
            offsets.append(int(offset))
            filters.append(filt.detach().cpu().numpy())
        _check_positive("frame_length", frame_length)
        _check_positive("frame_shift", frame_shift)
        if frame_style not in {"causal", "centered"}:
            raise ValueError(
                f"Expected frame_style to be one of 'causal', 'centered'; got '{frame_style}'"
            )
        if window is not None and tuple(window.shape) != (frame_length,):
            raise ValueError(f"Expected window.shape to be ({frame_length},); got {tuple(window.shape)}")
        if dft_size is None:
            dft_size = 2 ** math.ceil(math.log(frame_length, 2))
        elif dft_size < frame_length:
            raise ValueError(f"Expected dft_size to be gte {frame_length}; got {dft_size}")
        super().__init__()
        self.frame_length, self.frame_shift = frame_length, frame_shift
        self.offsets, self.centered = tuple(offsets), frame_style == "centered"
        self.dft_size, self.use_log, self.use_power = dft_size, use_log, use_power
        self.kaldi_shift, self.is_real, self.include_energy = kaldi_shift, is_real, include_energy
        self._frame_style = frame_style
        # Same parameter names as the reference (torch.py:362-366: ``filters.<i>`` and ``window``), so
        # its state_dicts load here and ours load there; learnable, like the reference's.
        self.filters = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(f).clone()) for f in filters])
        if window is None:
            self.register_parameter("window", None)
        else:
            self.window = torch.nn.Parameter(window.detach().clone())
        self._computer = None
        self._rebuild()
        self.register_load_state_dict_post_hook(lambda module, _: module._rebuild())

    def _versions(self):
        return tuple(p._version for p in self.parameters())

    def _rebuild(self) -> None:
        """(Re)build the kernel plan from the current parameters (construction, load_state_dict, and
        whenever an optimizer step or any other in-place update has touched them)"""
        self._built_versions = self._versions()
        window = None if self.window is None else self.window.detach().cpu().double().numpy()
        self._computer = STFTFrameComputer.from_tables(
            list(self.offsets), [f.detach().cpu().numpy() for f in self.filters], self.frame_length,
            self.frame_shift, self._frame_style, window, self.dft_size, self.use_log, self.use_power,
            self.include_energy, self.kaldi_shift, self.is_real,
        )

    @classmethod
    def from_stft_frame_computer(
        cls,
        computer: STFTFrameComputer,
        filter_type: torch.dtype = torch.cfloat,
        window_type: torch.dtype = torch.float,
    ):
        """Module equivalent to an already-built :class:`STFTFrameComputer` (shares its plan)"""
        pairs = [
            (o, torch.as_tensor(np.asarray(x)).to(filter_type))
            for o, x in zip(computer._filt_start_idxs, computer._truncated_filts)
        ]
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.frame_length, self.frame_shift = computer.frame_length, computer.frame_shift
        self.offsets = tuple(int(o) for o, _ in pairs)
        self.centered = computer.frame_style == "centered"
        self.dft_size, self.use_log, self.use_power = computer._dft_size, computer._log, computer._power
        self.kaldi_shift, self.is_real = computer._kaldi_shift, computer._real
        self.include_energy = computer._include_energy
        self._frame_style = computer.frame_style
        self.filters = torch.nn.ParameterList([torch.nn.Parameter(x) for _, x in pairs])
        self.window = torch.nn.Parameter(torch.as_tensor(computer._window).to(window_type))
        self._computer = computer  # full-precision tables: no float32/complex64 round trip
        self._built_versions = self._versions()
        self.register_load_state_dict_post_hook(lambda module, _: module._rebuild())
        return self

    def forward(self, signal: torch.Tensor) -> torch.Tensor:
        if signal.ndim != 1:
            raise RuntimeError(f"Expected x to be 1-dimensional; got {signal.ndim}")
        if signal.size(0) < self.frame_length // 2 + 1:
            # like the reference (torch.py:179-180): num_filts columns, without the energy column
            return signal.new_empty((0, len(self.offsets)))
        if self._versions() != self._built_versions:
            self._rebuild()  # the parameters were updated in place (optimizer step): new kernel tables
        if torch.is_grad_enabled() and (signal.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _STFTFunction.apply(self, signal, self.window, *self.filters)
        return self._kernel_forward(signal)

    def _kernel_forward(self, signal: torch.Tensor) -> torch.Tensor:
        computer = self._computer
        signal = signal.detach()
        d_sig, home = _on_device(signal)
        if d_sig.dtype not in (torch.float32, torch.int16):
            d_sig = d_sig.float()
        offsets = np.zeros(1, np.int64)
        lengths = np.array([d_sig.numel()], np.int64)
        feats, _ = computer.compute_packed_device(d_sig, offsets, lengths)
        return _result(feats, signal.dtype, home)


PyTorchSTFTFrameComputer = PyTorchShortTimeFourierTransformFrameComputer


def pytorch_stft_frame_computer(
    sig: torch.Tensor,
    filters: List[torch.Tensor],
    offsets: List[int],
    frame_length: int,
    frame_shift: int,
    centered: bool = True,
    window: Optional[torch.Tensor] = None,
    dft_size: Optional[int] = None,
    use_log: bool = True,
    use_power: bool = False,
    include_energy: bool = False,
    kaldi_shift: bool = False,
    is_real: bool = True,
    eps: float = config.LOG_FLOOR_VALUE,
) -> torch.Tensor:
    """Functional form (reference ``torch.py:142-235``); builds a plan per call, so prefer the module"""
    if dft_size is not None and dft_size < frame_length:
        raise RuntimeError(f"expected dft_size gte {frame_length}; got {dft_size}")
    if len(filters) != len(offsets):
        raise RuntimeError(
            f"filters ({len(filters)}) has different length than offsets ({len(offsets)})"
        )
    if sig.ndim != 1:
        raise RuntimeError(f"Expected x to be 1-dimensional; got {sig.ndim}")
    if window is not None and tuple(window.shape) != (frame_length,):
        raise RuntimeError(
            f"Expected window to have shape {(frame_length,)}; got {tuple(window.shape)}"
        )
    if eps != config.LOG_FLOOR_VALUE:
        raise NotImplementedError("the kernels floor at config.LOG_FLOOR_VALUE")
    module = PyTorchShortTimeFourierTransformFrameComputer(
        list(zip(offsets, filters)), frame_length, frame_shift,
        "centered" if centered else "causal", window, dft_size, use_log, use_power,
        include_energy, kaldi_shift, is_real,
    )
    return module(sig)


class PyTorchPostProcessorWrapper(torch.nn.Module):
    """Apply a :class:`PostProcessor` to a tensor (reference ``torch.py:435-472``)

    Same contract as the reference wrapper (``apply`` with its default ``axis=-1``); the
    post-processor itself runs its CUDA kernel.
    """

    def __init__(self, postprocessor: PostProcessor):
        super().__init__()
        self.postprocessor = postprocessor

    @classmethod
    def from_postprocessor(cls, postprocessor: PostProcessor):
        return cls(postprocessor)

    def forward(self, sig: torch.Tensor) -> torch.Tensor:
        out = self.postprocessor.apply(sig.detach().cpu().numpy())
        return torch.as_tensor(out, device=sig.device, dtype=sig.dtype)


class PyTorchShortIntegrationFrameComputer(torch.nn.Module):
    """Module around :class:`SIFrameComputer` (reference ``torch.py:475-519``)"""

    def __init__(self, si_frame_computer: SIFrameComputer):
        super().__init__()
        self.si_frame_computer = si_frame_computer

    @classmethod
    def from_si_frame_computer(cls, si_frame_computer: SIFrameComputer):
        return cls(si_frame_computer)

    def state_dict(self, *args, **kwargs):
        raise NotImplementedError

    def load_state_dict(self, *args, **kwargs):
        raise NotImplementedError

    def forward(self, sig: torch.Tensor) -> torch.Tensor:
        computer = self.si_frame_computer
        if sig.ndim != 1:
            raise RuntimeError(f"Expected x to be 1-dimensional; got {sig.ndim}")
        if computer.num_frames(sig.numel()) == 0:
            return sig.new_empty((0, computer.num_coeffs))
        d_sig, home = _on_device(sig)
        feats, _ = computer.compute_packed_device(
            d_sig.float(), np.zeros(1, np.int64), np.array([d_sig.numel()], np.int64)
        )
        return _result(feats, sig.dtype, home)


PyTorchSIFrameComputer = PyTorchShortIntegrationFrameComputer
