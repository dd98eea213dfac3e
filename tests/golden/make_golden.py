"""Generate the golden vectors in this directory by running the REAL reference.

Run in the build container (the reference is importable there, not on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed): ``stft.npz``, ``si.npz``, ``banks.npz``, ``post.npz``, ``kaldi.npz``, ``extra.npz``,
``si_long.npz``, ``torch_grad.npz`` (``python tests/golden/make_golden.py extra`` / ``si_long`` / ``torch_grad``
rebuild the last three alone).
The reference's numpy path is used (``config.USE_FFTPACK = False``); its scipy.fftpack branch
is pinned to it by the reference's own tests.  Nothing here is imported at test time except
``cases.py``.
"""

import os
import pickle
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
import pydrobert.speech.config as ref_config  # noqa: E402
from pydrobert.speech import compute, filters, post, pre  # noqa: E402
from pydrobert.speech.alias import alias_factory_subclass_from_arg as build  # noqa: E402

ref_config.USE_FFTPACK = False


def make_signal(spec):
    if spec[0] == "randn":
        _, seed, length = spec
        return np.random.default_rng(seed).standard_normal(length) * 1000.0
    _, length = spec
    with wave.open(os.path.join(REF, "extras", "test.wav")) as handle:
        data = np.frombuffer(handle.readframes(handle.getnframes()), dtype="<i2")
    return data[:length].astype(np.float64)


def ragged(prefix, arrays, out):
    arrays = [np.asarray(a) for a in arrays]
    out[prefix + "/offsets"] = np.cumsum([0] + [len(a) for a in arrays])
    out[prefix + "/values"] = np.concatenate(arrays) if arrays else np.zeros(0)


def stft_goldens():
    out = {}
    for name, (cfg, spec) in cases.STFT_CASES.items():
        computer = build(compute.FrameComputer, cfg)
        signal = make_signal(spec)
        out[name + "/signal"] = signal.astype(np.float32)  # what the kernels are fed
        feats = computer.compute_full(out[name + "/signal"].astype(np.float64))
        out[name + "/feats"] = feats
        out[name + "/window"] = computer._window
        out[name + "/dft_size"] = computer._dft_size
        out[name + "/starts"] = np.array(computer._filt_start_idxs)
        ragged(name + "/filts", computer._truncated_filts, out)
        out[name + "/geometry"] = np.array(
            [computer.frame_length, computer.frame_shift, int(computer.frame_style == "centered"),
             int(computer._kaldi_shift), int(computer._real)]
        )
        # linear (pre-log) features for the relative tolerance
        if cfg.get("use_log", True):
            lin = build(compute.FrameComputer, dict(cfg, use_log=False))
            out[name + "/feats_linear"] = lin.compute_full(out[name + "/signal"].astype(np.float64))
    computer = build(compute.FrameComputer, cases.README_FBANK)
    base = (np.random.default_rng(99).standard_normal(max(cases.EDGE_LENGTHS)) * 1000.0).astype(np.float32)
    out["edge/signal"] = base
    for n in cases.EDGE_LENGTHS:
        out[f"edge/feats_{n}"] = computer.compute_full(base[:n].astype(np.float64))
    # pre-emphasis fused in front of the README config
    sig = make_signal(("randn", 21, 5000)).astype(np.float32)
    out["preemph/signal"] = sig
    out["preemph/feats"] = computer.compute_full(pre.Preemphasize(0.97).apply(sig.astype(np.float64)))
    out["preemph/signal_out"] = pre.Preemphasize(0.97).apply(sig.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "stft.npz"), **out)


def si_goldens():
    out = {}
    for name, (cfg, spec) in cases.SI_CASES.items():
        computer = build(compute.FrameComputer, cfg)
        signal = make_signal(spec).astype(np.float32)
        out[name + "/signal"] = signal
        out[name + "/feats"] = computer.compute_full(signal.astype(np.float64))
        if cfg.get("use_log", True):
            lin = build(compute.FrameComputer, dict(cfg, use_log=False))
            out[name + "/feats_linear"] = lin.compute_full(signal.astype(np.float64))
        # impulse responses actually used (after roll + clamp), recovered from the stored DFTs
        irs = [computer._compute_idft(f.copy())[: computer._max_support] for f in computer._filts]
        out[name + "/impulse"] = np.array(irs)
        out[name + "/window"] = computer._window.flatten()
        out[name + "/geometry"] = np.array(
            [computer.frame_shift, computer._max_support, computer._translation,
             computer._frame_length, computer._dft_size, int(computer.frame_style == "centered")]
        )
        # frame counts over a sweep of lengths (causal computers lose a frame)
        lens = np.arange(0, 4 * computer.frame_shift + 3, 37)
        counts = []
        for n in lens:
            try:
                counts.append(computer.compute_full(np.zeros(n)).shape[0])
            except ValueError:
                counts.append(-1)
        out[name + "/sweep_lens"], out[name + "/sweep_counts"] = lens, np.array(counts)
    np.savez_compressed(os.path.join(HERE, "si.npz"), **out)


def bank_goldens():
    out = {}
    for name, (cfg, width) in cases.BANK_CASES.items():
        bank = build(filters.LinearFilterBank, cfg)
        out[name + "/supports"] = np.array(bank.supports)
        out[name + "/supports_hz"] = np.array(bank.supports_hz)
        out[name + "/flags"] = np.array([bank.is_real, bank.is_analytic, bank.is_zero_phase, bank.num_filts])
        starts, truncs, freqs, halves, imps = [], [], [], [], []
        for i in range(bank.num_filts):
            s, t = bank.get_truncated_response(i, width)
            starts.append(s)
            truncs.append(t)
            freqs.append(bank.get_frequency_response(i, width))
            halves.append(bank.get_frequency_response(i, width, half=True))
            imps.append(bank.get_impulse_response(i, width))
        out[name + "/starts"] = np.array(starts)
        ragged(name + "/trunc", truncs, out)
        out[name + "/freq"] = np.array(freqs)
        out[name + "/half"] = np.array(halves)
        out[name + "/impulse"] = np.array(imps)
    for i, (cfg, width) in enumerate(cases.WINDOW_CASES):
        out[f"window{i}"] = build(filters.WindowFunction, cfg).get_impulse_response(width)
    np.savez_compressed(os.path.join(HERE, "banks.npz"), **out)


def post_goldens():
    rng = np.random.default_rng(31)
    out = {}
    feats = rng.standard_normal((57, 7)) * 3 + 1
    out["feats"] = feats
    for order in (1, 2, 3):
        for ctx in (1, 2, 3):
            out[f"deltas_o{order}_w{ctx}"] = post.Deltas(order, context_window=ctx).apply(feats, axis=0)
    out["deltas_short"] = post.Deltas(2).apply(feats[:3], axis=0)
    std = post.Standardize()
    chunks = [rng.standard_normal((n, 7)) * (1 + np.arange(7)) + np.arange(7) for n in (30, 11, 64)]
    for c in chunks:
        std.accumulate(c.astype(np.float32))
    out["cmvn_chunks"] = np.concatenate(chunks).astype(np.float32)
    out["cmvn_chunk_lens"] = np.array([len(c) for c in chunks])
    out["cmvn_stats"] = std._stats
    out["cmvn_applied"] = std.apply(chunks[0].astype(np.float32))
    out["cmvn_applied_novar"] = post.Standardize(norm_var=False).apply(feats)
    out["cmvn_local"] = post.Standardize().apply(feats)
    np.savez_compressed(os.path.join(HERE, "post.npz"), **out)


def kaldi_goldens():
    """Carry the reference's own known-answer fixtures over (tests/data/*.pkl)"""
    data = os.path.join(REF, "tests", "data")
    out = {}
    with open(os.path.join(data, "noise.pkl"), "rb") as f:
        out["noise"] = np.asarray(pickle.load(f))
    with open(os.path.join(data, "kaldi_feats.pkl"), "rb") as f:
        out["kaldi_feats"] = np.asarray(pickle.load(f))
    with open(os.path.join(data, "kaldi_filts.pkl"), "rb") as f:
        filts = pickle.load(f)
    out["kaldi_filt_offsets"] = np.array([o for o, _ in filts])
    ragged("kaldi_filt", [v for _, v in filts], out)
    np.savez_compressed(os.path.join(HERE, "kaldi.npz"), **out)


def extra_goldens():
    """Round-2 additions (``extra.npz``): BASELINE config 1 at its full size (all of extras/test.wav:
    149 940 samples -> 937 x 41), one 60 s utterance of config 4 (SI + Gabor-41), ``post.Stack``"""
    out = {}
    with wave.open(os.path.join(REF, "extras", "test.wav")) as handle:
        wav = np.frombuffer(handle.readframes(handle.getnframes()), dtype="<i2")
    out["c1/signal"] = wav
    computer = build(compute.FrameComputer, cases.README_FBANK)
    out["c1/feats"] = computer.compute_full(wav.astype(np.float64)).astype(np.float32)
    lin = build(compute.FrameComputer, dict(cases.README_FBANK, use_log=False))
    out["c1/feats_linear"] = lin.compute_full(wav.astype(np.float64))
    si = build(compute.FrameComputer, cases.SI_GABOR_41)
    signal = (np.random.default_rng(cases.C4_SEED).standard_normal(cases.C4_SAMPLES) * 1000.0).astype(np.float32)
    out["c4_60s/feats"] = si.compute_full(signal.astype(np.float64)).astype(np.float32)
    rng = np.random.default_rng(41)
    feats = rng.standard_normal((23, 5))
    out["stack/feats"] = feats
    for name, kwargs, axis in cases.STACK_CASES:
        out["stack/" + name] = post.Stack(**kwargs).apply(feats, axis=axis)
    np.savez_compressed(os.path.join(HERE, "extra.npz"), **out)


def si_long_goldens():
    """``si_long.npz``: short-integration features of banks whose impulse responses are longer than
    one 1024-point block (the signals are regenerated from their seeds at test time)"""
    out = {}
    for name, (cfg, spec) in cases.SI_LONG_CASES.items():
        signal = make_signal(spec).astype(np.float32).astype(np.float64)
        computer = build(compute.FrameComputer, cfg)
        out[name + "/feats"] = computer.compute_full(signal).astype(np.float32)
        lin = build(compute.FrameComputer, dict(cfg, use_log=False))
        out[name + "/feats_linear"] = lin.compute_full(signal)
        out[name + "/geometry"] = np.array(
            [computer.frame_shift, computer._max_support, computer._translation,
             computer._frame_length, computer._dft_size, int(computer.frame_style == "centered")]
        )
    np.savez_compressed(os.path.join(HERE, "si_long.npz"), **out)


TORCH_GRAD_CASES = {  # real banks only: the reference's torch path and its NumPy path agree there
    "fbank10_power_log_energy": dict(bank={"name": "fbank", "num_filts": 10}, frame_length_ms=25, include_energy=True,
                                     window_function="hanning", use_power=True, use_log=True),
    "fbank8_magnitude_causal": dict(bank={"name": "fbank", "num_filts": 8}, frame_length_ms=20, frame_shift_ms=8,
                                    frame_style="causal", include_energy=False, window_function="hamming",
                                    use_power=False, use_log=False),
}


def torch_grad_goldens():
    """``torch_grad.npz``: outputs and gradients (signal, window, every filter) of the reference's OWN torch
    module (``torch.py:142-235``) in float64, for the backward pass of the torch mirror"""
    import torch
    from pydrobert.speech import torch as ref_torch

    out = {}
    rng = np.random.default_rng(77)
    for name, kwargs in TORCH_GRAD_CASES.items():
        computer = build(compute.FrameComputer, dict(kwargs, name="stft"))
        module = ref_torch.PyTorchSTFTFrameComputer.from_stft_frame_computer(computer, torch.cdouble, torch.double)
        signal = torch.tensor(rng.standard_normal(3000) * 100.0, dtype=torch.double, requires_grad=True)
        feats = module(signal)
        grad_out = torch.tensor(rng.standard_normal(tuple(feats.shape)))
        leaves = [signal, module.window] + list(module.filters)
        grads = torch.autograd.grad(feats, leaves, grad_out)
        out[name + "/signal"] = signal.detach().numpy()
        out[name + "/feats"] = feats.detach().numpy()
        out[name + "/grad_out"] = grad_out.numpy()
        out[name + "/grad_signal"] = grads[0].numpy()
        out[name + "/window"] = module.window.detach().numpy()
        out[name + "/grad_window"] = grads[1].numpy()
        out[name + "/offsets"] = np.array(computer._filt_start_idxs, dtype=np.int64)
        ragged(name + "/filters", [f.detach().numpy() for f in module.filters], out)
        ragged(name + "/grad_filters", [g.numpy() for g in grads[2:]], out)
        out[name + "/geometry"] = np.array(
            [computer.frame_length, computer.frame_shift, computer._dft_size, int(computer.frame_style == "centered")])
    np.savez_compressed(os.path.join(HERE, "torch_grad.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "torch_grad":
        torch_grad_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        extra_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "si_long":
        si_long_goldens()
        sys.exit(0)
    stft_goldens()
    si_goldens()
    bank_goldens()
    post_goldens()
    kaldi_goldens()
    extra_goldens()
    si_long_goldens()
    torch_grad_goldens()
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".npz"):
            print(name, os.path.getsize(os.path.join(HERE, name)))
