"""Bring-up / timing probe of stft_umma_kernel (PDS_STFT_KERNEL=u): features against the default
kernel and the oracle on a ragged batch, then the time of one launch over the bench corpus."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
mode = sys.argv[1] if len(sys.argv) > 1 else "check"
rng = np.random.default_rng(5)
lengths = [16000, 201, 7777, 160 * 32 + 240, 48000, 399, 160 * 64 + 241, 33333]
signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]


def run(kernel, config=cfg, sigs=signals):
    if kernel:
        os.environ["PDS_STFT_KERNEL"] = kernel
    else:
        os.environ.pop("PDS_STFT_KERNEL", None)
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, config)
    out = computer.compute_batch(sigs)
    return computer, out


if mode == "check":
    import oracle
    for name, config in (("readme", cfg), ("linear", dict(cfg, use_log=False)),
                         ("magnitude", dict(cfg, use_power=False, use_log=False)),
                         ("no-energy", dict(cfg, include_energy=False))):
        _, ref = run(None, config)
        computer, got = run("u", config)
        print(name, "kernel:", computer.kernel_name())
        for sig, a, b in zip(signals, got, ref):
            if not len(b):
                assert a.shape == b.shape
                continue
            want = oracle.stft_features(
                sig.astype(np.float64), computer._window, computer._dft_size, computer._filt_start_idxs,
                computer._truncated_filts, computer.frame_shift, computer.pad_left, computer._power, computer._log,
                computer.includes_energy, computer._real)
            if computer._log:
                print(f"  len {len(sig):6d}: vs oracle {np.abs(a - want).max():.3e}  default vs oracle {np.abs(b - want).max():.3e}")
            else:
                scale = np.maximum(np.abs(want), 1e-6 * np.abs(want).max(axis=1, keepdims=True))
                raw = np.abs(a - want) / np.maximum(np.abs(want), 1e-300)
                print(f"  len {len(sig):6d}: linear rel err {(np.abs(a - want) / scale).max():.3e} (un-floored {raw.max():.3e})"
                      f"  default {(np.abs(b - want) / scale).max():.3e}")
            bad = np.argwhere(~np.isfinite(a))
            if len(bad):
                print("   non-finite at", bad[:5])
else:
    dev = torch.device("cuda", 0)
    n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    lens = rng.integers(16000 * 8, 16000 * 14, n_utts).astype(np.int64)
    offsets, total = PackedSignals.layout(lens, 199 % 4)  # pad_left % 4: every tile starts on a 16-byte grid (TMA path)
    d_sig = torch.randn(total, device=dev) * 1000
    hours = lens.sum() / 16000 / 3600
    probe_ix = 0
    for kernel in ((None, "u") if mode == "time" else ("u",) * (4 if mode == "probe" else 1)):
        if mode == "probe":
            os.environ["PDS_W_PROBE"] = str(probe_ix); probe_ix += 1
        if kernel:
            os.environ["PDS_STFT_KERNEL"] = kernel
        else:
            os.environ.pop("PDS_STFT_KERNEL", None)
        computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
        layout = computer.plan_batch(offsets, lens, dev)
        for _ in range(3):
            feats = computer.run_batch(layout, d_sig)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            feats = computer.run_batch(layout, d_sig)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        print(f"{computer.kernel_name():26s} {best:8.3f} ms  {hours / (best * 1e-3):9.1f} audio-h/s  ({hours:.1f} h, "
              f"{int(feats.shape[0])} frames) checksum {float(feats.double().sum()):.6e}")
