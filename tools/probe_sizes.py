"""Device-side throughput of the fused STFT path for other DFT sizes (development aid):
power-of-two sizes on the tensor-core kernels, non-power-of-two sizes on the Bluestein kernel and on
the direct-DFT kernel it replaces."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)


def run(label, cfg, n_utts, env=None):
    for key, value in (env or {}).items():
        os.environ[key] = value
    comp = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
    rate = comp.sampling_rate
    lengths = (rate * rng.uniform(2, 20, n_utts)).astype(np.int64)
    offsets, total = PackedSignals.layout(lengths, comp.pad_left % 4)
    d_sig = torch.randn(total, device=dev) * 1000
    layout = comp.plan_batch(offsets, lengths, dev)
    out = torch.empty((layout.rows, comp.num_coeffs), device=dev)
    for _ in range(3):
        comp.run_batch(layout, d_sig, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        comp.run_batch(layout, d_sig, out=out)
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    hours = lengths.sum() / rate / 3600
    print(f"{label:34s} N={comp._dft_size:5d} kernel={comp.kernel_name():28s} {best:8.3f} ms  "
          f"{layout.rows / best / 1e6:6.3f} G frames/s  {hours / (best * 1e-3):8.1f} audio-h/s", flush=True)
    for key in (env or {}):
        os.environ.pop(key)


FBANK = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
         "window_function": "hanning", "use_power": True}
run("fbank 25 ms, padded (512)", dict(FBANK, pad_to_nearest_power_of_two=True), 2000)
run("fbank 25 ms, not padded (400)", dict(FBANK, pad_to_nearest_power_of_two=False), 2000)
run("  same on the direct-DFT kernel", dict(FBANK, pad_to_nearest_power_of_two=False), 200, {"PDS_STFT_NO_BLUESTEIN": "1"})
run("fbank 20 ms, not padded (320)", dict(FBANK, frame_length_ms=20, pad_to_nearest_power_of_two=False), 2000)
run("fbank 16 ms (256)", dict(FBANK, frame_length_ms=16), 2000)
run("fbank 50 ms (1024)", dict(FBANK, frame_length_ms=50), 1000)
run("fbank 100 ms (2048)", dict(FBANK, frame_length_ms=100), 500)
