"""Throughput of the fused STFT kernel at other DFT sizes (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
lengths = (16000 * rng.uniform(2, 20, 3000)).astype(np.int64)
for ms, shift in ((16, 10), (25, 10), (32, 10), (50, 10), (64, 16), (100, 20)):
    cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": ms, "frame_shift_ms": shift, "include_energy": True,
           "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
    offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
    d_sig = torch.randn(total, device=dev) * 1000
    layout = computer.plan_batch(offsets, lengths, dev)
    feats = torch.empty((layout.rows, computer.num_coeffs), device=dev)
    for _ in range(3):
        computer.run_batch(layout, d_sig, out=feats)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        computer.run_batch(layout, d_sig, out=feats)
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    hours = lengths.sum() / 16000 / 3600
    print(f"L={computer.frame_length} S={computer.frame_shift} N={computer._dft_size}: {best:.3f} ms  "
          f"{hours / (best * 1e-3):.0f} audio-h/s  frames/s={layout.rows / (best * 1e-3):.3e}")
