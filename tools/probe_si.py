"""Timing probe of the short-integration kernel (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

dev = torch.device("cuda", 0)
cfg = {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}}
si = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
n_long = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lengths = np.array([16000 * 60] * 20 + [16000 * 600] * n_long, dtype=np.int64)
offsets, total = PackedSignals.layout(lengths, 0)
d_sig = torch.randn(total, device=dev) * 1000
for _ in range(2):
    si.compute_packed_device(d_sig, offsets, lengths)
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    si.compute_packed_device(d_sig, offsets, lengths)
    t1.record()
    torch.cuda.synchronize()
    best = min(best, t0.elapsed_time(t1))
hours = lengths.sum() / 16000 / 3600
print(f"C4 SI gabor41: {best:.3f} ms  {hours / (best * 1e-3):.2f} audio-h/s")
