"""Quick device-side timing probe of the fused STFT kernel (development aid, not the bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
rng = np.random.default_rng(0)
lengths = (16000 * rng.uniform(2, 20, n_utts)).astype(np.int64)
offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
d_sig = torch.randn(total, device=dev, generator=gen) * 1000
audio_h = lengths.sum() / 16000 / 3600
layout = computer.plan_batch(offsets, lengths, dev)
frame_off = layout.frame_off
feats = torch.empty((layout.rows, computer.num_coeffs), device=dev)
for rep in range(3):
    computer.run_batch(layout, d_sig, out=feats)
torch.cuda.synchronize()
times = []
for rep in range(10):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    computer.run_batch(layout, d_sig, out=feats)
    t1.record(); torch.cuda.synchronize()
    times.append(t0.elapsed_time(t1))
ms = min(times)
frames = int(frame_off[-1])
print(f"utts={n_utts} audio_h={audio_h:.3f} frames={frames} best_ms={ms:.3f} mean_ms={np.mean(times):.3f}")
print(f"audio-h/s={audio_h / (ms * 1e-3):.1f}  frames/s={frames / (ms * 1e-3):.3e}")
bytes_alg = frames * 804
print(f"algorithmic GB/s={bytes_alg / (ms * 1e-3) / 1e9:.1f}  flops(14559/frame) TF/s={frames * 14559 / (ms * 1e-3) / 1e12:.2f}")
