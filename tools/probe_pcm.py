"""Device-side timing of the fused STFT kernel on int16 PCM input (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
rng = np.random.default_rng(0)
lengths = (16000 * rng.uniform(2, 20, 4000)).astype(np.int64)
offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
dev = torch.device("cuda", 0)
d_f32 = (torch.randn(total, device=dev) * 1000).round_()
d_i16 = d_f32.to(torch.int16)
layout = computer.plan_batch(offsets, lengths, dev)
feats_a = torch.empty((layout.rows, computer.num_coeffs), device=dev)
feats_b = torch.empty_like(feats_a)
for name, sig, out in (("float32", d_f32, feats_a), ("int16", d_i16, feats_b)):
    for _ in range(3):
        computer.run_batch(layout, sig, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        computer.run_batch(layout, sig, out=out)
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    print(f"{name}: {best:.3f} ms  frames/s={layout.rows / (best * 1e-3):.3e}")
for name, kw in (("float32 + preemph 0.97", dict(preemph=0.97)), ("float32 + dither 1.0", dict(dither=1.0)),
                 ("int16 + preemph 0.97", dict(preemph=0.97)),
                 ("float32 + dither, then preemph", dict(dither=1.0, preemph=0.97, dither_first=True)),
                 ("float32 + preemph, then dither", dict(dither=1.0, preemph=0.97, dither_first=False))):
    sig = d_i16 if name.startswith("int16") else d_f32
    for _ in range(3):
        computer.run_batch(layout, sig, out=feats_b, **kw)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        computer.run_batch(layout, sig, out=feats_b, **kw)
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    print(f"{name}: {best:.3f} ms  frames/s={layout.rows / (best * 1e-3):.3e}")
computer.run_batch(layout, d_i16, out=feats_b)
torch.cuda.synchronize()
print("max |float32 - int16| =", float((feats_a - feats_b).abs().max()))
