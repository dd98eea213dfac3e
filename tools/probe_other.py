"""Device-side timing probe of the non-headline kernels (development aid, not the bench)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

dev = torch.device("cuda", 0)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    return best


rng = np.random.default_rng(0)
# ---- C3: gammatone-64 STFT -----------------------------------------------------------------
cfg = {"name": "stft", "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 64, "erb": True},
       "frame_length_ms": 25, "use_power": True}
comp = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
lengths = (16000 * rng.uniform(2, 20, 1000)).astype(np.int64)
offsets, total = PackedSignals.layout(lengths, comp.pad_left % 4)
d_sig = torch.randn(total, device=dev) * 1000
layout = comp.plan_batch(offsets, lengths, dev)
out = torch.empty((layout.rows, comp.num_coeffs), device=dev)
ms = timeit(lambda: comp.run_batch(layout, d_sig, out=out))
hours = lengths.sum() / 16000 / 3600
print(f"C3 gammatone64 stft: {ms:.3f} ms  {hours / (ms * 1e-3):.1f} audio-h/s  frames/s={layout.rows / (ms * 1e-3):.3e} nnz={comp.folded_weights.nnz}")

# ---- C4: SI gabor-41 -----------------------------------------------------------------------
cfg = {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}}
si = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
lengths = np.array([16000 * 60] * 20 + [16000 * 600] * 2, dtype=np.int64)
offsets, total = PackedSignals.layout(lengths, 0)
d_sig = torch.randn(total, device=dev) * 1000
ms = timeit(lambda: si.compute_packed_device(d_sig, offsets, lengths), reps=3)
hours = lengths.sum() / 16000 / 3600
frames = sum(si.num_frames(int(n)) for n in lengths)
print(f"C4 SI gabor41: {ms:.3f} ms  {hours / (ms * 1e-3):.2f} audio-h/s  frames/s={frames / (ms * 1e-3):.3e} "
      f"max_support={si._max_support} GFMA/s={frames * 160 * 41 * si._max_support * 2 / (ms * 1e-3) / 1e9:.0f}")

# ---- C5 post: deltas + cmvn on 10.99M x 41 -------------------------------------------------
rows = 10_989_392
feats = torch.randn((rows, 41), device=dev)
row_off = torch.tensor([0, rows], dtype=torch.int64, device=dev)
deltas = pds.post.Deltas(2)
ms = timeit(lambda: deltas.apply_device(feats, row_off))
print(f"deltas(2) {rows}x41 -> x123: {ms:.3f} ms  {(rows * 41 * 4 + rows * 123 * 4) / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic")
full = deltas.apply_device(feats, row_off)
std = pds.post.Standardize()
lib = pds._lib.get_lib()
d_stats = torch.zeros((2, 124), dtype=torch.float64, device=dev)
from pydrobert_speech_b200._gpu import stream_ptr  # noqa: E402
ms = timeit(lambda: lib.pds_cmvn_accumulate(full.data_ptr(), rows, 123, d_stats.data_ptr(), stream_ptr(dev)))
print(f"cmvn stats {rows}x123: {ms:.3f} ms  {rows * 123 * 4 / (ms * 1e-3) / 1e9:.0f} GB/s")
std.accumulate_device(full)
ms = timeit(lambda: std.apply_device(full, out=full))
print(f"cmvn apply {rows}x123: {ms:.3f} ms  {2 * rows * 123 * 4 / (ms * 1e-3) / 1e9:.0f} GB/s")
