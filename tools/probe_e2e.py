"""Where does the end-to-end time go?  (development aid)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402
from pydrobert_speech_b200.pipeline import FeaturePipeline  # noqa: E402

dev = torch.device("cuda", 0)
cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
rng = np.random.default_rng(0)
lengths = (16000 * rng.uniform(2, 20, 10000)).astype(np.int64)
offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
host_sig = torch.randn(total).mul_(1000).pin_memory()
frames = sum(computer.num_frames(int(n)) for n in lengths)
host_out = torch.empty((frames, 41)).pin_memory()
hours = lengths.sum() / 16000 / 3600
chunk = 1 << 26


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {dt * 1e3:.1f} ms  ({hours / dt:.1f} audio-h/s)")


d_buf = torch.empty(chunk, device=dev)
d_out = torch.empty((chunk // 160 + 64, 41), device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def h2d_only():
    with torch.cuda.stream(s_in):
        for a in range(0, total, chunk):
            b = min(total, a + chunk)
            d_buf[: b - a].copy_(host_sig[a:b], non_blocking=True)


def both():
    rows_per = d_out.shape[0]
    with torch.cuda.stream(s_in):
        for a in range(0, total, chunk):
            b = min(total, a + chunk)
            d_buf[: b - a].copy_(host_sig[a:b], non_blocking=True)
    with torch.cuda.stream(s_out):
        for r in range(0, frames, rows_per):
            e = min(frames, r + rows_per)
            host_out[r:e].copy_(d_out[: e - r], non_blocking=True)


timed("H2D only, 256 MB chunks", h2d_only)
timed("H2D + D2H concurrently", both)
packed = PackedSignals(host_sig.numpy(), offsets, lengths)
for cs in (1 << 26, 1 << 25, 1 << 24, 1 << 23):
    pipe = FeaturePipeline(computer, chunk_samples=cs)
    timed(f"pipeline.run_host chunk={cs}", lambda: pipe.run_host(packed, out=host_out.numpy(), device=dev))
want = computer.compute_batch([host_sig.numpy()[offsets[u]:offsets[u] + lengths[u]] for u in (0, 5000, 9999)])
fo = np.concatenate([[0], np.cumsum([computer.num_frames(int(n)) for n in lengths])])
for w, u in zip(want, (0, 5000, 9999)):
    assert np.array_equal(w, host_out.numpy()[fo[u]:fo[u + 1]]), u
print("pipeline output matches per-utterance compute")
pipe = FeaturePipeline(computer, chunk_samples=chunk)
t0 = time.perf_counter()
for begin, end in pipe._chunks(packed.lengths):
    first = int(packed.offsets[begin]) // 4 * 4
    computer.plan_batch(packed.offsets[begin:end] - first, packed.lengths[begin:end], dev)
torch.cuda.synchronize()
print(f"planning all chunks (host + tile upload): {(time.perf_counter() - t0) * 1e3:.1f} ms")
