"""Device times of the fused Deltas(2) -> CMVN kernels on the benchmark's feature matrix
(10.99 M x 41, 10 000 utterances): statistics pass and apply pass (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.post import Deltas, Standardize  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
frames = ((16000 * rng.uniform(2, 20, 10000)).astype(np.int64) + 80) // 160
row_off = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)])).to(dev)
rows = int(frames.sum())
feats = torch.randn((rows, 41), device=dev)
out = torch.empty((rows, 123), device=dev)
deltas = Deltas(2)


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


def stats():
    Standardize().accumulate_device(deltas.lazy_device(feats, row_off))


cmvn = Standardize()
cmvn.accumulate_device(deltas.lazy_device(feats, row_off))


def apply():
    cmvn.apply_device(deltas.lazy_device(feats, row_off), out=out)


a, b = timeit(stats), timeit(apply)
print(f"{os.environ.get('PDS_DELTAS_KERNEL', 'streaming'):9s} run={os.environ.get('PDS_D25_RUN', '-'):>3s} grid={os.environ.get('PDS_D25_GRID', '-'):>2s}  "
      f"stats {a:.3f} ms ({rows * 41 * 4 / a / 1e6:.0f} GB/s)  apply {b:.3f} ms ({rows * 164 * 4 / b / 1e6:.0f} GB/s)  sum {a + b:.3f} ms")
