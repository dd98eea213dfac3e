"""A/B timing of the fused STFT kernel variants (PDS_STFT_KERNEL values) on the benchmark corpus.

    python tools/probe_variants.py [n_utts] [variant ...]      e.g.  10000 1 t

Every variant gets a fresh plan; outputs are compared with the first variant's.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

cfg = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
variants = sys.argv[2:] or ["1", "2", "p"]
rng = np.random.default_rng(0)
lengths = (16000 * rng.uniform(2, 20, n_utts)).astype(np.int64)
dev = torch.device("cuda", 0)
reference = None
for variant in variants:
    os.environ["PDS_STFT_KERNEL"] = variant.split(":")[0]
    os.environ["PDS_W_PROBE"] = variant.split(":")[1] if ":" in variant else "0"
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
    offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
    gen = torch.Generator(device=dev).manual_seed(0)
    d_sig = torch.randn(total, device=dev, generator=gen) * 1000
    layout = computer.plan_batch(offsets, lengths, dev)
    feats = torch.empty((layout.rows, computer.num_coeffs), device=dev)
    for rep in range(5):
        computer.run_batch(layout, d_sig, out=feats)
    torch.cuda.synchronize()
    times = []
    for rep in range(10):
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        computer.run_batch(layout, d_sig, out=feats)
        t1.record()
        torch.cuda.synchronize()
        times.append(t0.elapsed_time(t1))
    frames = layout.rows
    ms = float(np.median(times))
    note = ""
    if reference is None:
        reference = feats.clone()
    else:
        diff = (feats - reference).abs().max().item()
        note = f" max|diff vs {variants[0]}|={diff:.2e}"
    print(f"variant={variant!r} utts={n_utts} frames={frames} median_ms={ms:.3f} min_ms={min(times):.3f} "
          f"Gframes/s={frames / ms / 1e6:.3f} fp32_frac={frames * 14559 / (ms * 1e-3) / 74.45e12:.3f}{note}", flush=True)
    del computer, layout, feats, d_sig
