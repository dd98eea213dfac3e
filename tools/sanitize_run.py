"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200 import pipeline as pds_pipeline  # noqa: E402

rng = np.random.default_rng(0)
lengths = [0, 1, 201, 399, 5000, 16000, 33333, 160 * 32 + 240, 48000]
signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
fbank = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
         "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
for cfg in (fbank, dict(fbank, frame_length_ms=64, frame_shift_ms=16), dict(fbank, frame_length_ms=100, frame_shift_ms=20),
            dict(fbank, frame_shift_ms=10.0625),
            {"name": "stft", "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 64, "erb": True},
             "frame_length_ms": 25, "use_power": True}):
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
    # default kernels, then every opt-in variant (the plan cache is keyed on the switches)
    for kernel, bank in ((None, None), ("1", None), (None, "tf32"), ("p", None), ("scalar", None), ("ws", None)):
        for name, value in (("PDS_STFT_KERNEL", kernel), ("PDS_STFT_BANK", bank)):
            os.environ.pop(name, None)
            if value is not None:
                os.environ[name] = value
        feats = computer.compute_batch(signals)
        feats = computer.compute_batch(signals, preemph=0.97, dither=1.0, seed=3)
        feats = computer.compute_batch([s.astype(np.int16) for s in signals])
        print("stft ok", computer._dft_size, kernel, bank, computer.kernel_name(), sum(len(f) for f in feats))
    os.environ.pop("PDS_STFT_KERNEL", None)
    os.environ.pop("PDS_STFT_BANK", None)
# DFT sizes that are not a power of two: Bluestein (<= 512) and the direct-DFT kernel
for ms in (25, 20, 12.5):
    cfg = dict(fbank, frame_length_ms=ms, pad_to_nearest_power_of_two=False)
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cfg)
    feats = computer.compute_batch(signals)
    print("stft ok", computer._dft_size, computer.kernel_name(), sum(len(f) for f in feats))
si = pds.alias_factory_subclass_from_arg(
    pds.compute.FrameComputer, {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}})
out = si.compute_batch(signals)
print("si ok", sum(len(f) for f in out))
dev = torch.device("cuda", 0)
x = torch.randn(3000, 41, device=dev)
row_off = torch.tensor([0, 7, 1000, 1001, 3000], dtype=torch.int64, device=dev)
d = pds.post.Deltas(2)
full = d.apply_device(x, row_off)
cm = pds.post.Standardize()
lazy = d.lazy_device(x, row_off)
cm.accumulate_device(lazy)
y = cm.apply_device(lazy)
z = pds.post.Deltas(3, context_window=3).apply_device(x, row_off)
os.environ["PDS_DELTAS_KERNEL"] = "s"  # the shared-memory-staged Deltas(2) kernels
full_s = d.apply_device(x, row_off)
cm2 = pds.post.Standardize()
cm2.accumulate_device(d.lazy_device(x, row_off))
y2 = cm2.apply_device(d.lazy_device(x, row_off))
os.environ.pop("PDS_DELTAS_KERNEL")
torch.cuda.synchronize()
assert torch.equal(full, full_s) and torch.allclose(y, y2, rtol=1e-5, atol=1e-5)
pipe = pds_pipeline.FeaturePipeline(pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, fbank))
packed = pds_pipeline.PackedSignals.pack(signals, np.float32, pipe.computer.pad_left % 4)
corpus, _ = pipe.run_corpus(packed, pds.post.Standardize(), pds.post.Deltas(2))
torch.cuda.synchronize()
print("post ok", tuple(full.shape), tuple(y.shape), tuple(z.shape))
