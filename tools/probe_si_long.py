"""Timing probe of short integration with long supports (development aid): the triangular mel
bank (6 987 taps) on the long-support overlap-save kernel against the time-domain kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydrobert_speech_b200 as pds  # noqa: E402
from pydrobert_speech_b200.compute import PackedSignals  # noqa: E402

dev = torch.device("cuda", 0)
banks = {
    "fbank40 (6 987 taps)": "fbank",
    "fbank3 (wide filters)": {"name": "fbank", "num_filts": 3},
    "gabor128 (838 taps)": {"name": "gabor", "scaling_function": "mel", "num_filts": 128},
    "gammatone100 (934 taps)": {"name": "gammatone", "scaling_function": "mel", "num_filts": 100},
}
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
lengths = np.array([16000 * 60] * (seconds // 60), dtype=np.int64)
offsets, total = PackedSignals.layout(lengths, 0)
d_sig = torch.randn(total, device=dev) * 1000
hours = lengths.sum() / 16000 / 3600
for label, bank in banks.items():
    for kernel in ("", "unpaired", "direct"):
        os.environ.pop("PDS_SI_KERNEL", None)
        os.environ.pop("PDS_SI_PAIRS", None)
        if kernel == "unpaired":
            os.environ["PDS_SI_PAIRS"] = "0"
        elif kernel:
            os.environ["PDS_SI_KERNEL"] = kernel
        try:
            si = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, {"name": "si", "bank": bank})
            out = si.compute_packed_device(d_sig, offsets, lengths)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(2):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                si.compute_packed_device(d_sig, offsets, lengths)
                t1.record()
                torch.cuda.synchronize()
                best = min(best, t0.elapsed_time(t1))
            print(f"{label:26s} kernel={kernel or 'default':8s} {best:9.3f} ms  {hours / (best * 1e-3):8.2f} audio-h/s")
        except Exception as exc:  # the time-domain kernel does not fit every geometry
            print(f"{label:26s} kernel={kernel or 'default':8s} unavailable: {exc}")
