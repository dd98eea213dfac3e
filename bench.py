#!/usr/bin/env python
"""Benchmark of the frame-feature hot path (BASELINE.json: audio-hours/sec for 40-mel fbank).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (``config.workload``): BASELINE.json configs[1] -- the README fbank config (STFT,
40 triangular mel filters, 25 ms / 10 ms, Hann, power, log, energy) over 10 000 synthetic 16 kHz
utterances of 2-20 s (float32 ``N(0,1) * 1000`` samples, lengths from ``default_rng(rank)``):
30.5 audio-hours, 7.0 GB of samples per GPU.  With N GPUs every rank owns such a corpus shard of
its own (weak scaling; utterances are independent).

One *step* = one pass of the whole shard through the fused STFT kernel.

* ``value``     : whole-job audio-hours per second with the packed samples and the tile table
                  already resident in HBM; CUDA events on the launching stream, max over ranks.
* ``e2e``       : the same metric through ``FeaturePipeline.run_host`` with pinned HOST buffers,
                  the host->device copy of all samples and the device->host copy of all features
                  inside the timed region (float32 samples; ``int16_pcm_input`` repeats it with
                  16-bit PCM host buffers, what wav files hold).
* ``roofline``  : the fused kernel against its binding bound, the FP32 pipe (SURVEY.md 8(d):
                  14 559 flop/frame against SMs * 128 lanes * 2 * clock); the HBM figures
                  (804 B/frame against MEASURED_PEAKS.json) are reported beside it.
* ``sustained`` : the same kernel launched back to back for >= 3 s, with its own clock record.
* ``per_config``: the other BASELINE configs, each timed the same way (CUDA events, clocks sampled
                  inside the region): ``c3`` gammatone-64 STFT, ``c4`` short-integration Gabor-41
                  over 100 x 60 s + 10 x 600 s, ``c5`` fbank + Deltas(2) + corpus CMVN -- with
                  N > 1 ranks the c5 step contains the NCCL all-reduce of the statistics
                  (``collective_us`` times it alone).
* ``cpu_baseline``: the reference's own NumPy path (``oracle/_ref``, installed by
                  ``__graft_entry__.build()`` where ``/root/reference`` exists) on all host cores,
                  on a bounded sample generated outside the timed region; ``cpu_baseline_port``
                  is the vectorised float64 port (``oracle/``) on the same sample.

``--impl reference`` times the CPU arm alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

README_FBANK = {
    "name": "stft",
    "bank": "fbank",
    "frame_length_ms": 25,
    "include_energy": True,
    "pad_to_nearest_power_of_two": True,
    "window_function": "hanning",
    "use_power": True,
}
GAMMATONE_64 = {
    "name": "stft",
    "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 64, "erb": True},
    "frame_length_ms": 25,
    "use_power": True,
}
SI_GABOR_41 = {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}}
N_UTTS = 10000
RATE = 16000
BYTES_PER_FRAME = 804  # 4*S + 4*num_coeffs
FLOPS_PER_FRAME = 14559
METRIC = "audio-hours/sec, 40-mel fbank (README config), synthetic 16 kHz corpus"


def corpus_lengths(seed: int, n_utts: int = N_UTTS) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (RATE * rng.uniform(2, 20, n_utts)).astype(np.int64)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as handle:
            return json.load(handle), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def stft_flops_per_frame(computer) -> float:
    """2.5 N log2 N + L (+ 2 L energy) + 3 K + 2 nnz(W) + 2 C"""
    n = computer._dft_size
    k = n // 2 + 1
    length = computer.frame_length
    return (2.5 * n * np.log2(n) + length + (2 * length if computer._include_energy else 0) + 3 * k
            + 2 * computer.folded_weights.nnz + 2 * computer.num_coeffs)


def si_flops_per_frame(computer, n_fft: int = 1024) -> float:
    """SURVEY.md 8(d), overlap-save accounting: per block of V = N - M + 1 valid samples
    2.5 N log2 N + F (6 N + 5 N log2 N) + 7 V F"""
    filts = computer.num_coeffs
    valid = n_fft - computer._max_support + 1
    per_block = 2.5 * n_fft * np.log2(n_fft) + filts * (6 * n_fft + 5 * n_fft * np.log2(n_fft)) + 7 * valid * filts
    return per_block / valid * computer.frame_shift


# ------------------------------------------------------------------------------------------
# CPU arms: the real reference (oracle/_ref) and the NumPy port (oracle/)
# ------------------------------------------------------------------------------------------
_WORKER = {}
_SIGNALS = []  # generated in the parent BEFORE the pool forks, outside every timed region


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "pydrobert", "speech"))


def _cpu_init(kind):
    try:  # one BLAS/OpenMP thread per worker process; the pool supplies the parallelism
        import torch

        torch.set_num_threads(1)
    except Exception:
        pass
    if kind == "reference":
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from pydrobert.speech.alias import alias_factory_subclass_from_arg
        from pydrobert.speech.compute import FrameComputer

        _WORKER["computer"] = alias_factory_subclass_from_arg(FrameComputer, dict(README_FBANK))
    else:
        import oracle  # noqa: F401  (bench.py is one of the places allowed to use the oracle)
        import pydrobert_speech_b200 as pds

        computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, README_FBANK)
        _WORKER["args"] = (
            computer._window, computer._dft_size, computer._filt_start_idxs, computer._truncated_filts,
            computer.frame_shift, computer.pad_left, True, True, True, True,
        )
    _WORKER["kind"] = kind


def _cpu_one(index):
    signal = _SIGNALS[index]
    if _WORKER["kind"] == "reference":
        return _WORKER["computer"].compute_full(signal).shape[0]
    import oracle

    return oracle.stft_features(signal, *_WORKER["args"]).shape[0]


def make_cpu_sample(n_utts):
    """float64 signals of the first `n_utts` utterances of the workload, into the module global the
    forked workers inherit"""
    lengths = corpus_lengths(0)[:n_utts]
    del _SIGNALS[:]
    for i, n in enumerate(lengths):
        _SIGNALS.append(np.random.default_rng(1000 + i).standard_normal(int(n)) * 1000.0)
    return lengths


def cpu_rate(kind, n_utts, cores, passes=1, warm_passes=0):
    """(audio-hours/sec, seconds per pass) of a CPU arm on `cores` processes over _SIGNALS[:n_utts]"""
    import multiprocessing as mp

    hours = sum(len(s) for s in _SIGNALS[:n_utts]) / RATE / 3600.0
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(kind,)) as pool:
        pool.map(_cpu_one, range(min(cores, n_utts)))  # warm the workers (imports, tables)
        for _ in range(warm_passes):
            pool.map(_cpu_one, range(n_utts), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(passes):
            pool.map(_cpu_one, range(n_utts), chunksize=1)
        elapsed = (time.perf_counter() - t0) / passes
    return hours / elapsed, elapsed


def reference_sample_size(cores):
    # the reference runs ~30x real time per core: 15 utterances (~165 s of audio) per core and pass
    return int(min(N_UTTS, max(16, 15 * cores)))


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = "reference" if have_reference() else "port"
    n_utts = reference_sample_size(cores) if kind == "reference" else int(min(N_UTTS, max(64, 160 * cores)))
    lengths = make_cpu_sample(n_utts)
    hours = float(lengths.sum()) / RATE / 3600.0
    value, secs = cpu_rate(kind, n_utts, cores, passes=args.steps, warm_passes=args.warmup)
    what = ("the reference itself (pydrobert.speech STFTFrameComputer.compute_full, NumPy float64, installed "
            "into oracle/_ref)" if kind == "reference" else
            "float64 NumPy restatement of compute.py:388-460/574-607 vectorised over frames (oracle/stft.py)")
    sample = (f"{n_utts} utterances ({hours:.3f} audio-h) of the workload per step, signals generated outside "
              f"the timed region; {what}; one process per core x {cores}")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "audio-hours/sec",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": secs * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: README fbank over synthetic 2-20 s 16 kHz utterances (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/sec", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in the background for the whole run; `window()` summarises the samples taken
    inside one timed region"""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.file = None
        self.rows = []
        self.sm_max = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def _read(self):
        import datetime

        if self.file is None:
            return
        self.file.flush()
        self.file.seek(0)
        rows = []
        for row in self.file.read().splitlines():
            cells = [c.strip() for c in row.split(",")]
            if len(cells) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(cells[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clock = float(cells[1])
                self.sm_max = float(cells[2])
                power = float(cells[3])
            except ValueError:
                continue
            active = [name for name, cell in zip(self.NAMES, cells[5:9]) if cell.lower().startswith("active")]
            rows.append((stamp, clock, active, power))
        self.rows = rows

    def window(self, t_begin, t_end):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self._read()
        rows = self.rows
        inside = [r for r in rows if t_begin - 0.02 <= r[0] <= t_end + 0.02]
        where = "timed region"
        if not inside and rows:  # region shorter than the sampling period: nearest samples under the same load
            mid = 0.5 * (t_begin + t_end)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            where = "nearest to the timed region (same load)"
        if inside:
            out["sm_mhz"] = float(np.median([r[1] for r in inside]))
            out["sm_max_mhz"] = self.sm_max
            out["power_w"] = float(np.median([r[3] for r in inside]))
            out["samples"] = len(inside)
            out["sampled"] = where
            out["reasons"] = sorted({name for r in inside for name in r[2]})
        return out

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.close()
        os.unlink(self.file.name)
        self.proc = None


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import pydrobert_speech_b200 as pds
    from pydrobert_speech_b200.compute import PackedSignals
    from pydrobert_speech_b200.pipeline import FeaturePipeline
    from pydrobert_speech_b200.post import Deltas, Standardize

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.current_stream(device)
    peaks, peak_src = measured_peaks()
    prop = torch.cuda.get_device_properties(device)
    fp32_peak = prop.multi_processor_count * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    hbm_peak = float(peaks["hbm_gbs"])

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def timed(step, steps, warmup):
        """W warm-up steps, then K steps between barriers; returns (ms per step as the max over
        ranks, mean per-step device time on this rank, clock record of the region)"""
        for _ in range(warmup):
            step()
        barrier()
        wall_begin = time.time()
        begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        begin.record(stream)
        for _ in range(steps):
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            step()
            k1.record(stream)
            marks.append((k0, k1))
        end.record(stream)
        barrier()
        clocks = sampler.window(wall_begin, time.time()) if rank == 0 else None
        total_ms = reduce_max(begin.elapsed_time(end))
        step_ms = float(np.mean([a.elapsed_time(b) for a, b in marks]))
        return total_ms / steps, step_ms, clocks

    # ---- headline: configs[1] ---------------------------------------------------------------
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, README_FBANK)
    lengths = corpus_lengths(rank, args.utts)  # every rank: its own shard of the same size class
    offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    d_signal = torch.randn(total, device=device, generator=gen) * 1000.0
    audio_hours = float(lengths.sum()) / RATE / 3600.0
    layout = computer.plan_batch(offsets, lengths, device)
    frames = layout.rows
    d_feats = torch.empty((frames, computer.num_coeffs), dtype=torch.float32, device=device)
    job_hours = reduce_sum(audio_hours)

    def stft_step():
        computer.run_batch(layout, d_signal, out=d_feats)

    # a fresh box idles at 120 MHz: spin the same step (untimed) until the clocks have ramped
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < args.prewarm_s:
        stft_step()
        torch.cuda.synchronize(device)
    ms_per_step, kernel_ms, clocks = timed(stft_step, args.steps, args.warmup)
    value = job_hours / (ms_per_step * 1e-3)

    def stft_roofline(comp, n_frames, k_ms, bytes_per_frame):
        flops = stft_flops_per_frame(comp)
        tf = n_frames * flops / (k_ms * 1e-3) / 1e12
        gbs = n_frames * bytes_per_frame / (k_ms * 1e-3) / 1e9
        return {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                "flops_per_frame": flops, "hbm_achieved_gbs": gbs, "hbm_frac": gbs / hbm_peak}

    # ---- sustained: the same launch back to back for >= 3 s --------------------------------
    sustained = None
    if args.sustain_s > 0:
        n_launch = max(args.steps, int(np.ceil(args.sustain_s * 1e3 / kernel_ms)))
        s_ms, s_kernel_ms, s_clocks = timed(stft_step, n_launch, 0)
        sustained = {"value": job_hours / (s_ms * 1e-3), "unit": "audio-hours/sec", "launches": n_launch,
                     "seconds": s_ms * n_launch * 1e-3, "ms_per_step": s_ms, "clocks": s_clocks}

    # ---- the other BASELINE configs, timed the same way -------------------------------------
    per_config = {}
    if not args.no_configs:
        side_steps = max(3, min(args.steps, 10))
        # c5: fbank + Deltas(2) + corpus CMVN; the statistics never leave the GPU: accumulate ->
        # (all-reduce over NCCL) -> apply are stream ordered, no host sync inside a step
        deltas = Deltas(2)
        d_row_off = torch.from_numpy(layout.frame_off).to(device)
        d_full = torch.empty((frames, 3 * computer.num_coeffs), dtype=torch.float32, device=device)

        def c5_step():
            computer.run_batch(layout, d_signal, out=d_feats)
            lazy = deltas.lazy_device(d_feats, d_row_off)  # the Deltas output is never materialised
            cmvn = Standardize()
            cmvn.accumulate_device(lazy)
            if world > 1:
                cmvn.allreduce()
            cmvn.apply_device(lazy, out=d_full)

        c5_ms, c5_kernel_ms, c5_clocks = timed(c5_step, side_steps, 3)
        c5_bytes = 4 * computer.frame_shift + 4 * 123 + 2 * 4 * 123  # 2 116 B/frame
        c5_gbs = frames * c5_bytes / (c5_kernel_ms * 1e-3) / 1e9
        collective_us = None
        if world > 1:
            probe = Standardize()
            probe.device_stats(device, 123)

            def coll_step():
                probe.allreduce()

            coll_ms, _, _ = timed(coll_step, 50, 5)
            collective_us = coll_ms * 1e3
        per_config["c5"] = {
            "workload": "configs[4]: configs[1] + Deltas(2) along time + corpus Standardize, statistics "
                        + (f"all-reduced over NCCL across {world} ranks" if world > 1 else "of one rank (no collective at N=1)"),
            "value": job_hours / (c5_ms * 1e-3), "unit": "audio-hours/sec", "ms_per_step": c5_ms, "kernel_ms": c5_kernel_ms,
            "steps": side_steps, "launches_per_step": 3, "collective_us": collective_us,
            "roofline": {"bound": "hbm", "achieved": c5_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": c5_gbs / hbm_peak,
                         "bytes_per_frame": c5_bytes,
                         "note": "the chain's first stage (the fused STFT kernel) is FP32 bound; see roofline of the headline"},
            "clocks": c5_clocks,
        }
        del d_full
        # c3: gammatone-64 STFT on the same corpus
        comp3 = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, GAMMATONE_64)
        off3, total3 = PackedSignals.layout(lengths, comp3.pad_left % 4)
        if total3 <= total and np.array_equal(off3, offsets):
            sig3 = d_signal
        else:
            sig3 = torch.randn(total3, device=device, generator=gen) * 1000.0
        layout3 = comp3.plan_batch(off3, lengths, device)
        out3 = torch.empty((layout3.rows, comp3.num_coeffs), dtype=torch.float32, device=device)

        def c3_step():
            comp3.run_batch(layout3, sig3, out=out3)

        c3_ms, c3_kernel_ms, c3_clocks = timed(c3_step, side_steps, 3)
        per_config["c3"] = {
            "workload": "configs[2]: STFT + complex gammatone bank (64 filters, mel centres, ERB bandwidths), 512-point "
                        "DFT, log power, same corpus",
            "value": job_hours / (c3_ms * 1e-3), "unit": "audio-hours/sec", "ms_per_step": c3_ms, "kernel_ms": c3_kernel_ms,
            "steps": side_steps, "launches_per_step": 1,
            "roofline": stft_roofline(comp3, layout3.rows, c3_kernel_ms, 4 * comp3.frame_shift + 4 * comp3.num_coeffs),
            "clocks": c3_clocks,
        }
        del out3, sig3, layout3
        # c4: short-integration Gabor-41 over long utterances
        comp4 = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, SI_GABOR_41)
        len4 = np.array([RATE * 60] * 100 + [RATE * 600] * 10, dtype=np.int64)
        off4, total4 = PackedSignals.layout(len4, 0)
        sig4 = torch.randn(total4, device=device, generator=gen) * 1000.0
        hours4 = float(len4.sum()) / RATE / 3600.0
        frames4 = int(sum(comp4.num_frames(int(n)) for n in len4))

        def c4_step():
            comp4.compute_packed_device(sig4, off4, len4)

        c4_steps = max(3, min(args.steps, 5))
        c4_ms, c4_kernel_ms, c4_clocks = timed(c4_step, c4_steps, 3)
        flops4 = si_flops_per_frame(comp4)
        tf4 = frames4 * flops4 / (c4_kernel_ms * 1e-3) / 1e12
        per_config["c4"] = {
            "workload": "configs[3]: SIFrameComputer + Gabor bank (41 filters, mel), pooling over 100 x 60 s + 10 x 600 s "
                        "utterances per GPU",
            "value": reduce_sum(hours4) / (c4_ms * 1e-3), "unit": "audio-hours/sec", "ms_per_step": c4_ms,
            "kernel_ms": c4_kernel_ms, "steps": c4_steps, "audio_hours_per_gpu": hours4,
            "roofline": {"bound": "fp32", "achieved": tf4, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf4 / fp32_peak,
                         "flops_per_frame": flops4,
                         "note": "overlap-save accounting of SURVEY.md 8(d) (1024-point blocks)"},
            "clocks": c4_clocks,
        }
        del sig4

    # ---- end to end through the public pipeline API: host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pipeline = FeaturePipeline(computer, chunk_samples=args.chunk_samples)
        host_sig = torch.empty(total, dtype=torch.float32, pin_memory=True)
        host_sig.copy_(d_signal)
        host_out = torch.empty((frames, computer.num_coeffs), dtype=torch.float32, pin_memory=True)
        packed = PackedSignals(host_sig.numpy(), offsets, lengths)
        del d_signal, d_feats
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))

        def e2e_rate(batch):
            pipeline.run_host(batch, out=host_out.numpy(), device=device)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pipeline.run_host(batch, out=host_out.numpy(), device=device)
            torch.cuda.synchronize(device)
            return job_hours * e2e_steps / reduce_max(time.perf_counter() - t0)

        e2e_value = e2e_rate(packed)

        def copy_floor(host_in):
            """The same bytes moved with no kernel in between: all samples host->device and all
            features device->host, chunked like run_host, on two streams, all ranks at once"""
            chunk = args.chunk_samples
            d_in = torch.empty(chunk, dtype=host_in.dtype, device=device)
            d_out = torch.empty((min(frames, chunk // computer.frame_shift + 64), computer.num_coeffs),
                                dtype=torch.float32, device=device)
            s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

            def once():
                with torch.cuda.stream(s_in):
                    for first in range(0, total, chunk):
                        n = min(chunk, total - first)
                        d_in[:n].copy_(host_in[first:first + n], non_blocking=True)
                with torch.cuda.stream(s_out):
                    for first in range(0, frames, d_out.shape[0]):
                        n = min(d_out.shape[0], frames - first)
                        host_out[first:first + n].copy_(d_out[:n], non_blocking=True)
                s_in.synchronize()
                s_out.synchronize()

            once()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                once()
            return job_hours * e2e_steps / reduce_max(time.perf_counter() - t0)

        floor_f32 = copy_floor(host_sig)
        e2e_pcm = None
        if not args.no_pcm:  # the same corpus as 16-bit PCM (what wav files hold): half the bytes over PCIe
            host_pcm = torch.empty(total, dtype=torch.int16, pin_memory=True)
            host_pcm.copy_(host_sig.clamp(-32768, 32767).round_().to(torch.int16))
            del host_sig
            packed_pcm = PackedSignals(host_pcm.numpy(), offsets, lengths)
            pcm_value = e2e_rate(packed_pcm)
            e2e_pcm = {"value": pcm_value, "unit": "audio-hours/sec", "copy_floor": copy_floor(host_pcm),
                       "h2d_bytes_per_step": int(lengths.sum()) * 2 + layout.n_tiles * 32,
                       "note": "same corpus rounded to int16 PCM host buffers (supplementary; `e2e` is the float32 run)"}
        e2e = {
            "value": e2e_value,
            "unit": "audio-hours/sec",
            "h2d_bytes_per_step": int(lengths.sum()) * 4 + layout.n_tiles * 32,
            "d2h_bytes_per_step": int(frames) * computer.num_coeffs * 4,
            "steps": e2e_steps,
            "copy_floor": floor_f32,
            "copy_floor_note": "the same host->device and device->host bytes with no kernel in between, all ranks at "
                               "once (what the box's PCIe / host memory allows); e2e / copy_floor is the pipeline's efficiency",
            "api": "pydrobert_speech_b200.pipeline.FeaturePipeline.run_host (pinned host in/out)",
            "int16_pcm_input": e2e_pcm,
        }

    if rank == 0:
        sampler.stop()
        roofline = stft_roofline(computer, frames, kernel_ms, BYTES_PER_FRAME)
        roofline.update({
            "traffic": None,
            "peak_source": f"FP32 = {prop.multi_processor_count} SMs x 128 lanes x 2 x {peaks.get('sm_max_mhz', 1965.0):.0f} MHz "
                           f"(clock from MEASURED_PEAKS.json); HBM {hbm_peak:.0f} GB/s {peak_src}",
            "kernel": computer.kernel_name(device) + " (dft_size 512, power, float32 samples)",
            "kernel_ms": kernel_ms,
            "algorithmic_flops_per_launch": float(frames) * FLOPS_PER_FRAME,
            "algorithmic_bytes_per_launch": int(frames * BYTES_PER_FRAME),
            "note": "FP32 is the binding bound (SURVEY.md 8(d): 14.2 k audio-h/s against 22.6 k for HBM)",
        })
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as handle:
                per_frame = json.load(handle).get("dram_bytes_per_frame")
            if per_frame:
                roofline["traffic"] = per_frame * frames
        cores = os.cpu_count() or 1
        cpu = cpu_port = None
        if world == 1 and not args.no_cpu:
            n_ref = reference_sample_size(cores)
            n_port = int(min(N_UTTS, max(64, 160 * cores)))
            sample_lengths = make_cpu_sample(max(n_ref, n_port))
            rate, secs = cpu_rate("port", n_port, cores)
            cpu_port = {
                "value": rate, "unit": "audio-hours/sec", "cores": cores, "kind": "port",
                "sample": (f"first {n_port} utterances of the workload "
                           f"({float(sample_lengths[:n_port].sum()) / RATE / 3600:.3f} audio-h, {secs:.1f} s wall), generated "
                           "outside the timed region; float64 NumPy port vectorised over frames (oracle/stft.py), "
                           "one process per core"),
            }
            cpu = cpu_port
            if have_reference():
                rate, secs = cpu_rate("reference", n_ref, cores)
                cpu = {
                    "value": rate, "unit": "audio-hours/sec", "cores": cores, "kind": "reference",
                    "sample": (f"first {n_ref} utterances of the workload "
                               f"({float(sample_lengths[:n_ref].sum()) / RATE / 3600:.3f} audio-h, {secs:.1f} s wall), generated "
                               "outside the timed region; the reference's own STFTFrameComputer.compute_full (NumPy "
                               "float64, oracle/_ref), one process per core"),
                }
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "audio-hours/sec",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": ("configs[1]: README fbank (STFT N=512, 40 mel + energy, 25/10 ms, Hann, power, "
                             f"log) over {args.utts} synthetic 16 kHz utterances of 2-20 s per GPU"),
                "audio_hours_per_gpu": audio_hours,
                "frames_per_gpu": int(frames),
                "input_bytes_per_gpu": int(lengths.sum()) * 4,
                "l2_policy": "inputs (7 GB) and outputs (1.8 GB) far exceed the 126 MB L2; no flush needed",
                "parallelism": f"utterance shards, {world} rank(s); the only data-path collective is the CMVN "
                               "all-reduce of per_config.c5",
            },
            "roofline": roofline,
            "sustained": sustained,
            "per_config": per_config,
            "cpu_baseline": cpu,
            "cpu_baseline_port": cpu_port,
            "e2e": e2e,
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=("ours", "reference"), default="ours")
    parser.add_argument("--utts", type=int, default=N_UTTS, help="utterances per GPU")
    parser.add_argument("--chunk-samples", type=int, default=1 << 26)
    parser.add_argument("--e2e-steps", type=int, default=5)
    parser.add_argument("--sustain-s", type=float, default=3.0,
                        help="seconds of back-to-back launches for the `sustained` figure (0 = skip)")
    parser.add_argument("--no-pcm", action="store_true", help="skip the supplementary int16 end-to-end run")
    parser.add_argument("--prewarm-s", type=float, default=0.75,
                        help="seconds of untimed launches before the warm-up steps (clock ramp on a cold GPU)")
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--no-cpu", action="store_true")
    parser.add_argument("--no-configs", action="store_true", help="skip per_config (c3, c4, c5)")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
