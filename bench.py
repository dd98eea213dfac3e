#!/usr/bin/env python
"""Benchmark of the frame-feature hot path (BASELINE.json: audio-hours/sec for 40-mel fbank).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]

Workload (``config.workload``): BASELINE.json configs[1] -- the README fbank config (STFT, 40
triangular mel filters, 25 ms / 10 ms, Hann, power, log, energy) over 10 000 synthetic 16 kHz
utterances of 2-20 s (float32 ``N(0,1) * 1000`` samples, lengths from ``default_rng(0)``):
30.6 audio-hours, 7.0 GB of samples per GPU.  With N GPUs every rank owns such a corpus shard of
its own (weak scaling; utterances are independent, there is no data-path collective).

One *step* = one pass of the whole shard through the fused STFT kernel.

* ``value``   : whole-job audio-hours per second with the packed samples and the tile table already
                resident in HBM; timed with CUDA events on the launching stream, max over ranks.
* ``e2e``     : the same metric through ``FeaturePipeline.run_host`` with pinned HOST buffers, the
                host->device copy of all samples and the device->host copy of all features inside
                the timed region.
* ``roofline``: the fused kernel (``stft_tc_kernel``) against the measured HBM copy bandwidth (MEASURED_PEAKS.json);
                algorithmic bytes = 804 B/frame (640 B of new samples + 164 B of coefficients,
                SURVEY.md 8(d)).  The kernel is FP32-issue bound, not HBM bound, so the FP32
                figure (14 559 flop/frame against SMs*128*2*f) is reported beside it.
* ``cpu_baseline``: the float64 NumPy oracle (``oracle/``, a port of the reference's algorithm,
                vectorised over frames) on all host cores, on a bounded subset of the workload.

``--impl reference`` times that CPU port alone (rank 0 only), as the reference arm.
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

README_FBANK = {
    "name": "stft",
    "bank": "fbank",
    "frame_length_ms": 25,
    "include_energy": True,
    "pad_to_nearest_power_of_two": True,
    "window_function": "hanning",
    "use_power": True,
}
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


# ------------------------------------------------------------------------------------------
# CPU port (the oracle) -- used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_init():
    import oracle  # noqa: F401  (bench.py is one of the places allowed to use the oracle)
    import pydrobert_speech_b200 as pds

    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, README_FBANK)
    _WORKER["args"] = (
        computer._window, computer._dft_size, computer._filt_start_idxs, computer._truncated_filts,
        computer.frame_shift, computer.pad_left, True, True, True, True,
    )
    try:  # one BLAS/OpenMP thread per worker process; the pool supplies the parallelism
        import torch

        torch.set_num_threads(1)
    except Exception:
        pass


def _cpu_one(task):
    import oracle

    seed, length = task
    signal = np.random.default_rng(seed).standard_normal(length) * 1000.0
    feats = oracle.stft_features(signal, *_WORKER["args"])
    return feats.shape[0]


def cpu_port_rate(lengths, cores, repeats=1):
    """audio-hours/sec of the NumPy port on `cores` worker processes over the given utterances"""
    import multiprocessing as mp

    tasks = [(1000 + i, int(n)) for i, n in enumerate(lengths)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_one, tasks[: cores])  # warm the workers (imports, table construction)
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_one, tasks, chunksize=1)
            best = min(best, time.perf_counter() - t0)
    hours = float(np.sum(lengths)) / RATE / 3600.0
    return hours / best, best


def cpu_sample_lengths(cores):
    # sized for a few seconds of wall time per pass on all cores (~160 utterances = 0.5 audio-h per core)
    n = int(min(N_UTTS, max(64, 160 * cores)))
    return corpus_lengths(0)[:n]


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    lengths = cpu_sample_lengths(cores)
    import multiprocessing as mp

    tasks = [(1000 + i, int(n)) for i, n in enumerate(lengths)]
    hours = float(np.sum(lengths)) / RATE / 3600.0
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_one, tasks, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one, tasks, chunksize=1)
        elapsed = time.perf_counter() - t0
    value = hours * args.steps / elapsed
    sample = (f"{len(lengths)} utterances ({hours:.3f} audio-h) of the workload per step; float64 NumPy "
              "restatement of compute.py:388-460/574-607 vectorised over frames (oracle/stft.py), "
              f"one process per core x {cores}; /root/reference itself cannot travel to the GPU box")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "audio-hours/sec",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: README fbank over synthetic 2-20 s 16 kHz utterances (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock over the samples taken inside [t_begin, t_end] (wall clock of the timed
        region); the sampler is started before the warm-up so that it is already running."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        rows, sm_max = [], None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.file.read().splitlines():
            cells = [c.strip() for c in row.split(",")]
            if len(cells) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(cells[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clock = float(cells[1])
                sm_max = float(cells[2])
            except ValueError:
                continue
            active = [name for name, cell in zip(names, cells[5:9]) if cell.lower().startswith("active")]
            rows.append((stamp, clock, active))
        self.file.close()
        os.unlink(self.file.name)
        inside = [r for r in rows if t_begin is None or (t_begin - 0.02 <= r[0] <= t_end + 0.02)]
        where = "timed region"
        if not inside and rows:  # region shorter than the sampling period: nearest samples under the same load
            mid = 0.5 * (t_begin + t_end)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            where = "nearest to the timed region (same load)"
        if inside:
            out["sm_mhz"] = float(np.median([r[1] for r in inside]))
            out["sm_max_mhz"] = sm_max
            out["samples"] = len(inside)
            out["sampled"] = where
            out["reasons"] = sorted({name for r in inside for name in r[2]})
        return out


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

    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, README_FBANK)
    lengths = corpus_lengths(rank, args.utts)  # every rank: its own shard of the same size class
    offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    d_signal = torch.randn(total, device=device, generator=gen) * 1000.0
    audio_hours = float(lengths.sum()) / RATE / 3600.0
    layout = computer.plan_batch(offsets, lengths, device)
    frames = layout.rows
    d_feats = torch.empty((frames, computer.num_coeffs), dtype=torch.float32, device=device)
    c5 = args.workload == "c5"
    deltas = Deltas(2)
    d_row_off = torch.from_numpy(layout.frame_off).to(device)
    d_full = torch.empty((frames, 3 * computer.num_coeffs), dtype=torch.float32, device=device) if c5 else None

    def step():
        computer.run_batch(layout, d_signal, out=d_feats)
        if c5:  # fbank + Deltas(2) + corpus CMVN: stats summed per GPU, one allreduce, apply
            lazy = deltas.lazy_device(d_feats, d_row_off)  # Deltas output is never materialised
            cmvn = Standardize()
            cmvn.accumulate_device(lazy)
            if world > 1:
                cmvn.allreduce()
            cmvn.apply_device(lazy, out=d_full)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # a fresh box idles at 120 MHz: spin the same step (untimed) until the clocks have ramped, then
    # do the W warm-up steps the contract asks for
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < args.prewarm_s:
        step()
        torch.cuda.synchronize(device)
    for _ in range(args.warmup):
        step()
    barrier()
    wall_begin = time.time()
    stream = torch.cuda.current_stream(device)
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_events = []
    begin.record(stream)
    for _ in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        computer.run_batch(layout, d_signal, out=d_feats) if not c5 else step()
        k1.record(stream)
        kernel_events.append((k0, k1))
    end.record(stream)
    barrier()
    clocks = sampler.stop(wall_begin, time.time()) if rank == 0 else None
    elapsed_ms = begin.elapsed_time(end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    totals = torch.tensor([audio_hours, float(frames)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    job_hours, job_frames = float(totals[0].item()), float(totals[1].item())
    value = job_hours * args.steps / (elapsed_ms * 1e-3)

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
        pipeline.run_host(packed, out=host_out.numpy(), device=device)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipeline.run_host(packed, out=host_out.numpy(), device=device)
        torch.cuda.synchronize(device)
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # supplementary: the same corpus as 16-bit PCM (what wav files hold): half the bytes over PCIe
        e2e_pcm = None
        if not args.no_pcm and world == 1:  # supplementary, single-GPU runs only (3.5 GB more pinned memory per rank)
            host_pcm = torch.empty(total, dtype=torch.int16, pin_memory=True)
            host_pcm.copy_(host_sig.clamp(-32768, 32767).round_().to(torch.int16))
            packed_pcm = PackedSignals(host_pcm.numpy(), offsets, lengths)
            pipeline.run_host(packed_pcm, out=host_out.numpy(), device=device)  # warm-up
            barrier()
            t1 = time.perf_counter()
            for _ in range(e2e_steps):
                pipeline.run_host(packed_pcm, out=host_out.numpy(), device=device)
            torch.cuda.synchronize(device)
            tp = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            e2e_pcm = {"value": job_hours * e2e_steps / float(tp.item()), "unit": "audio-hours/sec",
                       "h2d_bytes_per_step": int(lengths.sum()) * 2 + layout.n_tiles * 32,
                       "note": "same corpus rounded to int16 PCM host buffers (supplementary; `e2e` is the float32 run)"}
        e2e = {
            "value": job_hours * e2e_steps / float(t.item()),
            "unit": "audio-hours/sec",
            "h2d_bytes_per_step": int(lengths.sum()) * 4 + layout.n_tiles * 32,
            "d2h_bytes_per_step": int(frames) * computer.num_coeffs * 4,
            "steps": e2e_steps,
            "api": "pydrobert_speech_b200.pipeline.FeaturePipeline.run_host (pinned host in/out)",
            "int16_pcm_input": e2e_pcm,
        }

    if rank == 0:
        peaks, peak_src = measured_peaks()
        alg_bytes = frames * BYTES_PER_FRAME
        achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
        prop = torch.cuda.get_device_properties(device)
        fp32_peak = prop.multi_processor_count * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        fp32_achieved = frames * FLOPS_PER_FRAME / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "hbm",
            "achieved": achieved_gbs,
            "peak": float(peaks["hbm_gbs"]),
            "unit": "GB/s",
            "frac": achieved_gbs / float(peaks["hbm_gbs"]),
            "traffic": None,
            "peak_source": peak_src,
            "kernel": "pds::stft_tc_kernel<512, true, float, kRows13>",
            "kernel_ms": kernel_ms,
            "algorithmic_bytes_per_launch": int(alg_bytes),
            "note": "kernel is FP32-issue bound (SURVEY.md 8(d)); see fp32_*",
            "fp32_achieved_tflops": fp32_achieved,
            "fp32_peak_tflops": fp32_peak,
            "fp32_frac": fp32_achieved / fp32_peak,
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as handle:
                per_frame = json.load(handle).get("dram_bytes_per_frame")
            if per_frame:
                roofline["traffic"] = per_frame * frames
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            sample_lengths = cpu_sample_lengths(cores)
            rate, secs = cpu_port_rate(sample_lengths, cores)
            cpu = {
                "value": rate,
                "unit": "audio-hours/sec",
                "cores": cores,
                "kind": "port",
                "sample": (f"first {len(sample_lengths)} utterances of the workload "
                           f"({float(sample_lengths.sum()) / RATE / 3600:.3f} audio-h, {secs:.1f} s wall); "
                           "float64 NumPy port vectorised over frames (oracle/stft.py), one process per core"),
            }
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "audio-hours/sec",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": ("configs[1]: README fbank (STFT N=512, 40 mel + energy, 25/10 ms, Hann, power, "
                             f"log) over {args.utts} synthetic 16 kHz utterances of 2-20 s per GPU"
                             + ("; + Deltas(2) + corpus CMVN (configs[4])" if c5 else "")),
                "audio_hours_per_gpu": audio_hours,
                "frames_per_gpu": int(frames),
                "input_bytes_per_gpu": int(lengths.sum()) * 4,
                "l2_policy": "inputs (7 GB) and outputs (1.8 GB) far exceed the 126 MB L2; no flush needed",
                "parallelism": f"utterance shards, {world} rank(s), no data-path collective",
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * (1 if not c5 else 3),
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
    parser.add_argument("--workload", choices=("c2", "c5"), default="c2")
    parser.add_argument("--utts", type=int, default=N_UTTS, help="utterances per GPU")
    parser.add_argument("--chunk-samples", type=int, default=1 << 26)
    parser.add_argument("--e2e-steps", type=int, default=5)
    parser.add_argument("--no-pcm", action="store_true", help="skip the supplementary int16 end-to-end run")
    parser.add_argument("--prewarm-s", type=float, default=0.75,
                        help="seconds of untimed launches before the warm-up steps (clock ramp on a cold GPU)")
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--no-cpu", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
