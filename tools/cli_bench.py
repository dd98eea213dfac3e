"""Throughput of ``signals-to-torch-feat-dir`` as users run it: wav files on disk in, ``.pt`` files out.

    python tools/cli_bench.py [n_ranks] [n_utts]

Writes `n_utts` synthetic 16-bit wav files (2-20 s, 16 kHz) to a temporary directory, runs the
command with the README fbank config under ``torch.distributed.run`` on `n_ranks` GPUs (1 = plain
call), and prints one JSON line: files/s and audio-hours/s including decoding and ``torch.save``,
per rank (``--report``) and for the whole job (wall clock of the launch, interpreter start-up
included), plus a check that a sample of the files equals a single-utterance computation."""
import json
import os
import subprocess
import sys
import tempfile
import time
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CFG = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
       "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}


def main():
    n_ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        wav_dir, feat_dir = os.path.join(tmp, "wav"), os.path.join(tmp, "feat")
        os.makedirs(wav_dir)
        total = 0
        with open(os.path.join(tmp, "map"), "w") as mp:
            for i in range(n_utts):
                n = int(16000 * rng.uniform(2, 20))
                total += n
                path = os.path.join(wav_dir, f"u{i:05d}.wav")
                with wave.open(path, "wb") as wv:
                    wv.setnchannels(1)
                    wv.setsampwidth(2)
                    wv.setframerate(16000)
                    wv.writeframes(rng.integers(-20000, 20000, n).astype(np.int16).tobytes())
                mp.write(f"u{i:05d} {path}\n")
        report = os.path.join(tmp, "report.jsonl")
        args = [os.path.join(tmp, "map"), json.dumps(CFG), feat_dir, "--num-workers=8", f"--report={report}"]
        module = [sys.executable, "-m", "pydrobert_speech_b200.command_line", "signals-to-torch-feat-dir"]
        if n_ranks > 1:
            module = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}",
                      "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "pydrobert_speech_b200.command_line",
                      "signals-to-torch-feat-dir"]
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
        t0 = time.perf_counter()
        subprocess.check_call(module + args, env=env, cwd=ROOT)
        wall = time.perf_counter() - t0
        with open(report) as handle:
            ranks = [json.loads(line) for line in handle]
        assert len(os.listdir(feat_dir)) == n_utts
        import torch

        import pydrobert_speech_b200 as pds

        computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, CFG)
        for i in (0, n_utts // 2, n_utts - 1):
            with wave.open(os.path.join(wav_dir, f"u{i:05d}.wav")) as wv:
                pcm = np.frombuffer(wv.readframes(wv.getnframes()), dtype="<i2")
            want = computer.compute_full(pcm.astype(np.float32))
            got = torch.load(os.path.join(feat_dir, f"u{i:05d}.pt")).numpy()
            assert got.shape == want.shape and np.array_equal(got, want), i
        hours = total / 16000 / 3600
        slowest = max(r["seconds"] for r in ranks)
        print(json.dumps({
            "command": "signals-to-torch-feat-dir (README fbank, 16-bit wav in, .pt out, tmpfs)",
            "n_ranks": n_ranks, "utterances": n_utts, "audio_hours": hours,
            "job_files_per_second": n_utts / slowest, "job_audio_hours_per_second": hours / slowest,
            "launch_wall_seconds": wall, "slowest_rank_seconds": slowest,
            "per_rank_audio_hours_per_second": [round(r["audio_hours_per_second"], 2) for r in ranks],
            "host_cores": os.cpu_count(),
        }))


if __name__ == "__main__":
    main()
