"""BASELINE config 5 as a chain (fbank + Deltas(2) + corpus CMVN): the statistics stay on the GPU,
nothing inside a step synchronises the host, and with two ranks the statistics are summed by an
NCCL all-reduce (reference: post.py:160-305 composed by the user)."""
import os
import socket

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu


def _corpus(seed=5, n=23):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal(int(k)) * 800).astype(np.float32) for k in rng.integers(300, 20000, n)]


def _oracle_chain(computer, signals):
    feats = []
    for sig in signals:
        f = oracle.stft_features(sig.astype(np.float64), computer._window, computer._dft_size,
                                 computer._filt_start_idxs, computer._truncated_filts, computer.frame_shift,
                                 computer.pad_left, True, True, True, True)
        feats.append(oracle.deltas(f, 2) if len(f) else np.zeros((0, 123)))
    stacked = np.concatenate(feats)
    mean, std = stacked.mean(0), stacked.std(0)
    return [(f - mean) / std for f in feats], stacked


def test_c5_step_has_no_host_sync(speech):
    """accumulate -> apply on device-resident statistics: torch's sync debug mode raises on any
    synchronising call (.item(), .cpu(), blocking copies) made inside the step"""
    import torch

    from pydrobert_speech_b200.compute import PackedSignals
    from pydrobert_speech_b200.post import Deltas, Standardize

    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.README_FBANK)
    signals = _corpus()
    packed = PackedSignals.pack(signals, np.float32, computer.pad_left % 4)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_sig = torch.from_numpy(packed.data).to(dev)
    layout = computer.plan_batch(packed.offsets, packed.lengths, dev)
    d_feats = torch.empty((layout.rows, 41), device=dev)
    d_full = torch.empty((layout.rows, 123), device=dev)
    d_row_off = torch.from_numpy(layout.frame_off).to(dev)
    deltas = Deltas(2)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for _ in range(2):
            computer.run_batch(layout, d_sig, out=d_feats)
            lazy = deltas.lazy_device(d_feats, d_row_off)
            cmvn = Standardize()
            cmvn.accumulate_device(lazy)
            cmvn.apply_device(lazy, out=d_full)
        # the whole step was only enqueued: the stream still has (or just had) work, nothing waited
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert not cmvn.check_zero_variance()
    want, stacked = _oracle_chain(computer, signals)
    got = d_full.cpu().numpy()
    assert np.abs(got - np.concatenate(want)).max() <= 2e-3
    stats = cmvn.stats  # pulled back from the GPU on demand
    assert stats.shape == (2, 124) and stats[0, -1] == len(stacked)
    assert np.allclose(stats[0, :-1], stacked.sum(0), rtol=1e-5, atol=1e-2)
    assert np.allclose(stats[1, :-1], (stacked ** 2).sum(0), rtol=1e-5)


def test_run_corpus_single_rank(speech, tmp_path):
    from pydrobert_speech_b200.compute import PackedSignals
    from pydrobert_speech_b200.pipeline import FeaturePipeline
    from pydrobert_speech_b200.post import Deltas, Standardize

    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.README_FBANK)
    signals = _corpus(seed=9)
    packed = PackedSignals.pack(signals, np.float32, computer.pad_left % 4)
    cmvn = Standardize()
    feats, frame_off = FeaturePipeline(computer).run_corpus(packed, cmvn, Deltas(2))
    want, _ = _oracle_chain(computer, signals)
    for u, w in enumerate(want):
        assert np.abs(feats[frame_off[u]:frame_off[u + 1]] - w).max() <= 2e-3
    # the statistics are the reference's format: a reference-style Standardize loads and applies them
    path = str(tmp_path / "cmvn.npy")
    cmvn.save(path)
    again = Standardize(path)
    assert np.allclose(again.stats, cmvn.stats)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, tmp):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in (root, os.path.join(root, "tests", "golden")):
        sys.path.insert(0, path)
    import torch
    import torch.distributed as dist

    import pydrobert_speech_b200 as pds
    from pydrobert_speech_b200.compute import PackedSignals
    from pydrobert_speech_b200.pipeline import FeaturePipeline, shard_utterances
    from pydrobert_speech_b200.post import Deltas, Standardize

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    computer = pds.alias_factory_subclass_from_arg(pds.compute.FrameComputer, cases.README_FBANK)
    signals = _corpus(seed=13, n=31)  # the same corpus on every rank; each takes its shard
    mine = shard_utterances([len(s) for s in signals], world)[rank]
    packed = PackedSignals.pack([signals[i] for i in mine], np.float32, computer.pad_left % 4)
    cmvn = Standardize()
    feats, frame_off = FeaturePipeline(computer).run_corpus(packed, cmvn, Deltas(2))
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), feats=feats, frame_off=frame_off, mine=mine, stats=cmvn.stats)
    dist.destroy_process_group()


def test_run_corpus_two_ranks_nccl(speech, tmp_path):
    """Standardize.allreduce on NCCL: two ranks, each with a shard, agree on the global statistics
    and produce the features the single-process oracle chain gives for the whole corpus"""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.README_FBANK)
    signals = _corpus(seed=13, n=31)
    want, stacked = _oracle_chain(computer, signals)
    seen = []
    stats = []
    for rank in range(world):
        data = np.load(tmp_path / f"rank{rank}.npz")
        stats.append(data["stats"])
        for j, u in enumerate(data["mine"]):
            got = data["feats"][data["frame_off"][j]:data["frame_off"][j + 1]]
            assert np.abs(got - want[u]).max() <= 2e-3
            seen.append(int(u))
    assert sorted(seen) == list(range(31))
    assert np.array_equal(stats[0], stats[1]) and stats[0][0, -1] == len(stacked)
    assert np.allclose(stats[0][0, :-1], stacked.sum(0), rtol=1e-5, atol=1e-2)
