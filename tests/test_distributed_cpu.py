"""World-size-2 gloo tests (CPU): utterance sharding and the CMVN statistics all-reduce -- the
host-side logic of the N > 1 path.  No kernels run here; statistics are fed in directly."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    import pydrobert_speech_b200 as pds
    from pydrobert_speech_b200.pipeline import shard_utterances

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same corpus on every rank
    lengths = rng.integers(20, 200, 37)
    feats = [rng.standard_normal((n, 5)) * (1 + np.arange(5)) + np.arange(5) for n in lengths]
    mine = shard_utterances(lengths, world)[rank]
    std = pds.post.Standardize()
    stats = np.zeros((2, 6))
    for u in mine:  # what pds_cmvn_accumulate produces, computed on the host for this CPU test
        stats[0, :5] += feats[u].sum(0)
        stats[1, :5] += (feats[u] ** 2).sum(0)
        stats[0, 5] += len(feats[u])
    std._stats = stats
    std.allreduce()
    np.save(os.path.join(tmp, f"stats{rank}.npy"), std.stats)
    np.save(os.path.join(tmp, f"shard{rank}.npy"), mine)
    dist.destroy_process_group()


def test_cmvn_allreduce_and_sharding_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    lengths = rng.integers(20, 200, 37)
    feats = np.concatenate([rng.standard_normal((n, 5)) * (1 + np.arange(5)) + np.arange(5) for n in lengths])
    want = np.zeros((2, 6))
    want[0, :5], want[1, :5], want[0, 5] = feats.sum(0), (feats ** 2).sum(0), len(feats)
    stats = [np.load(tmp_path / f"stats{r}.npy") for r in range(world)]
    assert np.allclose(stats[0], want, rtol=1e-12) and np.array_equal(stats[0], stats[1])
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    assert sorted(np.concatenate(shards)) == list(range(37))  # a partition
    loads = [lengths[s].sum() for s in shards]
    assert abs(loads[0] - loads[1]) <= lengths.max()  # balanced to within one utterance


def test_shard_utterances_properties():
    from pydrobert_speech_b200.pipeline import shard_utterances

    rng = np.random.default_rng(1)
    lengths = (16000 * rng.uniform(2, 20, 1000)).astype(np.int64)
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lengths, world)
        assert sorted(np.concatenate(shards)) == list(range(1000))
        loads = np.array([lengths[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lengths.max()
