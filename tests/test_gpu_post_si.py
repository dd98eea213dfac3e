"""Parity of the post-processing, pre-processing and short-integration kernels (through the C
ABI) against the oracle and the goldens produced by the real reference."""
import warnings

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu


def build(speech, base, cfg):
    return speech.alias_factory_subclass_from_arg(base, cfg)


# ---- Deltas ---------------------------------------------------------------------------------
@pytest.mark.parametrize("order", (1, 2, 3))
@pytest.mark.parametrize("ctx", (1, 2, 3))
def test_deltas_match_reference(speech, golden, order, ctx):
    data = golden("post")
    got = speech.post.Deltas(order, context_window=ctx).apply(data["feats"], axis=0)
    want = data[f"deltas_o{order}_w{ctx}"]
    assert got.shape == want.shape and got.dtype == data["feats"].dtype
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_deltas_short_and_axes(speech, golden):
    data = golden("post")
    got = speech.post.Deltas(2).apply(data["feats"][:3], axis=0)
    assert np.abs(got - data["deltas_short"]).max() <= 1e-5
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6, 40, 5))
    # filter along axis 1, stack on a new leading axis: compare with the oracle slice by slice
    got = speech.post.Deltas(1, target_axis=0, concatenate=False).apply(x, axis=1)
    assert got.shape == (2, 6, 40, 5)
    for i in range(6):
        want = oracle.deltas(x[i], 1)[:, 5:]
        assert np.abs(got[1, i] - want).max() <= 1e-5
    # the reference CLI's default: axis=-1
    feats = rng.standard_normal((30, 8))
    got = speech.post.Deltas(2).apply(feats)
    want = oracle.deltas(feats.T, 2)  # filter along coefficients
    assert got.shape == (30, 24)
    assert np.abs(got[:, 8:16] - want[:, 30:60].T).max() <= 1e-5


def test_deltas_wide_and_long(speech):
    """column chunking (cols > 256) and many row tiles, against the oracle"""
    rng = np.random.default_rng(7)
    feats = rng.standard_normal((700, 41))
    got = speech.post.Deltas(2).apply(feats)  # CLI default: filter along the 41 coefficients, 700 columns
    want = oracle.deltas(feats.T, 2)
    assert got.shape == (700, 123)
    assert np.abs(got[:, 41:82] - want[:, 700:1400].T).max() <= 1e-5
    assert np.abs(got[:, 82:] - want[:, 1400:].T).max() <= 1e-5
    tall = rng.standard_normal((5000, 300)).astype(np.float32)
    got = speech.post.Deltas(3, context_window=3).apply(tall, axis=0)
    assert np.abs(got - oracle.deltas(tall, 3, 3)).max() <= 1e-4


def test_deltas_respect_utterance_boundaries(speech):
    import torch

    rng = np.random.default_rng(1)
    lens = [5, 0, 1, 70, 33]
    feats = rng.standard_normal((sum(lens), 41)).astype(np.float32)
    row_off = torch.tensor(np.cumsum([0] + lens), dtype=torch.int64, device="cuda")
    got = speech.post.Deltas(2).apply_device(torch.from_numpy(feats).cuda(), row_off).cpu().numpy()
    begin = 0
    for n in lens:
        if n:
            want = oracle.deltas(feats[begin : begin + n], 2)
            assert np.abs(got[begin : begin + n] - want).max() <= 1e-5
        begin += n


# ---- Standardize ------------------------------------------------------------------------------
def test_cmvn_matches_reference(speech, golden):
    data = golden("post")
    std = speech.post.Standardize()
    begin = 0
    for n in data["cmvn_chunk_lens"]:
        std.accumulate(data["cmvn_chunks"][begin : begin + n])
        begin += n
    assert np.allclose(std.stats, data["cmvn_stats"], rtol=1e-12)  # float64 accumulation, exact inputs
    first = data["cmvn_chunks"][: data["cmvn_chunk_lens"][0]]
    got = std.apply(first)
    assert got.dtype == np.float64
    assert np.abs(got - data["cmvn_applied"]).max() <= 1e-5
    feats = data["feats"]
    assert np.abs(speech.post.Standardize().apply(feats) - data["cmvn_local"]).max() <= 1e-5
    assert np.abs(speech.post.Standardize(norm_var=False).apply(feats) - data["cmvn_applied_novar"]).max() <= 1e-5


def test_cmvn_errors_and_io(speech, tmp_path):
    std = speech.post.Standardize()
    with pytest.raises(ValueError, match="empty"):
        std.accumulate(np.zeros((0, 3)))
    with pytest.raises(ValueError, match="global statistics"):
        std.apply(np.ones(4))
    rng = np.random.default_rng(2)
    # (the reference's raw-binary sanity check rejects negative sums, post.py:130-131)
    std.accumulate(rng.standard_normal((100, 4)) + 5.0, axis=1)
    with pytest.raises(ValueError, match="Expected feature vector of length 4; got 5"):
        std.accumulate(np.ones((3, 5)))
    for name in ("stats.npy", "stats.npz", "stats.bin"):
        path = str(tmp_path / name)
        std.save(path)
        # raw binaries carry no type information: the reference needs force_as="file" as well
        loaded = speech.post.Standardize(path, **({"force_as": "file"} if name.endswith(".bin") else {}))
        assert np.allclose(loaded.stats, std.stats)
    const = np.ones((10, 3))
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        out = speech.post.Standardize().apply(const)
    assert any("0 variance" in str(w.message) for w in caught)
    assert np.allclose(out, 0)


def test_cmvn_large_matches_float64(speech):
    import torch

    rng = np.random.default_rng(3)
    feats = (rng.standard_normal((200003, 123)) * 7 + 3).astype(np.float32)
    std = speech.post.Standardize()
    std.accumulate_device(torch.from_numpy(feats).cuda())
    want = oracle.cmvn_accumulate(feats)
    assert np.allclose(std.stats, want, rtol=1e-12)
    got = std.apply_device(torch.from_numpy(feats).cuda()).cpu().numpy()
    assert np.abs(got - oracle.cmvn_apply(feats, want)).max() <= 2e-5
    assert abs(got.mean()) < 1e-4 and abs(got.std() - 1) < 1e-3


# ---- pre-processors ---------------------------------------------------------------------------
def test_preemphasis_kernel(speech, golden):
    data = golden("stft")
    got = speech.pre.Preemphasize(0.97).apply(data["preemph/signal"].astype(np.float64))
    assert got.dtype == np.float64
    want = data["preemph/signal_out"]
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    two_d = np.stack([data["preemph/signal"][:100], data["preemph/signal"][100:200]])
    with pytest.warns(DeprecationWarning):
        got = speech.pre.Preemphasize(0.5).apply(two_d, axis=1)
    assert np.allclose(got[1], oracle.preemphasize(two_d[1], 0.5), rtol=1e-5, atol=1e-3)


def test_dither_kernel_statistics(speech):
    # reference tests/test_pre.py:32-38
    np.random.seed(1)
    signal = np.zeros(200000)
    out = speech.pre.Dither(2.0).apply(signal)
    assert abs(out.std() - 2.0) < 1e-2 and abs(out.mean()) < 2e-2
    assert np.all(signal == 0)  # not in place
    np.random.seed(1)
    assert np.array_equal(out, speech.pre.Dither(2.0).apply(signal))


def test_fused_deltas_cmvn_matches_separate_passes(speech):
    """Deltas -> Standardize through the fused kernels (pds_deltas_cmvn_*): same statistics and
    output as materialising the deltas first, across utterance boundaries; other filter lengths
    fall back to the separate kernels"""
    import torch

    rng = np.random.default_rng(12)
    dev = torch.device("cuda", 0)
    lengths = [1, 3, 7, 8, 9, 130, 257, 1000, 2, 513]
    feats = torch.from_numpy(rng.standard_normal((sum(lengths), 41)).astype(np.float32) * 3 + 1).to(dev)
    row_off = torch.tensor(np.concatenate([[0], np.cumsum(lengths)]), dtype=torch.int64, device=dev)
    for deltas in (speech.post.Deltas(2), speech.post.Deltas(1), speech.post.Deltas(2, context_window=3)):
        full = deltas.apply_device(feats, row_off)
        sep, fused = speech.post.Standardize(), speech.post.Standardize()
        sep.accumulate_device(full)
        fused.accumulate_device(deltas.lazy_device(feats, row_off))
        assert np.allclose(sep._stats, fused._stats, rtol=1e-12, atol=1e-9)
        want = sep.apply_device(full)
        got = fused.apply_device(deltas.lazy_device(feats, row_off))
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    # float64 reference of the whole chain on one utterance
    one = feats[:777]
    lazy = speech.post.Deltas(2).lazy_device(one)
    cmvn = speech.post.Standardize()
    cmvn.accumulate_device(lazy)
    got = cmvn.apply_device(lazy).cpu().numpy().astype(np.float64)
    ref = oracle.deltas(one.cpu().numpy().astype(np.float64), 2, 2)
    ref = (ref - ref.mean(0)) / ref.std(0)
    assert np.abs(got - ref).max() <= 2e-5


# ---- short integration ------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.SI_CASES))
def test_si_matches_reference(speech, golden, name):
    cfg, _ = cases.SI_CASES[name]
    data = golden("si")
    computer = build(speech, speech.compute.FrameComputer, cfg)
    signal = data[name + "/signal"]
    want = data[name + "/feats"]
    got = computer.compute_full(signal)
    assert got.shape == want.shape and got.dtype == signal.dtype
    if cfg.get("use_log", True):
        assert np.abs(got - want).max() <= 1e-3
        lin = build(speech, speech.compute.FrameComputer, dict(cfg, use_log=False))
        lin_got, lin_want = lin.compute_full(signal).astype(np.float64), data[name + "/feats_linear"]
    else:
        lin_got, lin_want = got.astype(np.float64), want
    scale = np.maximum(np.abs(lin_want), 1e-6 * np.abs(lin_want).max(axis=1, keepdims=True))
    assert (np.abs(lin_got - lin_want) / scale).max() <= 1e-4


def test_si_chunked_and_batch(speech):
    rng = np.random.default_rng(4)
    computer = build(speech, speech.compute.FrameComputer, cases.SI_CASES["si_gabor41"][0])
    for n in (0, 1, 256, 1024, 3000):
        sig = rng.standard_normal(n)
        full = computer.compute_full(sig)
        chunked = speech.compute.frame_by_frame_calculation(computer, sig, chunk_size=333)
        assert chunked.shape == full.shape == (computer.num_frames(n), 41)
        assert np.allclose(full, chunked, atol=1e-5)
    sigs = [rng.standard_normal(n).astype(np.float32) for n in (2000, 0, 50, 4321)]
    for sig, feats in zip(sigs, computer.compute_batch(sigs)):
        assert np.array_equal(feats, computer.compute_full(sig))
    with pytest.raises(ValueError, match="float type"):
        computer.compute_full(np.zeros(100, dtype=np.int16))


# ---- pipeline ---------------------------------------------------------------------------------
def test_pipeline_matches_step_by_step(speech):
    from pydrobert_speech_b200.pipeline import FeaturePipeline

    rng = np.random.default_rng(5)
    computer = build(speech, speech.compute.FrameComputer, cases.README_FBANK)
    sigs = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in (16000, 300, 0, 52000, 8000)]
    pipe = FeaturePipeline(computer, [speech.pre.Preemphasize(0.97)], [speech.post.Deltas(2)], chunk_samples=20000)
    out = pipe(sigs)
    for sig, feats in zip(sigs, out):
        base = oracle.stft_features(
            oracle.preemphasize(sig, 0.97), computer._window, 512, computer._filt_start_idxs,
            computer._truncated_filts, 160, 199, True, True, True, True)
        want = oracle.deltas(base, 2) if len(base) else np.zeros((0, 123))
        assert feats.shape == want.shape
        if len(want):
            assert np.abs(feats - want).max() <= 1e-3


def test_config4_sixty_second_utterance(speech, golden):
    """BASELINE config 4 on one 60 s utterance (6000 x 41) against the reference's own output
    (tests/golden/make_golden.py extra); the signal is regenerated from its seed"""
    want = golden("extra")["c4_60s/feats"]
    signal = (np.random.default_rng(cases.C4_SEED).standard_normal(cases.C4_SAMPLES) * 1000.0).astype(np.float32)
    si = build(speech, speech.compute.FrameComputer, cases.SI_GABOR_41)
    got = si.compute_full(signal)
    assert got.shape == want.shape == (6000, 41)
    assert np.abs(got - want).max() <= 1e-3


# ---- short integration with supports longer than one 1024-point block ---------------------------
def _seeded(spec):
    _, seed, length = spec
    return (np.random.default_rng(seed).standard_normal(length) * 1000.0).astype(np.float32)


@pytest.mark.parametrize("name", sorted(cases.SI_LONG_CASES))
def test_si_long_supports_match_reference(speech, golden, name):
    """the 1024 * R-point overlap-save kernel (R = 2 ... 16) against the reference's own output
    (tests/golden/make_golden.py si_long): 838 ... 6 987 taps, real and complex filters, centered and
    causal pooling"""
    cfg, spec = cases.SI_LONG_CASES[name]
    data = golden("si_long")
    signal = _seeded(spec)
    computer = build(speech, speech.compute.FrameComputer, cfg)
    assert computer._max_support == int(data[name + "/geometry"][1])
    got = computer.compute_full(signal)
    want = data[name + "/feats"]
    assert got.shape == want.shape
    lin = build(speech, speech.compute.FrameComputer, dict(cfg, use_log=False))
    lin_got, lin_want = lin.compute_full(signal).astype(np.float64), data[name + "/feats_linear"]
    if cfg.get("use_log", True):
        assert np.abs(got - want).max() <= 1e-3
    scale = np.maximum(np.abs(lin_want), 1e-6 * np.abs(lin_want).max(axis=1, keepdims=True))
    worst = (np.abs(lin_got - lin_want) / scale).max()
    print(f"WORST {name}: linear relative error {worst:.3g}")
    assert worst <= 1e-4
    # batches, ragged lengths and the chunked interface go through the same tiles
    pieces = [signal[:n] for n in (0, 1, 700, 5000, len(signal))]
    for piece, feats in zip(pieces, computer.compute_batch(pieces)):
        assert np.array_equal(feats, computer.compute_full(piece))
    chunked = speech.compute.frame_by_frame_calculation(computer, signal[:6000], chunk_size=1111)
    assert np.allclose(chunked, computer.compute_full(signal[:6000]), atol=2e-5 if cfg.get("use_log", True) else 0, rtol=2e-5)


@pytest.mark.parametrize("name", sorted(cases.SI_CASES))
def test_si_long_kernel_on_short_supports(speech, golden, monkeypatch, name):
    """PDS_SI_KERNEL=big forces the long-support kernel onto the standard cases (R = 2)"""
    monkeypatch.setenv("PDS_SI_KERNEL", "big")
    cfg, _ = cases.SI_CASES[name]
    data = golden("si")
    computer = build(speech, speech.compute.FrameComputer, cfg)
    signal = data[name + "/signal"]
    want = data[name + "/feats"]
    got = computer.compute_full(signal)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= (1e-3 if cfg.get("use_log", True) else 1e-4 * np.abs(want).max())


@pytest.mark.parametrize("bank", ["fbank", {"name": "fbank", "num_filts": 7}, {"name": "fbank", "num_filts": 8},
                                  {"name": "fbank", "num_filts": 3}, {"name": "fbank", "num_filts": 2}])
def test_si_real_banks_share_transforms(speech, monkeypatch, bank):
    """real banks: filters 2 p and 2 p + 1 ride one complex transform (y1 - i y2) on both overlap-save
    kernels; against one transform per filter (PDS_SI_PAIRS=0) and against the oracle, odd and even
    numbers of filters"""
    rng = np.random.default_rng(31)
    signal = (rng.standard_normal(9000) * 1000).astype(np.float32)
    for cfg in ({"name": "si", "bank": bank}, {"name": "si", "bank": bank, "use_power": True, "use_log": False},
                {"name": "si", "bank": bank, "frame_shift_ms": 20}):
        monkeypatch.setenv("PDS_SI_PAIRS", "0")
        single = build(speech, speech.compute.FrameComputer, cfg).compute_full(signal)
        monkeypatch.delenv("PDS_SI_PAIRS")
        computer = build(speech, speech.compute.FrameComputer, cfg)
        paired = computer.compute_full(signal)
        want = oracle.si_features(
            signal.astype(np.float64), computer._impulse_responses, computer._window.reshape(-1), computer.frame_shift,
            computer._zero_pad, computer._pool_start, computer._frames_lost, cfg.get("use_power", False),
            cfg.get("use_log", True))
        assert paired.shape == single.shape == want.shape
        if cfg.get("use_log", True):
            assert np.abs(paired - want).max() <= 1e-3 and np.abs(paired - single).max() <= 2e-4
        else:
            scale = np.maximum(np.abs(want), 1e-6 * np.abs(want).max(axis=1, keepdims=True))
            assert (np.abs(paired - want) / scale).max() <= 1e-4
