"""The hot path at BASELINE.json's full size (configs[1]: 10 000 synthetic 16 kHz utterances of
2-20 s, 30.5 audio-hours, 11 M frames), checked through size-independent properties and against
the oracle on a sample of utterances.  Runs in a few seconds on a B200."""
import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

LOG_TOL = 1e-3
N_UTTS = 10000


def oracle_feats(computer, signal):
    return oracle.stft_features(
        signal, computer._window, computer._dft_size, computer._filt_start_idxs, computer._truncated_filts,
        computer.frame_shift, computer.pad_left, computer._power, computer._log, computer.includes_energy,
        computer._real)


@pytest.fixture(scope="module")
def corpus(speech):
    import torch

    from pydrobert_speech_b200.compute import PackedSignals

    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.README_FBANK)
    rng = np.random.default_rng(0)
    lengths = (16000 * rng.uniform(2, 20, N_UTTS)).astype(np.int64)
    offsets, total = PackedSignals.layout(lengths, computer.pad_left % 4)
    device = torch.device("cuda", 0)
    gen = torch.Generator(device=device).manual_seed(99)
    d_signal = torch.randn(total, device=device, generator=gen) * 1000.0
    layout = computer.plan_batch(offsets, lengths, device)
    feats = computer.run_batch(layout, d_signal)
    torch.cuda.synchronize()
    return computer, lengths, offsets, d_signal, layout, feats


def test_shape_and_finiteness(corpus):
    import torch

    computer, lengths, _, _, layout, feats = corpus
    want_rows = sum(computer.num_frames(int(n)) for n in lengths)
    assert layout.rows == want_rows == 10_989_392
    assert feats.shape == (want_rows, 41)
    assert bool(torch.isfinite(feats).all())
    # white noise of variance 1e6 in a 40-mel bank: every coefficient far above the log floor (-11.5)
    assert float(feats.min()) > -5.0 and float(feats.max()) < 30.0


def test_scaling_by_a_power_of_two_shifts_the_log_features(corpus):
    """power features are homogeneous of degree 2: feats(4x) = feats(x) + ln 16, exactly in float32
    arithmetic up to the log approximation -- checked on all 11 M frames"""
    import torch

    computer, _, _, d_signal, layout, feats = corpus
    scaled = computer.run_batch(layout, d_signal * 4.0)
    shift = (scaled - feats) - float(np.log(16.0))
    assert float(shift.abs().max()) <= 2e-5


def test_sampled_utterances_match_the_oracle_and_single_runs(corpus):
    """utterances inside the 10 000-utterance batch: within tolerance of the float64 oracle and
    bitwise equal to the same utterance computed on its own (different tiling, same arithmetic)"""
    computer, lengths, offsets, d_signal, layout, feats = corpus
    rng = np.random.default_rng(1)
    picks = [0, 1, N_UTTS - 1, int(np.argmin(lengths)), int(np.argmax(lengths))] + list(rng.integers(0, N_UTTS, 6))
    for u in picks:
        sig = d_signal[int(offsets[u]) : int(offsets[u] + lengths[u])].cpu().numpy()
        got = feats[int(layout.frame_off[u]) : int(layout.frame_off[u + 1])].cpu().numpy()
        want = oracle_feats(computer, sig.astype(np.float64))
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= LOG_TOL
        assert np.array_equal(got, computer.compute_full(sig))


def test_post_chain_statistics_at_full_size(speech, corpus):
    """fbank -> Deltas(2) -> corpus CMVN (config 5) on all frames: zero mean / unit variance of the
    normalised output, fused and separate passes agree"""
    import torch

    computer, _, _, _, layout, feats = corpus
    row_off = torch.from_numpy(layout.frame_off).to(feats.device)
    deltas = speech.post.Deltas(2)
    fused = speech.post.Standardize()
    lazy = deltas.lazy_device(feats, row_off)
    fused.accumulate_device(lazy)
    out = fused.apply_device(lazy)
    assert out.shape == (layout.rows, 123)
    mean = out.double().mean(0)
    var = (out.double() ** 2).mean(0) - mean ** 2
    assert float(mean.abs().max()) <= 2e-4
    assert float((var - 1).abs().max()) <= 2e-4
    separate = speech.post.Standardize()
    separate.accumulate_device(deltas.apply_device(feats, row_off))
    assert np.allclose(fused._stats, separate._stats, rtol=1e-11, atol=1e-6)


def test_short_integration_on_long_utterances(speech):
    """config 4 (SI + Gabor-41 with frame pooling) on long utterances (20 x 60 s + 2 x 600 s):
    homogeneity of the pooled magnitude, independence of far-away samples, oracle on a prefix"""
    import torch

    from pydrobert_speech_b200.compute import PackedSignals

    cfg = {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}}
    si = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)
    lengths = np.array([16000 * 60] * 20 + [16000 * 600] * 2, dtype=np.int64)
    offsets, total = PackedSignals.layout(lengths, 0)
    device = torch.device("cuda", 0)
    gen = torch.Generator(device=device).manual_seed(7)
    d_signal = torch.randn(total, device=device, generator=gen) * 1000.0
    feats, frame_off = si.compute_packed_device(d_signal, offsets, lengths)
    assert feats.shape == (int(frame_off[-1]), 41) and bool(torch.isfinite(feats).all())
    assert int(frame_off[-1]) == sum(si.num_frames(int(n)) for n in lengths)
    # log |y| pooled: feats(4x) = feats(x) + ln 4
    scaled, _ = si.compute_packed_device(d_signal * 4.0, offsets, lengths)
    assert float(((scaled - feats) - float(np.log(4.0))).abs().max()) <= 2e-5
    # the first 2 s of a 600 s utterance: same frames as the prefix alone (away from its end) and
    # within tolerance of the float64 oracle
    u = 21
    prefix = d_signal[int(offsets[u]) : int(offsets[u]) + 32000].cpu().numpy()
    inside = feats[int(frame_off[u]) : int(frame_off[u]) + 190].cpu().numpy()
    alone = si.compute_full(prefix)
    assert np.allclose(inside, alone[:190], rtol=0, atol=2e-6)
    want = oracle.si_features(
        prefix.astype(np.float64), si._impulse_responses, si._window.reshape(-1), si.frame_shift,
        si._zero_pad, si._pool_start, si._frames_lost, si._power, si._log)
    assert np.abs(alone - want).max() <= LOG_TOL
