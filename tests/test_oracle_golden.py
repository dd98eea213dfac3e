"""Pin the oracle: every function of oracle/ against goldens from the real reference and against
the reference's own Kaldi known-answer fixtures.  CPU only."""
import numpy as np
import pytest

import cases
import oracle
from conftest import load_ragged


@pytest.mark.parametrize("name", sorted(cases.STFT_CASES))
def test_stft_oracle_matches_reference(golden, name):
    data = golden("stft")
    L, S, centered, kaldi, real = data[name + "/geometry"]
    cfg, _ = cases.STFT_CASES[name]
    args = dict(
        window=data[name + "/window"],
        dft_size=int(data[name + "/dft_size"]),
        start_idxs=data[name + "/starts"],
        truncated_filts=load_ragged(data, name + "/filts"),
        frame_shift=int(S),
        pad_left=oracle.stft.pad_left_for(int(L), int(S), bool(centered), bool(kaldi)),
        use_power=cfg.get("use_power", False),
        use_log=cfg.get("use_log", True),
        include_energy=cfg.get("include_energy", False),
        is_real=bool(real),
    )
    signal = data[name + "/signal"].astype(np.float64)
    got = oracle.stft_features(signal, **args)
    want = data[name + "/feats"]
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    if cfg.get("use_log", True):
        lin = oracle.stft_features(signal, linear=True, **args)
        assert np.allclose(lin, data[name + "/feats_linear"], rtol=1e-12, atol=0)
    looped = oracle.stft_features_looped(signal[:1500], **args)
    assert np.allclose(looped, oracle.stft_features(signal[:1500], **args), rtol=1e-12, atol=1e-12)


def test_stft_oracle_edge_lengths(golden):
    data = golden("stft")
    name = "readme_fbank_noise"
    args = dict(
        window=data[name + "/window"], dft_size=512, start_idxs=data[name + "/starts"],
        truncated_filts=load_ragged(data, name + "/filts"), frame_shift=160, pad_left=199,
        use_power=True, use_log=True, include_energy=True, is_real=True,
    )
    for n in cases.EDGE_LENGTHS:
        got = oracle.stft_features(data["edge/signal"][:n].astype(np.float64), **args)
        want = data[f"edge/feats_{n}"]
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_kaldi_known_answers(golden, speech):
    """tests/test_compute.py:190-208 and tests/test_filters.py:211-223 of the reference"""
    kaldi = golden("kaldi")
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.KALDI_FBANK)
    feats = oracle.stft_features(
        kaldi["noise"].astype(np.float64), computer._window, computer._dft_size,
        computer._filt_start_idxs, computer._truncated_filts, computer.frame_shift,
        computer.pad_left, True, True, False, True,
    )
    feats += 2 * np.log(0.5 * (computer.frame_length - 1))  # undo the unit-normalised Hann window
    feats -= np.log(2)  # undo the Hermitian doubling
    assert feats.shape == kaldi["kaldi_feats"].shape
    assert np.allclose(feats, kaldi["kaldi_feats"])
    bank = speech.alias_factory_subclass_from_arg(speech.filters.LinearFilterBank, cases.KALDI_FBANK["bank"])
    for i, (offset, weights) in enumerate(zip(kaldi["kaldi_filt_offsets"], load_ragged(kaldi, "kaldi_filt"))):
        start, trunc = bank.get_truncated_response(i, 512)
        assert start == offset
        assert np.allclose(trunc[: len(weights)] ** 2, weights, atol=1e-5)
        assert np.allclose(trunc[len(weights) :] ** 2, 0.0)


def test_preemphasis_oracle(golden):
    data = golden("stft")
    got = oracle.preemphasize(data["preemph/signal"], 0.97)
    assert np.allclose(got, data["preemph/signal_out"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", sorted(cases.SI_CASES))
def test_si_oracle_matches_reference(golden, name):
    data = golden("si")
    cfg, _ = cases.SI_CASES[name]
    S, max_support, translation, _, _, centered = (int(v) for v in data[name + "/geometry"])
    if centered:
        pad_left, frame_start, lost = max(0, S - translation), max(0, translation - S), 0
    else:
        pad_left, frame_start, lost = 0, translation, int(max_support - translation <= S)
    got = oracle.si_features(
        data[name + "/signal"], data[name + "/impulse"], data[name + "/window"], S, pad_left,
        frame_start, lost, cfg.get("use_power", False), cfg.get("use_log", True),
    )
    want = data[name + "/feats"]
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)


def test_post_oracle_matches_reference(golden):
    data = golden("post")
    feats = data["feats"]
    for order in (1, 2, 3):
        for ctx in (1, 2, 3):
            got = oracle.deltas(feats, order, ctx)
            assert np.allclose(got, data[f"deltas_o{order}_w{ctx}"], rtol=1e-12, atol=1e-13)
    assert np.allclose(oracle.deltas(feats[:3], 2), data["deltas_short"], rtol=1e-12, atol=1e-13)
    stats, begin = None, 0
    for n in data["cmvn_chunk_lens"]:
        stats = oracle.cmvn_accumulate(data["cmvn_chunks"][begin : begin + n], stats)
        begin += n
    assert np.allclose(stats, data["cmvn_stats"], rtol=1e-13)
    first = data["cmvn_chunks"][: data["cmvn_chunk_lens"][0]]
    assert np.allclose(oracle.cmvn_apply(first, stats), data["cmvn_applied"], rtol=1e-12, atol=1e-13)
    local = oracle.cmvn_accumulate(feats)
    assert np.allclose(oracle.cmvn_apply(feats, local), data["cmvn_local"], rtol=1e-10, atol=1e-12)
    assert np.allclose(oracle.cmvn_apply(feats, local, norm_var=False), data["cmvn_applied_novar"], atol=1e-12)


def test_stft_oracle_config1_full_wav(golden, speech):
    """BASELINE config 1 at full size: the oracle on all of extras/test.wav (149 940 samples) against
    the reference's 937 x 41 output (stored as float32), tables built by the package"""
    data = golden("extra")
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cases.README_FBANK)
    got = oracle.stft_features(
        data["c1/signal"].astype(np.float64), computer._window, computer._dft_size, computer._filt_start_idxs,
        computer._truncated_filts, computer.frame_shift, computer.pad_left, True, True, True, True)
    assert got.shape == (937, 41)
    assert np.abs(got - data["c1/feats"]).max() <= 2e-6  # float32 storage of the golden
    lin = oracle.stft_features(
        data["c1/signal"].astype(np.float64), computer._window, computer._dft_size, computer._filt_start_idxs,
        computer._truncated_filts, computer.frame_shift, computer.pad_left, True, True, True, True, linear=True)
    assert np.allclose(lin, data["c1/feats_linear"], rtol=1e-11, atol=0)


@pytest.mark.parametrize("name", ["si_fbank8_energy", "si_gammatone100_causal", "si_fbank16_power_nolog"])
def test_si_oracle_long_supports(speech, golden, name):
    """impulse responses longer than 1024 samples: the oracle on the HOST tables of the package (built
    without a GPU) against the reference's output (make_golden.py si_long)"""
    cfg, (_, seed, length) = cases.SI_LONG_CASES[name]
    data = golden("si_long")
    signal = (np.random.default_rng(seed).standard_normal(length) * 1000.0).astype(np.float32)
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)
    S, max_support, translation, _, _, centered = (int(v) for v in data[name + "/geometry"])
    assert (computer.frame_shift, computer._max_support, computer._translation) == (S, max_support, translation)
    got = oracle.si_features(
        signal, computer._impulse_responses, computer._window.reshape(-1), S, computer._zero_pad,
        computer._pool_start, computer._frames_lost, cfg.get("use_power", False), cfg.get("use_log", True),
    )
    want = data[name + "/feats"]
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)  # the golden is stored as float32


@pytest.mark.parametrize("name", ["fbank10_power_log_energy", "fbank8_magnitude_causal"])
def test_torch_restatement_gradients_match_the_reference(speech, golden, name):
    """``torch._stft_math`` (the backward pass of the torch mirror) in float64 on the CPU: outputs and
    gradients of the reference's own torch module (tests/golden/make_golden.py torch_grad)"""
    import torch

    import pydrobert_speech_b200.torch as pt

    data = golden("torch_grad")
    edges = data[name + "/filters/offsets"]
    filters = [torch.tensor(data[name + "/filters/values"][a:b], requires_grad=True) for a, b in zip(edges[:-1], edges[1:])]
    frame_length, frame_shift, dft_size, centered = (int(v) for v in data[name + "/geometry"])
    full = name.endswith("energy")
    signal = torch.tensor(data[name + "/signal"], requires_grad=True)
    window = torch.tensor(data[name + "/window"], requires_grad=True)
    feats = pt._stft_math(signal, filters, [int(o) for o in data[name + "/offsets"]], frame_length, frame_shift,
                          bool(centered), window, dft_size, full, full, full, False, True, 1e-5)
    assert np.allclose(feats.detach().numpy(), data[name + "/feats"], rtol=1e-10, atol=1e-10)
    grads = torch.autograd.grad(feats, [signal, window] + filters, torch.tensor(data[name + "/grad_out"]))
    assert np.allclose(grads[0].numpy(), data[name + "/grad_signal"], rtol=1e-8, atol=1e-12)
    assert np.allclose(grads[1].numpy(), data[name + "/grad_window"], rtol=1e-8, atol=1e-9)
    for g, (a, b) in zip(grads[2:], zip(edges[:-1], edges[1:])):
        assert np.allclose(g.numpy(), data[name + "/grad_filters/values"][a:b], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("name", sorted(cases.STFT_CASES))
def test_torch_restatement_matches_every_stft_golden(speech, golden, name):
    """the differentiable restatement behind the torch mirror's backward pass computes what the kernels
    compute: every STFT golden of the reference's NumPy path (complex banks included), float64 on the CPU"""
    import torch

    import pydrobert_speech_b200.torch as pt

    cfg, _ = cases.STFT_CASES[name]
    data = golden("stft")
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)
    filters = [torch.as_tensor(np.asarray(f)) for f in computer._truncated_filts]
    feats = pt._stft_math(
        torch.tensor(data[name + "/signal"].astype(np.float64)), filters, computer._filt_start_idxs, computer.frame_length,
        computer.frame_shift, computer.frame_style == "centered", torch.as_tensor(computer._window), computer._dft_size,
        computer._log, computer._power, computer._include_energy, computer._kaldi_shift, computer._real, 1e-5)
    want = data[name + "/feats"]
    assert tuple(feats.shape) == want.shape
    assert np.allclose(feats.numpy(), want, rtol=1e-9, atol=1e-9)
