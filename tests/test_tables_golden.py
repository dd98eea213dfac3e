"""Host-side tables (scales, banks, windows, folded weights, geometry) against goldens dumped
from the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import cases
from conftest import load_ragged

RTOL = 1e-10  # float64 tables; array vs scalar libm differ in the last ulps only


def build(speech, base, cfg):
    return speech.alias_factory_subclass_from_arg(base, cfg)


@pytest.mark.parametrize("name", sorted(cases.BANK_CASES))
def test_bank_tables(speech, golden, name):
    cfg, width = cases.BANK_CASES[name]
    data = golden("banks")
    bank = build(speech, speech.filters.LinearFilterBank, cfg)
    flags = data[name + "/flags"]
    assert [bank.is_real, bank.is_analytic, bank.is_zero_phase, bank.num_filts] == list(flags)
    assert np.array_equal(np.array(bank.supports), data[name + "/supports"])
    assert np.allclose(np.array(bank.supports_hz), data[name + "/supports_hz"], rtol=RTOL)
    truncs = load_ragged(data, name + "/trunc")
    for i in range(bank.num_filts):
        start, trunc = bank.get_truncated_response(i, width)
        assert start == data[name + "/starts"][i]
        assert trunc.shape == truncs[i].shape and trunc.dtype == truncs[i].dtype
        assert np.allclose(trunc, truncs[i], rtol=RTOL, atol=1e-300)
        for half, key in ((False, "/freq"), (True, "/half")):
            got = bank.get_frequency_response(i, width, half=half)
            assert got.shape == data[name + key][i].shape
            assert np.allclose(got, data[name + key][i], rtol=RTOL, atol=1e-300)
        got = bank.get_impulse_response(i, width)
        want = data[name + "/impulse"][i]
        assert got.dtype == want.dtype
        assert np.allclose(got, want, rtol=1e-9, atol=1e-14 * np.abs(want).max())


@pytest.mark.parametrize("idx", range(len(cases.WINDOW_CASES)))
def test_windows(speech, golden, idx):
    cfg, width = cases.WINDOW_CASES[idx]
    window = build(speech, speech.filters.WindowFunction, cfg).get_impulse_response(width)
    assert np.allclose(window, golden("banks")[f"window{idx}"], rtol=1e-13, atol=0)


def test_scales_round_trip(speech):
    # reference tests/test_scales.py:21-26
    for cfg in ("mel", "bark", {"name": "linear", "low_hz": 20.0, "slope_hz": 2.0},
                {"name": "octave", "low_hz": 20.0}):
        scale = build(speech, speech.scales.ScalingFunction, cfg)
        for hertz in np.linspace(20, 8000, 200):
            assert np.isclose(scale.scale_to_hertz(scale.hertz_to_scale(hertz)), hertz)


def test_alias_resolution(speech):
    f, s, c = speech.filters, speech.scales, speech.compute
    assert isinstance(build(speech, f.WindowFunction, "tri"), f.BartlettWindow)
    assert isinstance(build(speech, f.LinearFilterBank, {"name": "tri", "scaling_function": "mel"}),
                      f.TriangularOverlappingFilterBank)
    assert isinstance(build(speech, f.LinearFilterBank, "fbank"), f.Fbank)
    assert isinstance(build(speech, f.LinearFilterBank, {"alias": "tonebank", "scaling_function": "mel"}),
                      f.ComplexGammatoneFilterBank)
    assert isinstance(build(speech, s.ScalingFunction, "mel"), s.MelScaling)
    assert isinstance(build(speech, speech.post.PostProcessor, "cmvn"), speech.post.Standardize)
    assert isinstance(build(speech, speech.pre.PreProcessor, "preemph"), speech.pre.Preemphasize)
    assert isinstance(build(speech, c.FrameComputer, cases.README_FBANK), c.STFTFrameComputer)
    bank = f.Fbank()
    assert build(speech, f.LinearFilterBank, bank) is bank
    with pytest.raises(ValueError, match="Cannot find subclass with alias 'nope'"):
        build(speech, f.LinearFilterBank, "nope")
    with pytest.raises(ValueError, match="Invalid frame style"):
        c.STFTFrameComputer("fbank", frame_style="sideways")
    with pytest.raises(ValueError, match="Invalid frequency range"):
        f.Fbank(low_hz=-1)


@pytest.mark.parametrize("name", sorted(cases.STFT_CASES))
def test_stft_tables_and_dense_restatement(speech, golden, name):
    """Product tables == reference tables, and log(|rfft|^p @ W^T) with the *replayed* fold
    reproduces the reference's features to float64 round-off."""
    cfg, _ = cases.STFT_CASES[name]
    data = golden("stft")
    computer = build(speech, speech.compute.FrameComputer, cfg)
    L, S, centered, kaldi, real = data[name + "/geometry"]
    assert (computer.frame_length, computer.frame_shift) == (L, S)
    assert (computer.frame_style == "centered", computer.kaldi_shift, bool(computer._real)) == (
        bool(centered), bool(kaldi), bool(real))
    assert computer._dft_size == data[name + "/dft_size"]
    assert np.allclose(computer._window, data[name + "/window"], rtol=1e-13, atol=0)
    assert np.array_equal(computer._filt_start_idxs, data[name + "/starts"])
    for mine, ref in zip(computer._truncated_filts, load_ragged(data, name + "/filts")):
        assert np.allclose(mine, ref, rtol=RTOL, atol=1e-300)
    # dense float64 restatement with the folded weights
    import oracle

    signal = data[name + "/signal"].astype(np.float64)
    frames = oracle.frame_signal(signal, L, S, computer.pad_left)
    spect = np.abs(np.fft.rfft(frames * computer._window, n=computer._dft_size))
    vals = (spect ** 2 if computer._power else spect) @ computer.folded_weights.dense.T
    want = data[name + ("/feats_linear" if computer._log else "/feats")]
    want = want[:, 1:] if computer.includes_energy else want
    assert np.allclose(vals, want, rtol=1e-11, atol=1e-13 * np.abs(want).max())
    w = computer.folded_weights
    assert w.taps.dtype == np.float32 and w.nnz == len(w.taps)
    for f in range(len(w.lo)):  # the band hull loses nothing
        row = np.zeros(w.dense.shape[1])
        row[w.lo[f] : w.lo[f] + w.length[f]] = w.dense[f, w.lo[f] : w.lo[f] + w.length[f]]
        assert np.array_equal(row, w.dense[f])


def test_replayed_fold_is_not_the_true_fold(speech):
    """Regression guard for SURVEY.md H1: nobody "fixes" the reference's conjugate branch."""
    computer = build(speech, speech.compute.FrameComputer, cases.GAMMATONE_64)
    N = computer._dft_size
    true = np.zeros_like(computer.folded_weights.dense)
    for f, (start, filt) in enumerate(zip(computer._filt_start_idxs, computer._truncated_filts)):
        full = np.zeros(N)
        np.add.at(full, (start + np.arange(len(filt))) % N, np.abs(filt) ** 2)
        true[f] = full[: N // 2 + 1]
        true[f, 1 : N // 2] += full[N - 1 : N // 2 : -1]
    assert not np.allclose(true, computer.folded_weights.dense, rtol=1e-3)


def test_readme_geometry(speech):
    computer = build(speech, speech.compute.FrameComputer, cases.README_FBANK)
    assert (computer.frame_length, computer.frame_shift, computer.dft_size) == (400, 160, 512)
    assert (computer.num_coeffs, computer.pad_left, computer.folded_weights.nnz) == (41, 199, 493)
    assert [computer.num_frames(n) for n in (0, 200, 201, 240, 149940)] == [0, 0, 1, 2, 937]
    assert computer.frame_length_ms == 25 and computer.frame_shift_ms == 10


def test_si_tables(speech, golden):
    data = golden("si")
    for name, (cfg, _) in cases.SI_CASES.items():
        computer = build(speech, speech.compute.FrameComputer, cfg)
        S, max_support, translation, frame_length, dft_size, centered = data[name + "/geometry"]
        assert (computer.frame_shift, computer._max_support, computer._translation) == (S, max_support, translation)
        assert (computer.frame_length, computer._dft_size) == (frame_length, dft_size)
        assert (computer.frame_style == "centered") == bool(centered)
        want = data[name + "/impulse"]
        assert computer._impulse_responses.shape == want.shape
        assert np.allclose(computer._impulse_responses, want, rtol=1e-8, atol=1e-12 * np.abs(want).max())
        assert np.allclose(computer._window.reshape(-1), data[name + "/window"], rtol=1e-13)
        counts = data[name + "/sweep_counts"]
        mine = np.array([computer.num_frames(n) for n in data[name + "/sweep_lens"]])
        ok = counts >= 0  # the reference raises for a few lengths of tiny causal banks
        assert np.array_equal(mine[ok], counts[ok])


def test_packed_layout_alignment(speech):
    offsets, total = speech.compute.PackedSignals.layout([5, 0, 17, 4], lead=3)
    assert list(offsets) == [3, 11, 11, 31] and total == 8 + 0 + 20 + 4 + 3 + 4
    packed = speech.compute.PackedSignals.pack([np.arange(5), np.arange(7)], np.float32, lead=3)
    assert np.array_equal(packed.data[packed.offsets[1] : packed.offsets[1] + 7], np.arange(7))
    assert all((o - 3) % 4 == 0 for o in packed.offsets)
