"""Parity of the fused STFT kernel (through the C ABI) against the oracle and the golden
vectors produced by the real reference.

Tolerances (stated by BASELINE.json's north star, checked here):
  * log features:           max |cuda - reference| <= 1e-3
  * linear power/magnitude: |cuda - reference| <= 1e-4 * max(reference value, row scale)
    i.e. 1e-4 relative, where coefficients more than ~60 dB below the strongest coefficient of
    their frame are compared against that frame scale (float32 FFT round-off is relative to
    the frame's spectrum, not to each bin).
"""
import numpy as np
import pytest

import cases
import oracle
from conftest import load_ragged

pytestmark = pytest.mark.gpu

LOG_TOL = 1e-3
LIN_RTOL = 1e-4


def build(speech, cfg):
    return speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)


def oracle_feats(computer, signal, linear=False):
    return oracle.stft_features(
        signal,
        computer._window,
        computer._dft_size,
        computer._filt_start_idxs,
        computer._truncated_filts,
        computer.frame_shift,
        computer.pad_left,
        computer._power,
        computer._log,
        computer.includes_energy,
        computer._real,
        linear=linear,
    )


WORST = {}  # test name -> (worst floored, worst un-floored) linear relative error; printed at the end


def check_linear(got, want, name=None):
    scale = np.maximum(np.abs(want), 1e-6 * np.abs(want).max(axis=1, keepdims=True))
    err = np.abs(got - want) / scale
    if name is not None:  # the relaxation has a number: the error relative to each coefficient itself
        raw = np.abs(got - want) / np.maximum(np.abs(want), np.finfo(np.float64).tiny)
        WORST[name] = (float(err.max()), float(raw.max()))
    assert err.max() <= LIN_RTOL, f"linear relative error {err.max():.3g}"


@pytest.fixture(scope="module", autouse=True)
def report_worst_linear_errors():
    yield
    if WORST:
        lines = [f"  {k:28s} floored {a:.2e}   per coefficient (un-floored) {b:.2e}" for k, (a, b) in sorted(WORST.items())]
        print("\nworst linear relative error vs the reference, per config (tolerance 1e-4 on the floored figure):\n"
              + "\n".join(lines))


@pytest.mark.parametrize("name", sorted(cases.STFT_CASES))
def test_matches_reference_golden(speech, golden, name):
    cfg, _ = cases.STFT_CASES[name]
    data = golden("stft")
    signal = data[name + "/signal"]
    want = data[name + "/feats"]
    computer = build(speech, cfg)
    got = computer.compute_full(signal)
    assert got.dtype == signal.dtype and got.shape == want.shape
    if cfg.get("use_log", True):
        assert np.abs(got - want).max() <= LOG_TOL
        lin = build(speech, dict(cfg, use_log=False))
        check_linear(lin.compute_full(signal).astype(np.float64), data[name + "/feats_linear"], name)
    else:
        check_linear(got.astype(np.float64), want, name)


def test_config1_full_length_wav(speech, golden):
    """BASELINE config 1 at its real size: all 149 940 samples of extras/test.wav (16-bit PCM) ->
    937 x 41, against the reference's float64 output (tests/golden/make_golden.py extra)"""
    data = golden("extra")
    wav = data["c1/signal"]
    assert wav.dtype == np.int16 and len(wav) == 149940
    computer = build(speech, cases.README_FBANK)
    got = computer.compute_full(wav.astype(np.float32))
    assert got.shape == (937, 41)
    assert np.abs(got - data["c1/feats"]).max() <= LOG_TOL
    assert np.array_equal(got, computer.compute_batch([wav])[0])  # the int16 staging path, same bits
    lin = build(speech, dict(cases.README_FBANK, use_log=False))
    check_linear(lin.compute_full(wav.astype(np.float32)).astype(np.float64), data["c1/feats_linear"], "c1_full_wav")


@pytest.mark.parametrize("n", cases.EDGE_LENGTHS)
def test_edge_lengths(speech, golden, n):
    data = golden("stft")
    signal = data["edge/signal"][:n]
    want = data[f"edge/feats_{n}"]
    got = build(speech, cases.README_FBANK).compute_full(signal)
    assert got.shape == want.shape
    if len(want):
        assert np.abs(got - want).max() <= LOG_TOL


def test_batch_equals_single_and_oracle(speech):
    rng = np.random.default_rng(5)
    computer = build(speech, cases.README_FBANK)
    lengths = [0, 1, 200, 201, 399, 5000, 16000, 33333, 48000, 160 * 32 + 240, 160 * 64 + 241]
    signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
    batch = computer.compute_batch(signals)
    assert len(batch) == len(signals)
    for sig, feats in zip(signals, batch):
        want = oracle_feats(computer, sig.astype(np.float64))
        assert feats.shape == want.shape
        if len(want):
            assert np.abs(feats - want).max() <= LOG_TOL
            assert np.array_equal(feats, computer.compute_full(sig))


def test_int16_input_matches_float(speech):
    rng = np.random.default_rng(6)
    computer = build(speech, cases.README_FBANK)
    pcm = rng.integers(-32768, 32767, 20000).astype(np.int16)
    a = computer.compute_batch([pcm])[0]
    b = computer.compute_batch([pcm.astype(np.float32)])[0]
    assert np.array_equal(a, b)


def test_fused_preemphasis_matches_reference(speech, golden):
    data = golden("stft")
    computer = build(speech, cases.README_FBANK)
    got = computer.compute_batch([data["preemph/signal"]], preemph=0.97)[0]
    assert np.abs(got - data["preemph/feats"]).max() <= LOG_TOL


def test_fused_dither_statistics(speech):
    # dither is pinned statistically only (reference tests/test_pre.py:32-38): energy of a
    # silent signal + N(0, c^2) is c^2 on average
    computer = build(speech, dict(cases.README_FBANK, use_log=False))
    feats = computer.compute_batch([np.zeros(160000, np.float32)], dither=3.0, seed=1)[0]
    assert abs(feats[:, 0].mean() - 9.0) < 0.1
    again = computer.compute_batch([np.zeros(160000, np.float32)], dither=3.0, seed=1)[0]
    other = computer.compute_batch([np.zeros(160000, np.float32)], dither=3.0, seed=2)[0]
    assert np.array_equal(feats, again) and not np.array_equal(feats, other)


@pytest.mark.parametrize("frame_ms,shift_ms", [(64, 16), (100, 20), (70, 10), (25, 10.0625), (68.875, 10.0625)])
def test_large_dft_sizes(speech, frame_ms, shift_ms):
    """1024- and 2048-point transforms (long frames / high sampling rates): the tensor-core kernel
    with 32- and 16-frame tiles against the oracle, ragged batch incl. utterance-edge tiles"""
    cfg = dict(cases.README_FBANK, frame_length_ms=frame_ms, frame_shift_ms=shift_ms)
    from pydrobert_speech_b200._gpu import current_device
    from pydrobert_speech_b200._lib import get_lib

    rng = np.random.default_rng(31)
    computer = build(speech, cfg)
    assert computer._dft_size in (512, 1024, 2048)  # 10.0625 ms = 161 samples: odd frame shift
    plan = computer._plan(current_device())
    assert get_lib().pds_stft_is_fast_path(plan.handle) == 1
    lengths = [0, 700, 1601, 5000, 16000, 40001, computer.frame_shift * 33 + 7]
    signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
    for sig, got in zip(signals, computer.compute_batch(signals)):
        want = oracle_feats(computer, sig.astype(np.float64))
        assert got.shape == want.shape
        if len(want):
            assert np.abs(got - want).max() <= LOG_TOL


@pytest.mark.parametrize("order", ["dither", "dither_preemph", "preemph_dither"])
def test_fused_preprocessing_matches_standalone_passes(speech, order):
    """The vectorised staging paths (four samples per Philox call, aligned vector loads, reflected
    edges per element) against the stand-alone pds_dither / pds_preemphasize passes followed by the
    plain kernel: same random stream (seed, utterance, sample), same features"""
    import torch

    from pydrobert_speech_b200.pre import _launch_rows, _rows_on_device

    rng = np.random.default_rng(21)
    computer = build(speech, cases.README_FBANK)
    for n in (201, 403, 4000, 16001, 33333):
        sig = (rng.standard_normal(n) * 100).astype(np.float32)
        device, d_in, offsets, lengths, _ = _rows_on_device(sig, None)
        seed = 1234
        if order == "dither":
            d_pre = _launch_rows("pds_dither", d_in, offsets, lengths, device, 2.0, seed)
            kwargs = dict(dither=2.0)
        elif order == "dither_preemph":
            d_pre = _launch_rows("pds_dither", d_in, offsets, lengths, device, 2.0, seed)
            d_pre = _launch_rows("pds_preemphasize", d_pre, offsets, lengths, device, 0.97)
            kwargs = dict(dither=2.0, preemph=0.97, dither_first=True)
        else:
            d_pre = _launch_rows("pds_preemphasize", d_in, offsets, lengths, device, 0.97)
            d_pre = _launch_rows("pds_dither", d_pre, offsets, lengths, device, 2.0, seed)
            kwargs = dict(dither=2.0, preemph=0.97, dither_first=False)
        want = computer.compute_batch([d_pre.cpu().numpy()])[0]
        got = computer.compute_batch([sig], seed=seed, **kwargs)[0]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-4


def test_chunked_equals_full(speech):
    rng = np.random.default_rng(7)
    for cfg in (cases.README_FBANK, cases.KALDI_FBANK, cases.GAMMATONE_64):
        computer = build(speech, cfg)
        for n in (0, 1, 256, 1024, 5000):
            sig = rng.standard_normal(n)
            full = computer.compute_full(sig)
            chunked = speech.compute.frame_by_frame_calculation(computer, sig, chunk_size=333)
            assert chunked.shape == full.shape
            assert np.allclose(full, chunked, atol=1e-5)
            assert not computer.started


def test_already_started_raises(speech):
    computer = build(speech, cases.README_FBANK)
    computer.compute_chunk(np.zeros(10))
    with pytest.raises(ValueError, match="Already started"):
        computer.compute_full(np.zeros(1000))
    computer.finalize()


def test_host_buffer_entry_point(speech):
    """pds_stft_compute_host: the call a non-PyTorch host binds"""
    import ctypes

    from pydrobert_speech_b200._gpu import current_device
    from pydrobert_speech_b200._lib import get_lib, check

    rng = np.random.default_rng(8)
    computer = build(speech, cases.README_FBANK)
    plan = computer._plan(current_device())
    sigs = [(rng.standard_normal(n) * 100).astype(np.float32) for n in (4000, 100, 9001)]
    packed = speech.compute.PackedSignals.pack(sigs, np.float32, computer.pad_left % 4)
    i64p = ctypes.POINTER(ctypes.c_int64)
    frame_off = np.zeros(4, np.int64)
    cap = sum(computer.num_frames(len(s)) for s in sigs)
    out = np.zeros((cap, computer.num_coeffs), np.float32)
    check(get_lib().pds_stft_compute_host(
        plan.handle, packed.data.ctypes.data, 0, len(packed.data), 3,
        packed.offsets.ctypes.data_as(i64p), packed.lengths.ctypes.data_as(i64p),
        out.ctypes.data, cap, frame_off.ctypes.data_as(i64p), 0))
    for u, sig in enumerate(sigs):
        assert np.array_equal(out[frame_off[u]:frame_off[u + 1]], computer.compute_full(sig))


@pytest.mark.parametrize("cfg_name", ["readme", "kaldi", "gammatone_L512", "magnitude"])
def test_kernel_variants_match_each_other_and_oracle(speech, monkeypatch, cfg_name):
    """Every fused kernel on a ragged batch -- the default (stft_tc2_kernel: tensor-core bank, phases
    re-cut), stft_tc_kernel (PDS_STFT_KERNEL=1), the warp-specialised pipeline with 16-frame tiles
    (=p), the scalar bank (=scalar) and the round-1 software pipeline (=ws): same features, within
    tolerance of the oracle"""
    cfg = {
        "readme": cases.README_FBANK,
        "kaldi": cases.KALDI_FBANK,
        "gammatone_L512": dict(cases.GAMMATONE_64, frame_length_ms=32),
        "magnitude": dict(cases.README_FBANK, use_power=False, use_log=False),
    }[cfg_name]
    rng = np.random.default_rng(11)
    computer = build(speech, cfg)
    lengths = [0, 1, 201, 399, 5000, 16000, 33333, 160 * 32 + 240, 160 * 64 + 241, 160 * 95, 48000] * 3
    signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
    results = {}
    for name, value in (("ws", "ws"), ("scalar", "scalar"), ("tc", "1"), ("pipe", "p"), ("tc2", None)):
        if value is None:
            monkeypatch.delenv("PDS_STFT_KERNEL")
        else:
            monkeypatch.setenv("PDS_STFT_KERNEL", value)
        results[name] = computer.compute_batch(signals)
    assert computer.kernel_name() == "pds::stft_tc2_kernel"
    monkeypatch.setenv("PDS_STFT_BANK", "tf32")  # the split-tf32 bank instead of the two-term bf16 one
    results["tc2_tf32"] = computer.compute_batch(signals)
    monkeypatch.delenv("PDS_STFT_BANK")
    for i, sig in enumerate(signals):
        want = oracle_feats(computer, sig.astype(np.float64))
        a, b = results["ws"][i], results["scalar"][i]
        assert all(r[i].shape == want.shape for r in results.values())
        if len(want):
            assert np.allclose(a, b, rtol=2e-6, atol=2e-6)  # same math, different FMA contraction
            for name in ("tc", "pipe", "tc2", "tc2_tf32"):
                got = results[name][i]
                # tensor-core bank on sums of non-negative terms: two-term bf16 splits 3 * 2^-17 relative
                # (the default), split-tf32 2^-20 (PDS_STFT_BANK=tf32 and the pipelined kernel)
                tol = 5e-6 if name in ("pipe", "tc2_tf32") and not cfg_name.startswith("gammatone") else 5e-5
                if name == "tc2_tf32":
                    tol = 5e-6
                assert np.allclose(got, b, rtol=tol, atol=tol), name
                if computer._log:
                    assert np.abs(got - want).max() <= LOG_TOL
                else:
                    check_linear(got.astype(np.float64), want)
            assert np.array_equal(results["tc"][i], results["tc2"][i])  # same arithmetic, other phase cut


@pytest.mark.parametrize("cfg_name", ["fbank400", "gabor320_mag", "odd_size"])
def test_non_power_of_two_dft_bluestein(speech, monkeypatch, cfg_name):
    """pad_to_nearest_power_of_two = false (compute.py:344-347): dft_size = frame length.  Sizes up
    to 512 run Bluestein's algorithm on the 1024-point in-register FFT; checked against the float64
    oracle and against the O(L K) direct-DFT kernel it replaces (PDS_STFT_NO_BLUESTEIN=1)."""
    cfg = {
        "fbank400": dict(cases.README_FBANK, pad_to_nearest_power_of_two=False),  # N = 400
        "gabor320_mag": {"name": "stft", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 30},
                         "frame_length_ms": 20, "pad_to_nearest_power_of_two": False, "use_power": False,
                         "use_log": False, "include_energy": True},               # N = 320
        "odd_size": {"name": "stft", "bank": {"name": "tri", "scaling_function": "mel", "num_filts": 20}, "frame_length_ms": 20.7,
                     "frame_shift_ms": 7.3, "pad_to_nearest_power_of_two": False, "use_power": True},  # N = 331
    }[cfg_name]
    rng = np.random.default_rng(17)
    lengths = [0, 1, 250, 3000, 16000, 40000, 160 * 8 + 399, 160 * 16 + 1]
    signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
    computer = build(speech, cfg)
    assert computer._dft_size == computer.frame_length and computer.kernel_name() == "pds::stft_bluestein_kernel"
    got = computer.compute_batch(signals)
    monkeypatch.setenv("PDS_STFT_NO_BLUESTEIN", "1")
    direct_computer = build(speech, cfg)
    assert direct_computer.kernel_name() == "pds::stft_direct_kernel"
    direct = direct_computer.compute_batch(signals)
    for sig, a, b in zip(signals, got, direct):
        want = oracle_feats(computer, sig.astype(np.float64))
        assert a.shape == want.shape == b.shape
        if len(want):
            if computer._log:
                assert np.abs(a - want).max() <= LOG_TOL and np.abs(b - want).max() <= LOG_TOL
            else:
                check_linear(a.astype(np.float64), want, "bluestein_" + cfg_name)
    # fused pre-processing and 16-bit input take the same staging code as every other kernel
    pcm = rng.integers(-20000, 20000, 30000).astype(np.int16)
    monkeypatch.delenv("PDS_STFT_NO_BLUESTEIN")
    a = computer.compute_batch([pcm], preemph=0.97)[0]
    want = oracle_feats(computer, oracle.preemphasize(pcm.astype(np.float64), 0.97))
    if computer._log:
        assert np.abs(a - want).max() <= LOG_TOL
    else:
        check_linear(a.astype(np.float64), want)


# ---- stft_umma_kernel: the transform on tcgen05 (opt-in, PDS_STFT_KERNEL=u) ----------------------
UMMA_KERNEL = "pds::stft_umma_kernel"


@pytest.mark.parametrize("name", sorted(cases.STFT_CASES))
def test_tcgen05_kernel_matches_reference_golden(speech, golden, monkeypatch, name):
    """Every reference golden through the tcgen05 transform (four 128-point real DFTs per frame as fp16
    two-term GEMMs, radix-4 + filter bank in the epilogue warps); configurations the kernel does not
    take (dft_size other than 512, operands beyond shared memory) must fall back and still be right"""
    monkeypatch.setenv("PDS_STFT_KERNEL", "u")
    cfg, _ = cases.STFT_CASES[name]
    data = golden("stft")
    signal = data[name + "/signal"]
    want = data[name + "/feats"]
    computer = build(speech, cfg)
    got = computer.compute_full(signal)
    eligible = computer._dft_size == 512 and computer.frame_shift % 4 == 0 and len(computer._truncated_filts) <= 64
    if name in ("readme_fbank_wav", "readme_fbank_noise", "kaldi_fbank", "gammatone64"):
        assert eligible and computer.kernel_name() == UMMA_KERNEL
    print(f"{name}: {computer.kernel_name()}")
    assert got.shape == want.shape
    if cfg.get("use_log", True):
        assert np.abs(got - want).max() <= LOG_TOL
        lin = build(speech, dict(cfg, use_log=False))
        check_linear(lin.compute_full(signal).astype(np.float64), data[name + "/feats_linear"], "tcgen05 " + name)
    else:
        check_linear(got.astype(np.float64), want, "tcgen05 " + name)


@pytest.mark.parametrize("cfg_name", ["readme", "kaldi", "gammatone", "magnitude", "no_energy"])
def test_tcgen05_kernel_ragged_batches_and_inputs(speech, monkeypatch, cfg_name):
    """ragged batch (empty, shorter than a frame, partial tiles, several tiles) against the oracle and the
    default kernel; 16-bit PCM input; fused pre-emphasis and dither (same Philox stream as the default
    kernel, so the two agree to rounding)"""
    cfg = {
        "readme": cases.README_FBANK,
        "kaldi": cases.KALDI_FBANK,
        "gammatone": cases.GAMMATONE_64,
        "magnitude": dict(cases.README_FBANK, use_power=False, use_log=False),
        "no_energy": dict(cases.README_FBANK, include_energy=False),
    }[cfg_name]
    rng = np.random.default_rng(12)
    lengths = [0, 1, 201, 399, 5000, 16000, 33333, 160 * 32 + 240, 160 * 64 + 241, 160 * 95, 48000] * 2
    signals = [(rng.standard_normal(n) * 1000).astype(np.float32) for n in lengths]
    # a quiet utterance next to loud ones, and one with an offset: the per-frame scale at work
    signals.append((rng.standard_normal(20000) * 1e-3).astype(np.float32))
    signals.append((rng.standard_normal(20000) * 1000 + 2000).astype(np.float32))
    computer = build(speech, cfg)
    monkeypatch.delenv("PDS_STFT_KERNEL", raising=False)
    ref = computer.compute_batch(signals)
    monkeypatch.setenv("PDS_STFT_KERNEL", "u")
    got = computer.compute_batch(signals)
    assert computer.kernel_name() == UMMA_KERNEL
    for sig, a, b in zip(signals, got, ref):
        want = oracle_feats(computer, sig.astype(np.float64))
        assert a.shape == b.shape == want.shape
        if not len(want):
            continue
        if computer._log:
            assert np.abs(a - want).max() <= LOG_TOL
        else:
            check_linear(a.astype(np.float64), want)
        assert np.array_equal(a, computer.compute_full(sig))  # batches and single signals: same bits
    pcm = rng.integers(-20000, 20000, 30000).astype(np.int16)
    a16 = computer.compute_batch([pcm])[0]
    assert np.array_equal(a16, computer.compute_batch([pcm.astype(np.float32)])[0])
    for kwargs in (dict(preemph=0.97), dict(dither=2.0, preemph=0.97, dither_first=True, seed=7),
                   dict(dither=1.0, seed=3)):
        monkeypatch.setenv("PDS_STFT_KERNEL", "u")
        a = computer.compute_batch(signals[4:8] + [pcm], **kwargs)
        monkeypatch.delenv("PDS_STFT_KERNEL")
        b = computer.compute_batch(signals[4:8] + [pcm], **kwargs)
        for x, y in zip(a, b):
            if computer._log:
                assert np.abs(x - y).max() <= 2e-4
            else:
                assert np.allclose(x, y, rtol=1e-4, atol=1e-4 * np.abs(y).max())
