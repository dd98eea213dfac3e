"""Guard-band checks of every kernel family: no write lands outside the output, and nothing read
outside the input reaches a result.

Each kernel runs twice: once on plain tensors, once with its input sitting in the middle of a
NaN-filled buffer and its output in the middle of a sentinel-filled one.  The two results must be
bit-identical (a read past either end of the input would turn some output into NaN -- masked loads
of the neighbouring bytes are fine, using them is not) and the sentinels must survive (a write
outside the output would flip one).  This is the in-suite stand-in for a memory checker.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096  # elements on either side; a multiple of 4 keeps the inner view 16-byte aligned
SENTINEL = -12345.0

LENGTHS = [0, 1, 201, 399, 5000, 16000, 33333, 160 * 32 + 240, 48000, 7]
FBANK = {"name": "stft", "bank": "fbank", "frame_length_ms": 25, "include_energy": True,
         "pad_to_nearest_power_of_two": True, "window_function": "hanning", "use_power": True}
GAMMATONE = {"name": "stft", "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 64, "erb": True},
             "frame_length_ms": 25, "use_power": True}


def guarded_input(torch, host, fill=float("nan")):
    """`host` copied into the middle of a buffer filled with `fill`; returns (whole, inner view)"""
    n = host.numel()
    whole = torch.full((n + 2 * GUARD,), fill, dtype=host.dtype, device="cuda")
    whole[GUARD:GUARD + n] = host.to("cuda").reshape(-1)
    return whole, whole[GUARD:GUARD + n].view(host.shape)


def guarded_output(torch, shape):
    n = int(np.prod(shape))
    whole = torch.full((n + 2 * GUARD,), SENTINEL, dtype=torch.float32, device="cuda")
    return whole, whole[GUARD:GUARD + n].view(shape)


def sentinels_intact(whole, n):
    return bool((whole[:GUARD] == SENTINEL).all()) and bool((whole[GUARD + n:] == SENTINEL).all())


def signals(dtype=np.float32):
    rng = np.random.default_rng(11)
    return [(rng.standard_normal(n) * 1000).astype(dtype) for n in LENGTHS]


STFT_CASES = [
    ("default", FBANK, {}),
    ("tc", FBANK, {"PDS_STFT_KERNEL": "1"}),
    ("tf32_bank", FBANK, {"PDS_STFT_BANK": "tf32"}),
    ("pipe", dict(FBANK, include_energy=False), {"PDS_STFT_KERNEL": "p"}),
    ("scalar", GAMMATONE, {"PDS_STFT_KERNEL": "scalar"}),
    ("ws", GAMMATONE, {"PDS_STFT_KERNEL": "ws"}),
    ("gammatone", GAMMATONE, {}),
    ("dft1024", dict(FBANK, frame_length_ms=64, frame_shift_ms=16), {}),
    ("dft2048", dict(FBANK, frame_length_ms=100, frame_shift_ms=20), {}),
    ("dft256", dict(FBANK, frame_length_ms=12.5, frame_shift_ms=5), {}),
    ("odd_shift", dict(FBANK, frame_shift_ms=10.0625), {}),
    ("bluestein400", dict(FBANK, pad_to_nearest_power_of_two=False), {}),
    ("bluestein200", dict(FBANK, frame_length_ms=12.5, pad_to_nearest_power_of_two=False), {}),
    ("direct800", dict(FBANK, frame_length_ms=50, pad_to_nearest_power_of_two=False), {}),
]


@pytest.mark.parametrize("name,cfg,env", STFT_CASES, ids=[c[0] for c in STFT_CASES])
@pytest.mark.parametrize("pcm", (False, True), ids=("f32", "i16"))
@pytest.mark.parametrize("pre", (False, True), ids=("plain", "dither_preemph"))
def test_stft_kernels_stay_inside_their_buffers(speech, monkeypatch, name, cfg, env, pcm, pre):
    import torch

    for key in ("PDS_STFT_KERNEL", "PDS_STFT_BANK"):
        monkeypatch.delenv(key, raising=False)
    for key, value in env.items():
        monkeypatch.setenv(key, value)
    computer = speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)
    dtype = np.int16 if pcm else np.float32
    packed = speech.compute.PackedSignals.pack(signals(dtype), dtype, computer.pad_left % 4)
    kwargs = dict(preemph=0.97, dither=1.0, seed=5) if pre else {}
    layout = computer.plan_batch(packed.offsets, packed.lengths, torch.device("cuda", 0))
    host = torch.from_numpy(packed.data)
    plain = computer.run_batch(layout, host.to("cuda"), **kwargs)
    fill = float("nan") if not pcm else 32767
    whole_in, inner_in = guarded_input(torch, host, fill)
    whole_out, inner_out = guarded_output(torch, (layout.rows, computer.num_coeffs))
    got = computer.run_batch(layout, inner_in, out=inner_out, **kwargs)
    torch.cuda.synchronize()
    assert got.data_ptr() == inner_out.data_ptr()
    assert sentinels_intact(whole_out, inner_out.numel()), f"{name}: write outside the output"
    assert not bool((inner_out == SENTINEL).any()), f"{name}: rows left unwritten"
    assert torch.isfinite(inner_out).all(), f"{name}: a guard value reached a result"
    assert torch.equal(plain, inner_out), f"{name}: result depends on what surrounds the input"
    if pcm:  # int16 has no NaN: move the surroundings and ask for the same bits again
        whole_in[:GUARD] = -32768
        whole_in[GUARD + host.numel():] = -32768
        again = computer.run_batch(layout, inner_in, **kwargs)
        assert torch.equal(plain, again), f"{name}: result depends on what surrounds the input"


@pytest.mark.parametrize("kernel", (None, "direct"))
def test_si_kernels_stay_inside_their_buffers(speech, monkeypatch, kernel):
    import torch

    monkeypatch.delenv("PDS_SI_KERNEL", raising=False)
    if kernel:
        monkeypatch.setenv("PDS_SI_KERNEL", kernel)
    computer = speech.alias_factory_subclass_from_arg(
        speech.compute.FrameComputer, {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}})
    packed = speech.compute.PackedSignals.pack(signals(), np.float32, 0)
    host = torch.from_numpy(packed.data)
    plain, frame_off = computer.compute_packed_device(host.to("cuda"), packed.offsets, packed.lengths)
    whole_in, inner_in = guarded_input(torch, host)
    got, _ = computer.compute_packed_device(inner_in, packed.offsets, packed.lengths)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert torch.equal(plain, got)
    # the output side: run the launch by hand into a guarded buffer
    import ctypes

    from pydrobert_speech_b200._gpu import TILE_DTYPE
    from pydrobert_speech_b200._lib import check, get_lib

    lib, plan = get_lib(), computer._plan(torch.device("cuda", 0))
    i64p = ctypes.POINTER(ctypes.c_int64)
    n_tiles = ctypes.c_int64(0)
    offs = np.zeros(len(packed.lengths) + 1, dtype=np.int64)
    check(lib.pds_si_layout(plan.handle, len(packed.lengths), packed.lengths.ctypes.data_as(i64p),
                            offs.ctypes.data_as(i64p), ctypes.byref(n_tiles)))
    tiles = np.empty(n_tiles.value, dtype=TILE_DTYPE)
    check(lib.pds_si_fill_tiles(plan.handle, len(packed.lengths), packed.offsets.ctypes.data_as(i64p),
                                packed.lengths.ctypes.data_as(i64p), offs.ctypes.data_as(i64p), tiles.ctypes.data))
    whole_out, inner_out = guarded_output(torch, tuple(plain.shape))
    computer._launch(plan, inner_in, tiles, inner_out)
    torch.cuda.synchronize()
    assert sentinels_intact(whole_out, inner_out.numel())
    assert torch.equal(plain, inner_out)


@pytest.mark.parametrize("staged", (False, True), ids=("streaming", "staged"))
@pytest.mark.parametrize("cols", (41, 40, 13, 300))
def test_post_kernels_stay_inside_their_buffers(speech, monkeypatch, staged, cols):
    import torch

    monkeypatch.delenv("PDS_DELTAS_KERNEL", raising=False)
    if staged:
        monkeypatch.setenv("PDS_DELTAS_KERNEL", "s")
    rng = np.random.default_rng(3)
    rows = 3001
    host = torch.from_numpy(rng.standard_normal((rows, cols)).astype(np.float32))
    row_off = torch.tensor([0, 7, 7, 8, 1000, 1001, 2990, rows], dtype=torch.int64, device="cuda")
    whole_in, inner_in = guarded_input(torch, host)
    for deltas in (speech.post.Deltas(2), speech.post.Deltas(1), speech.post.Deltas(3, context_window=3)):
        plain = deltas.apply_device(host.to("cuda"), row_off)
        whole_out, inner_out = guarded_output(torch, tuple(plain.shape))
        deltas.apply_device(inner_in, row_off, out=inner_out)
        torch.cuda.synchronize()
        assert sentinels_intact(whole_out, inner_out.numel())
        assert torch.equal(plain, inner_out)
    # Deltas(2) fused with CMVN: statistics and the standardised result
    deltas = speech.post.Deltas(2)
    want_cmvn = speech.post.Standardize()
    want_cmvn.accumulate_device(deltas.lazy_device(host.to("cuda"), row_off))
    want = want_cmvn.apply_device(deltas.lazy_device(host.to("cuda"), row_off))
    cmvn = speech.post.Standardize()
    cmvn.accumulate_device(deltas.lazy_device(inner_in, row_off))
    whole_out, inner_out = guarded_output(torch, tuple(want.shape))
    cmvn.apply_device(deltas.lazy_device(inner_in, row_off), out=inner_out)
    torch.cuda.synchronize()
    # float64 atomics: the order of the additions, hence the last bits, may differ between runs
    np.testing.assert_allclose(cmvn._stats, want_cmvn._stats, rtol=1e-11)
    assert sentinels_intact(whole_out, inner_out.numel())
    assert torch.allclose(want, inner_out, rtol=1e-5, atol=1e-6)
    # plain CMVN over a tensor
    feats = deltas.apply_device(host.to("cuda"), row_off)
    whole_in2, inner_in2 = guarded_input(torch, feats.cpu())
    plain_cmvn = speech.post.Standardize()
    plain_cmvn.accumulate_device(feats)
    guarded_cmvn = speech.post.Standardize()
    guarded_cmvn.accumulate_device(inner_in2)
    np.testing.assert_allclose(guarded_cmvn._stats, plain_cmvn._stats, rtol=1e-11)
    whole_out, inner_out = guarded_output(torch, tuple(feats.shape))
    guarded_cmvn.apply_device(inner_in2, out=inner_out)
    torch.cuda.synchronize()
    assert sentinels_intact(whole_out, inner_out.numel())
    assert torch.allclose(plain_cmvn.apply_device(feats), inner_out, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("entry,args", (("pds_preemphasize", (0.97,)), ("pds_dither", (1.0, 9))))
def test_stand_alone_pre_processing_stays_inside(speech, entry, args):
    import torch

    from pydrobert_speech_b200._gpu import stream_ptr
    from pydrobert_speech_b200._lib import check, get_lib

    rng = np.random.default_rng(5)
    device = torch.device("cuda", 0)
    lengths_h = [1, 2, 0, 1023, 1024, 40001, 5]
    offsets_h = np.concatenate([[0], np.cumsum(lengths_h)[:-1]])
    total = int(sum(lengths_h))
    host = torch.from_numpy((rng.standard_normal(total) * 100).astype(np.float32))
    offsets = torch.from_numpy(offsets_h.astype(np.int64)).to(device)
    lengths = torch.tensor(lengths_h, dtype=torch.int64, device=device)

    def run(d_in, d_out):
        check(getattr(get_lib(), entry)(d_in.data_ptr(), d_out.data_ptr(), len(lengths_h), offsets.data_ptr(),
                                        lengths.data_ptr(), total, *args, stream_ptr(device)))
        torch.cuda.synchronize()
        return d_out

    plain = run(host.to(device), torch.empty(total, dtype=torch.float32, device=device))
    whole_in, inner_in = guarded_input(torch, host)
    whole_out, inner_out = guarded_output(torch, (total,))
    run(inner_in, inner_out)
    assert sentinels_intact(whole_out, total)
    assert torch.isfinite(inner_out).all()
    assert torch.equal(plain, inner_out)
