"""The PyTorch modules and the signals-to-torch-feat-dir command, mirroring the reference's
tests/test_torch.py and tests/test_command_line.py:89-179 (which pass unmodified semantics)."""
import json
import os
import wave

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu


def build(speech, cfg):
    return speech.alias_factory_subclass_from_arg(speech.compute.FrameComputer, cfg)


@pytest.mark.parametrize("include_energy", [True, False])
def test_pytorch_stft_frame_computer(speech, include_energy):
    import torch

    from pydrobert_speech_b200.torch import PyTorchSTFTFrameComputer, pytorch_stft_frame_computer

    torch.manual_seed(1)
    signal = torch.randn(16000)
    cfg = dict(cases.KALDI_FBANK, include_energy=include_energy)
    computer = build(speech, cfg)
    module = PyTorchSTFTFrameComputer.from_stft_frame_computer(computer)
    exp = computer.compute_full(signal.numpy())
    act = module(signal)
    assert act.device.type == "cpu" and act.dtype == torch.float32
    assert np.allclose(exp, act.detach().numpy(), atol=1e-5)
    on_gpu = module(signal.cuda().double())
    assert on_gpu.is_cuda and on_gpu.dtype == torch.float64
    assert np.allclose(exp, on_gpu.detach().cpu().numpy(), atol=1e-5)
    # built from explicit tables (float32 / complex64 like the reference's from_* default)
    rebuilt = PyTorchSTFTFrameComputer(
        list(zip(module.offsets, module.filters)), module.frame_length, module.frame_shift, "centered",
        module.window, module.dft_size, module.use_log, module.use_power, module.include_energy,
        module.kaldi_shift, module.is_real)
    assert np.allclose(exp, rebuilt(signal).detach().numpy(), atol=1e-4)
    func = pytorch_stft_frame_computer(
        signal, module.filters, list(module.offsets), module.frame_length, module.frame_shift, True,
        module.window, module.dft_size, module.use_log, module.use_power, module.include_energy,
        module.kaldi_shift, module.is_real)
    assert np.allclose(exp, func.detach().numpy(), atol=1e-4)
    assert module(torch.zeros(10)).shape == (0, len(module.offsets))  # as the reference, torch.py:179-180
    with pytest.raises(RuntimeError, match="1-dimensional"):
        module(torch.zeros(3, 100))
    with pytest.raises(ValueError, match="dft_size"):
        PyTorchSTFTFrameComputer([(0, torch.ones(3))], 400, 160, dft_size=256)


def test_pytorch_pre_and_post(speech):
    import torch

    from pydrobert_speech_b200 import torch as pt

    torch.manual_seed(2)
    dither = pt.PyTorchDither.from_dither(speech.pre.Dither(2.0))
    noisy = dither(torch.zeros(1_000_000))
    assert torch.isclose(noisy.mean(), torch.tensor(0.0), atol=1e-2)
    assert torch.isclose(noisy.std(), torch.tensor(2.0), atol=1e-2)
    torch.manual_seed(2)
    assert torch.equal(noisy, dither(torch.zeros(1_000_000)))
    signal = torch.randn(1000, dtype=torch.float64)
    emph = pt.PyTorchPreemphasize.from_preemphasize(speech.pre.Preemphasize(0.9))(signal)
    assert np.allclose(emph.numpy(), oracle.preemphasize(signal.numpy(), 0.9), atol=1e-5)
    feats = torch.randn(50, 8)
    wrapped = pt.PyTorchPostProcessorWrapper.from_postprocessor(speech.post.Deltas(2))(feats)
    assert wrapped.shape == (50, 24)
    assert np.allclose(wrapped.numpy(), speech.post.Deltas(2).apply(feats.numpy()), atol=1e-6)
    si = build(speech, cases.SI_CASES["si_gabor_energy_power"][0])
    module = pt.PyTorchSIFrameComputer.from_si_frame_computer(si)
    sig = torch.randn(3000)
    assert np.allclose(module(sig).numpy(), si.compute_full(sig.numpy()), atol=1e-5)
    with pytest.raises(NotImplementedError):
        module.state_dict()


def test_signals_to_torch_feat_dir(speech, tmp_path):
    import torch

    from pydrobert_speech_b200 import command_line

    torch.manual_seed(50)
    feat_dir, raw_dir = str(tmp_path / "feat"), str(tmp_path / "raw")
    map_path, manifest_path = str(tmp_path / "map"), str(tmp_path / "manifest.txt")
    computer_path, pre_path = str(tmp_path / "fbank.json"), str(tmp_path / "pre.json")
    with open(computer_path, "w") as f:
        json.dump(cases.KALDI_FBANK, f)
    with open(pre_path, "w") as f:
        f.write('["dither"]\n')
    os.makedirs(raw_dir)
    num_utts, utt2signal, utt_ids = 100, dict(), []
    with open(map_path, "w") as mp:
        for idx in range(num_utts):
            utt_id = "utt{:03d}".format(idx)
            utt_ids.append(utt_id)
            n = torch.randint(1, 1600, (1,)).item()
            signal = torch.randint(-(2 ** 15), 2 ** 15 - 1, (n,), dtype=torch.float32)
            utt2signal[utt_id] = signal
            kind = idx % 3
            if kind == 2:
                path = os.path.join(raw_dir, f"{idx}.wav")
                with wave.open(path, "wb") as wv:
                    wv.setnchannels(1)
                    wv.setsampwidth(2)
                    wv.setframerate(16000)
                    wv.writeframes(signal.to(torch.int16).numpy().tobytes())
            elif kind == 1:
                path = os.path.join(raw_dir, f"{idx}.npy")
                np.save(path, signal.numpy())
            else:
                path = os.path.join(raw_dir, f"{idx}.pt")
                torch.save(signal, path)
            mp.write(f"{utt_id} {path}\n")
    args = [map_path, computer_path, feat_dir]
    assert not command_line.signals_to_torch_feat_dir(args)
    computer = build(speech, cases.KALDI_FBANK)
    for utt_id in utt_ids:
        feat = torch.load(os.path.join(feat_dir, f"{utt_id}.pt"))
        assert feat.dtype == torch.float32 and feat.shape[-1] == 40
        want = computer.compute_full(utt2signal[utt_id].numpy())
        assert feat.shape == want.shape and np.allclose(feat.numpy(), want, atol=1e-5)
    # a small batch size must not change anything
    assert not command_line.signals_to_torch_feat_dir(args + ["--batch-samples=3000", "--num-workers=3"])
    for utt_id in utt_ids[::7]:
        feat = torch.load(os.path.join(feat_dir, f"{utt_id}.pt"))
        assert np.allclose(feat.numpy(), computer.compute_full(utt2signal[utt_id].numpy()), atol=1e-5)
    args.pop(1)  # no computer: store the raw audio as (S, 1)
    assert not command_line.signals_to_torch_feat_dir(args)
    for utt_id, exp in utt2signal.items():
        act = torch.load(os.path.join(feat_dir, f"{utt_id}.pt"))
        assert act.shape == (len(exp), 1) and torch.allclose(exp, act.flatten())
    args += ["--seed=1", f"--preprocess={pre_path}"]
    assert not command_line.signals_to_torch_feat_dir(args)
    for utt_id in utt_ids:
        noisy = torch.load(os.path.join(feat_dir, f"{utt_id}.pt"))
        assert not torch.allclose(utt2signal[utt_id], noisy.flatten())
        utt2signal[utt_id] = noisy
    args += ["--num-workers=2", f"--manifest={manifest_path}"]
    assert not command_line.signals_to_torch_feat_dir(args)  # same seed: same noise
    for utt_id, exp in utt2signal.items():
        assert torch.allclose(exp, torch.load(os.path.join(feat_dir, f"{utt_id}.pt")))
    with open(manifest_path) as f:
        utts = [x.strip() for x in f]
    assert sorted(utts) == sorted(utt2signal)
    # utterances already in the manifest are not recomputed, missing ones are
    utt1, utt2 = utts[:2]
    exp1, exp2 = utt2signal[utt1], torch.randn_like(utt2signal[utt1])
    torch.save(exp2, os.path.join(feat_dir, f"{utt1}.pt"))
    torch.save(exp2, os.path.join(feat_dir, f"{utt2}.pt"))
    with open(manifest_path, "w") as f:
        f.write("\n".join(utts[1:]) + "\n")
    assert not command_line.signals_to_torch_feat_dir(args)
    assert torch.allclose(exp1, torch.load(os.path.join(feat_dir, f"{utt1}.pt")))
    assert torch.allclose(exp2, torch.load(os.path.join(feat_dir, f"{utt2}.pt")))


def test_cli_fused_preprocessing_and_postprocessing(speech, tmp_path):
    """dither + preemphasis fused into the kernel, Deltas applied with the CLI's axis=-1"""
    import torch

    from pydrobert_speech_b200 import command_line

    rng = np.random.default_rng(3)
    feat_dir, map_path = str(tmp_path / "feat"), str(tmp_path / "map")
    signals = {f"u{i}": (rng.standard_normal(n) * 300).astype(np.float32) for i, n in enumerate((4000, 9000, 250))}
    with open(map_path, "w") as mp:
        for utt, sig in signals.items():
            np.save(str(tmp_path / f"{utt}.npy"), sig)
            mp.write(f"{utt} {tmp_path / (utt + '.npy')}\n")
    args = [map_path, json.dumps(cases.README_FBANK), feat_dir, "--seed=5",
            '--preprocess=[{"name": "preemph", "coeff": 0.95}]', '--postprocess=[{"name": "deltas", "num_deltas": 1}]']
    assert command_line.main(["signals-to-torch-feat-dir"] + args) == 0
    computer = build(speech, cases.README_FBANK)
    for utt, sig in signals.items():
        feat = torch.load(os.path.join(feat_dir, f"{utt}.pt")).numpy()
        base = oracle.stft_features(
            oracle.preemphasize(sig, 0.95), computer._window, 512, computer._filt_start_idxs,
            computer._truncated_filts, 160, 199, True, True, True, True)
        want = np.concatenate([base, oracle.deltas(base.T, 1)[:, base.shape[0]:].T], axis=1)
        assert feat.shape == want.shape == (computer.num_frames(len(sig)), 82)
        assert np.abs(feat - want).max() <= 1e-3
    assert command_line.main(["nonsense"]) == 2
    assert command_line.signals_to_torch_feat_dir(["--help"]) == 0


def test_cli_sharded_ranks_write_identical_files(speech, tmp_path, monkeypatch):
    """torchrun-style launch (RANK / WORLD_SIZE / LOCAL_RANK): every rank computes a shard of the
    map and writes its own files.  With the same --seed the files of a 3-rank run are byte-identical
    to those of a 1-rank run (the dither stream is keyed by the position in the map), every utterance
    is written exactly once, and the shared manifest ends up complete."""
    import torch

    from pydrobert_speech_b200 import command_line

    rng = np.random.default_rng(8)
    map_path = str(tmp_path / "map")
    with open(map_path, "w") as mp:
        for i in range(37):
            n = int(rng.integers(300, 9000))
            if i % 2:
                path = str(tmp_path / f"s{i}.wav")
                with wave.open(path, "wb") as wv:
                    wv.setnchannels(1)
                    wv.setsampwidth(2)
                    wv.setframerate(16000)
                    wv.writeframes(rng.integers(-3000, 3000, n).astype(np.int16).tobytes())
            else:
                path = str(tmp_path / f"s{i}.npy")
                np.save(path, (rng.standard_normal(n) * 500).astype(np.float32))
            mp.write(f"utt{i:02d} {path}\n")
    common = [map_path, json.dumps(cases.README_FBANK)]
    tail = ["--seed=11", '--preprocess=[{"name": "dither", "coeff": 2.0}, "preemph"]', "--batch-samples=20000"]
    single, sharded = str(tmp_path / "single"), str(tmp_path / "sharded")
    report = str(tmp_path / "report.jsonl")
    assert command_line.signals_to_torch_feat_dir(common + [single] + tail) == 0
    manifest = str(tmp_path / "manifest")
    for rank in range(3):
        monkeypatch.setenv("WORLD_SIZE", "3")
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("LOCAL_RANK", "0")
        before = set(os.listdir(sharded)) if os.path.isdir(sharded) else set()
        assert command_line.signals_to_torch_feat_dir(
            common + [sharded] + tail + [f"--manifest={manifest}", f"--report={report}", "--num-workers=2"]) == 0
        written = set(os.listdir(sharded)) - before
        assert 0 < len(written) < 37  # a proper shard, disjoint from the other ranks'
    monkeypatch.delenv("WORLD_SIZE")
    monkeypatch.delenv("RANK")
    monkeypatch.delenv("LOCAL_RANK")
    names = sorted(os.listdir(single))
    assert names == sorted(os.listdir(sharded)) and len(names) == 37
    for name in names:
        a = torch.load(os.path.join(single, name))
        b = torch.load(os.path.join(sharded, name))
        assert a.shape == b.shape and torch.equal(a, b), name
    with open(manifest) as f:
        assert sorted(x.strip() for x in f) == [f"utt{i:02d}" for i in range(37)]
    with open(report) as f:
        lines = [json.loads(x) for x in f]
    assert [x["rank"] for x in lines] == [0, 1, 2] and sum(x["utterances"] for x in lines) == 37
    assert all(x["audio_hours_per_second"] > 0 and x["files_per_second"] > 0 for x in lines)
    # without a seed a multi-rank run with dither is refused (every rank would draw its own)
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    assert command_line.signals_to_torch_feat_dir(common + [sharded, tail[1]]) == 1


def test_pcm_wav_goes_to_the_kernel_as_int16(speech, tmp_path):
    """16-bit wav data is packed, copied and staged as int16 (half the PCIe bytes): same features as
    the float32 copy of the same samples"""
    rng = np.random.default_rng(21)
    pcm = [rng.integers(-20000, 20000, n).astype(np.int16) for n in (5000, 16001, 777)]
    computer = build(speech, cases.README_FBANK)
    from pydrobert_speech_b200.pipeline import FeaturePipeline

    pipe = FeaturePipeline(computer)
    as_pcm = pipe.run_list(pcm)
    assert pipe._pinned_buffers[("in", "int16")].numel() > 0
    as_float = pipe.run_list([s.astype(np.float32) for s in pcm])
    for a, b, s in zip(as_pcm, as_float, pcm):
        assert a.shape == b.shape == (computer.num_frames(len(s)), 41)
        assert np.array_equal(a, b)


def test_torch_module_state_dict_is_the_references(speech):
    """Parameter names and shapes follow the reference (torch.py:362-366: ``filters.<i>``, ``window``):
    a state_dict saved there loads here.  Loading new values rebuilds the kernel plan."""
    import torch

    import pydrobert_speech_b200.torch as pt

    computer = build(speech, cases.KALDI_FBANK)
    module = pt.PyTorchSTFTFrameComputer.from_stft_frame_computer(computer)
    state = module.state_dict()
    assert sorted(state) == sorted([f"filters.{i}" for i in range(40)] + ["window"])
    assert state["window"].shape == (400,) and state["filters.0"].dtype == torch.cfloat
    assert all(p.requires_grad for p in module.parameters())  # learnable, like the reference's
    sig = torch.randn(6000)
    before = module(sig)
    halved = {k: (v * 0.5 if k.startswith("filters") else v) for k, v in state.items()}
    module.load_state_dict(halved)  # power features: filters at half amplitude = -log(4)
    after = module(sig)
    assert torch.allclose(after, before - float(np.log(4.0)), atol=2e-3)
    # too short a signal: like the reference's functional form, (0, num_filts)
    energy = pt.PyTorchSTFTFrameComputer.from_stft_frame_computer(build(speech, cases.README_FBANK))
    assert energy(torch.zeros(10)).shape == (0, 40)
    assert energy(torch.zeros(5000)).shape[1] == 41
    assert module(torch.randn(3000, requires_grad=True)).grad_fn is not None  # differentiable (next test)
    with torch.no_grad():
        assert module(torch.randn(3000, requires_grad=True)).shape[1] == 40


def test_compute_feats_from_kaldi_tables(speech, tmp_path):
    """The second drop-in command (reference command_line.py:245-359, its test
    tests/test_command_line.py:21-86): wave scp in, float-matrix ark out, through the built-in
    Kaldi table I/O.  Features equal compute_full; post-processors are validated but not applied
    (as in the reference); the same --seed reproduces the dither."""
    from pydrobert_speech_b200 import _kaldi_io as kio
    from pydrobert_speech_b200 import command_line

    rng = np.random.default_rng(5)
    wav_scp, feat_ark = str(tmp_path / "wav.scp"), str(tmp_path / "feat.ark")
    signals = {}
    with open(wav_scp, "w") as scp:
        for i in range(40):
            utt = f"utt{i:02d}"
            n = int(rng.integers(0, 32000))
            pcm = rng.integers(-(2 ** 15), 2 ** 15 - 1, n).astype(np.int16)
            path = str(tmp_path / f"{utt}.wav")
            with wave.open(path, "wb") as wv:
                wv.setnchannels(1)
                wv.setsampwidth(2)
                wv.setframerate(16000 if i != 7 else 8000)  # one file at the wrong rate: skipped
                wv.writeframes(pcm.tobytes())
            signals[utt] = pcm
            scp.write(f"{utt} {path}\n")
    cfg = str(tmp_path / "fbank.json")
    with open(cfg, "w") as f:
        json.dump(cases.KALDI_FBANK, f)
    args = ["scp,s:" + wav_scp, "ark:" + feat_ark, cfg, "--postprocess=[\"unit\"]", "--batch-samples=100000"]
    assert command_line.compute_feats_from_kaldi_tables(args) == 0
    computer = build(speech, cases.KALDI_FBANK)
    got = dict(kio.read_matrix_table("ark:" + feat_ark))
    assert sorted(got) == sorted(u for u in signals if u != "utt07")
    for utt, feat in got.items():
        want = computer.compute_full(signals[utt].astype(np.float32))
        assert feat.dtype == np.float32 and feat.shape == want.shape
        assert np.array_equal(feat, want) and (feat.shape[0] == 0 or feat.shape[1] == 40)
    # dither: --seed makes it reproducible, and it changes the features
    noisy = args + ["--seed=30", '--preprocess=["dither"]']
    assert command_line.main(["compute-feats-from-kaldi-tables"] + noisy) == 0
    first = dict(kio.read_matrix_table("ark:" + feat_ark))
    assert command_line.compute_feats_from_kaldi_tables(noisy + ["--batch-samples=30000"]) == 0
    second = dict(kio.read_matrix_table("ark:" + feat_ark))
    assert all(np.array_equal(first[u], second[u]) for u in first)
    assert any(len(first[u]) and not np.array_equal(first[u], got[u]) for u in first)
    # unreadable table / bad computer config: return code 1, no exception
    assert command_line.compute_feats_from_kaldi_tables(["scp:" + str(tmp_path / "nope"), "ark:" + feat_ark, cfg]) == 1
    assert command_line.compute_feats_from_kaldi_tables(["scp:" + wav_scp, "ark:" + feat_ark, '{"name": "nonsense"}']) == 1


def _grad_case(golden, name):
    import torch

    data = golden("torch_grad")
    edges = data[name + "/filters/offsets"]
    filters = [torch.tensor(data[name + "/filters/values"][a:b]) for a, b in zip(edges[:-1], edges[1:])]
    grad_filters = [data[name + "/grad_filters/values"][a:b] for a, b in zip(edges[:-1], edges[1:])]
    frame_length, frame_shift, dft_size, centered = (int(v) for v in data[name + "/geometry"])
    kwargs = dict(frame_length=frame_length, frame_shift=frame_shift, dft_size=dft_size,
                  frame_style="centered" if centered else "causal",
                  use_log=name.endswith("energy"), use_power=name.endswith("energy"),
                  include_energy=name.endswith("energy"), kaldi_shift=False, is_real=True)
    return data, filters, grad_filters, [int(o) for o in data[name + "/offsets"]], kwargs


@pytest.mark.parametrize("name", ["fbank10_power_log_energy", "fbank8_magnitude_causal"])
def test_torch_module_is_differentiable(speech, golden, name):
    """forward = the fused kernel, backward = autograd through the plain-torch restatement: outputs and the
    gradients with respect to the signal, the window and every filter against the reference's OWN torch
    module (float64, tests/golden/make_golden.py torch_grad); an in-place parameter update (what an optimizer
    step does) reaches the next forward"""
    import torch

    import pydrobert_speech_b200.torch as pt

    data, filters, grad_filters, offsets, kwargs = _grad_case(golden, name)
    module = pt.PyTorchSTFTFrameComputer(
        list(zip(offsets, filters)), kwargs["frame_length"], kwargs["frame_shift"], kwargs["frame_style"],
        torch.tensor(data[name + "/window"]), kwargs["dft_size"], kwargs["use_log"], kwargs["use_power"],
        kwargs["include_energy"], kwargs["kaldi_shift"], kwargs["is_real"])
    signal = torch.tensor(data[name + "/signal"], dtype=torch.float32, device="cuda", requires_grad=True)
    feats = module(signal)
    want = data[name + "/feats"]
    assert feats.is_cuda and feats.grad_fn is not None and tuple(feats.shape) == want.shape
    got = feats.detach().cpu().numpy()
    if kwargs["use_log"]:
        assert np.abs(got - want).max() <= 1e-3
    else:
        assert (np.abs(got - want) / np.maximum(np.abs(want), 1e-6 * np.abs(want).max())).max() <= 1e-4
    feats.backward(torch.tensor(data[name + "/grad_out"], dtype=torch.float32, device="cuda"))

    def close(a, b):
        a = a.detach().cpu().numpy()
        return np.abs(a - b).max() <= 2e-4 * np.abs(b).max()

    assert close(signal.grad, data[name + "/grad_signal"])
    assert close(module.window.grad, data[name + "/grad_window"])
    for p, g in zip(module.filters, grad_filters):
        assert close(p.grad, g)
    # one step of plain SGD on the window, then the forward pass must see the new window
    before = module(signal.detach())
    with torch.no_grad():
        module.window.mul_(0.5)
    after = module(signal.detach())
    shift = float(np.log(4.0)) if kwargs["use_log"] else None
    cols = slice(1, None) if kwargs["include_energy"] else slice(None)
    if shift is not None:
        assert torch.allclose(after[:, cols], before[:, cols] - shift, atol=2e-3)
    else:
        assert torch.allclose(after[:, cols], before[:, cols] * 0.5, rtol=1e-3, atol=1e-3 * float(before.detach().abs().max()))
    with torch.no_grad():
        assert module(signal).grad_fn is None
