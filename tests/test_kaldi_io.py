"""The built-in Kaldi table I/O (``_kaldi_io``) against hand-built bytes of Kaldi's wire format
(kaldi-holder-inl.h / kaldi-matrix.cc): CPU only."""
import io
import os
import struct
import wave

import numpy as np
import pytest


def _wav_bytes(pcm: np.ndarray, rate=16000, channels=1) -> bytes:
    buf = io.BytesIO()
    with wave.open(buf, "wb") as wv:
        wv.setnchannels(channels)
        wv.setsampwidth(2)
        wv.setframerate(rate)
        wv.writeframes(pcm.astype("<i2").tobytes())
    return buf.getvalue()


def test_matrix_archive_bytes_and_scp(tmp_path, speech):
    from pydrobert_speech_b200 import _kaldi_io as kio

    ark, scp = str(tmp_path / "f.ark"), str(tmp_path / "f.scp")
    a = np.arange(6, dtype=np.float32).reshape(2, 3) + 0.5
    b = np.zeros((0, 3), dtype=np.float32)
    with kio.MatrixTableWriter(f"ark,scp:{ark},{scp}") as writer:
        writer.write("utt1", a)
        writer.write("utt2", b)
        with pytest.raises(ValueError):
            writer.write("bad key", a)
    want = (b"utt1 \0BFM \4" + struct.pack("<i", 2) + b"\4" + struct.pack("<i", 3) + a.tobytes()
            + b"utt2 \0BFM \4" + struct.pack("<i", 0) + b"\4" + struct.pack("<i", 3))
    with open(ark, "rb") as f:
        assert f.read() == want
    with open(scp) as f:
        assert f.read().splitlines() == [f"utt1 {ark}:5", f"utt2 {ark}:{5 + 15 + 24 + 5}"]
    for spec in (f"ark:{ark}", f"scp:{scp}"):
        back = list(kio.read_matrix_table(spec))
        assert [k for k, _ in back] == ["utt1", "utt2"]
        assert np.array_equal(back[0][1], a) and back[1][1].shape == (0, 3)
    # double and text flavours
    with kio.MatrixTableWriter(f"ark:{ark}", double=True) as writer:
        writer.write("d", a)
    with open(ark, "rb") as f:
        assert f.read() == b"d \0BDM \4" + struct.pack("<i", 2) + b"\4" + struct.pack("<i", 3) + a.astype("<f8").tobytes()
    with kio.MatrixTableWriter(f"ark,t:{ark}") as writer:
        writer.write("t1", a)
        writer.write("t2", a * 2)
    with open(ark) as f:
        assert f.read().startswith("t1  [\n  0.5  1.5  2.5\n  3.5  4.5  5.5 ]\n")
    back = dict(kio.read_matrix_table(f"ark:{ark}"))
    assert np.array_equal(back["t1"], a) and np.array_equal(back["t2"], a * 2)


def test_wave_tables(tmp_path, speech):
    from pydrobert_speech_b200 import _kaldi_io as kio

    rng = np.random.default_rng(0)
    mono = rng.integers(-3000, 3000, 1234).astype(np.int16)
    stereo = rng.integers(-3000, 3000, (500, 2)).astype(np.int16)
    ark = str(tmp_path / "w.ark")
    with open(ark, "wb") as f:  # Kaldi WaveHolder: key, space, the RIFF file itself
        f.write(b"a " + _wav_bytes(mono) + b"b " + _wav_bytes(stereo, 8000, 2))
    with kio.WaveTableReader(f"ark:{ark}") as reader:
        items = list(reader.items())
    assert [k for k, _ in items] == ["a", "b"]
    data, rate, dur = items[0][1]
    assert data.shape == (1, 1234) and rate == 16000 and dur == pytest.approx(1234 / 16000)
    assert np.array_equal(data[0], mono.astype(np.float32))
    data, rate, dur = items[1][1]
    assert data.shape == (2, 500) and rate == 8000 and np.array_equal(data, stereo.T.astype(np.float32))
    scp = str(tmp_path / "w.scp")
    path = str(tmp_path / "m.wav")
    with open(path, "wb") as f:
        f.write(_wav_bytes(mono))
    with open(scp, "w") as f:
        f.write(f"m {path}\n\n")
    with kio.WaveTableReader(f"scp,s,cs:{scp}") as reader:
        (key, (data, rate, _)), = list(reader.items())
    assert key == "m" and np.array_equal(data[0], mono.astype(np.float32))
    with pytest.raises(IOError):
        kio.WaveTableReader(f"scp:{tmp_path / 'missing.scp'}")
    with pytest.raises(IOError):
        kio.parse_specifier("ark:gunzip -c x.ark.gz |")
    assert kio.parse_specifier("ark,scp,t:a,b") == (["ark", "scp"], ["t"], ["a", "b"])
