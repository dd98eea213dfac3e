"""The C-ABI library: it builds, loads, and exports every symbol include/pds_b200.h declares.
No compute calls are made here (no GPU in the CPU test tier)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pds_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pds_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from pydrobert_speech_b200 import _lib

    assert declared_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from pydrobert_speech_b200 import _lib

    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.get_lib()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.pds_version() >= 100
    assert isinstance(lib.pds_last_error(), bytes)
    assert ctypes.sizeof(_lib.PdsTile) == 32


def test_no_cpu_fallback(speech):
    """Without a CUDA device the product must fail loudly rather than compute on the host"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pydrobert_speech_b200._lib import PdsError

    computer = speech.compute.STFTFrameComputer("fbank", frame_length_ms=25)
    with pytest.raises(PdsError, match="no CPU fallback"):
        computer.compute_full(np.zeros(16000))
    assert computer.compute_full(np.zeros(10)).shape == (0, 40)  # too short: no device work at all


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pydrobert-speech_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name


def test_fft_emulation_matches_numpy(tmp_path):
    """csrc/emu_fft.cpp runs the device FFT templates and index maps on the CPU"""
    src = os.path.join(ROOT, "pydrobert-speech_b200", "csrc", "emu_fft.cpp")
    lib_file = str(tmp_path / "libpds_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib_file, src])
    lib = ctypes.CDLL(lib_file)
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(0)
    for N, L in [(128, 100), (256, 200), (512, 400), (512, 512), (512, 33), (1024, 801), (2048, 1102)]:
        for power in (1, 0):
            x = (rng.standard_normal(L) * 1000).astype(np.float32)
            w = np.hanning(L).astype(np.float32)
            P = np.full(N // 2 + 1, np.nan, np.float32)
            assert lib.pds_emu_frame_spectrum(N, x.ctypes.data_as(fp), L, w.ctypes.data_as(fp),
                                              P.ctypes.data_as(fp), power) == 0
            X = np.abs(np.fft.rfft(x.astype(np.float64) * w, n=N))
            want = X ** 2 if power else X
            assert not np.isnan(P).any()
            assert np.abs(P - want).max() <= 2e-6 * want.max()
