"""Minimal Kaldi table I/O for ``compute-feats-from-kaldi-tables``.

The reference delegates table handling to the optional package pydrobert-kaldi
(``command_line.py:245-359``: ``kaldi_open(rspec, "wm", value_style="bsd")`` for the wave table,
``kaldi_open(wspec, "bm", mode="w")`` for the feature table).  That package is not part of this
build, so the two table types the command touches are implemented here from Kaldi's published
wire format (kaldi/src/util/kaldi-holder-inl.h, kaldi/src/matrix/kaldi-matrix.cc):

* wave tables (read): ``scp[,opts]:<file>`` of ``<key> <path-to-wav>`` lines, or ``ark[,opts]:<file>``
  in which every entry is ``<key><space>`` followed by a RIFF/WAVE file (PCM 8/16/24/32 bit or
  IEEE float).  Entries decode to ``(data (channels, samples) float32, sample rate, duration)`` --
  the ``"bsd"`` value style of pydrobert-kaldi.
* float matrix tables (write, and read back for tests): ``ark[,scp][,t]:<ark>[,<scp>]``.  Binary
  entries are ``<key><space>\\0B`` + ``FM `` (float32) or ``DM `` (float64) + ``\\4<int32 rows>\\4<int32
  cols>`` + row-major data; text entries are ``<key>  [`` rows ``]``.  With ``scp`` every entry also
  gets a ``<key> <ark>:<offset>`` line pointing at the byte after the key.

Options other than ``t`` (``s``, ``cs``, ``o``, ``p``, ...) are accepted and ignored: they tune Kaldi's
error handling, not the format.  Pipes (``cmd |``) are not supported.
"""

import io
import os
import struct

from typing import Iterator, List, Optional, Tuple

import numpy as np

__all__ = ["parse_specifier", "WaveTableReader", "MatrixTableWriter", "read_matrix_table"]


def parse_specifier(spec: str) -> Tuple[List[str], List[str], List[str]]:
    """``"ark,scp,t:a.ark,a.scp"`` -> ``(["ark", "scp"], ["t"], ["a.ark", "a.scp"])``"""
    if ":" not in spec:
        raise IOError(f"'{spec}' is not a Kaldi table specifier (expected <type>[,<opts>]:<path>)")
    head, tail = spec.split(":", 1)
    words = [w.strip() for w in head.split(",")]
    kinds = [w for w in words if w in ("ark", "scp")]
    options = [w for w in words if w not in ("ark", "scp")]
    if not kinds or len(kinds) != len(set(kinds)):
        raise IOError(f"'{spec}': expected 'ark' and/or 'scp' before the colon")
    paths = tail.split(",") if len(kinds) == 2 else [tail]
    if len(paths) != len(kinds) or any(p.strip().endswith("|") or p.strip().startswith("|") for p in paths):
        raise IOError(f"'{spec}': one plain file per table type is supported (no pipes)")
    return kinds, options, [p.strip() for p in paths]


# ---- RIFF/WAVE ---------------------------------------------------------------------------------
def _read_exact(handle, count: int) -> bytes:
    data = handle.read(count)
    if len(data) != count:
        raise IOError("unexpected end of file inside a wave entry")
    return data


def _read_wave(handle) -> Tuple[np.ndarray, float, float]:
    """One RIFF/WAVE object from the current position -> ((channels, samples) float32, rate, seconds)"""
    riff, _, wave_tag = struct.unpack("<4sI4s", _read_exact(handle, 12))
    if riff != b"RIFF" or wave_tag != b"WAVE":
        raise IOError("not a RIFF/WAVE object")
    fmt = None
    while True:
        tag, size = struct.unpack("<4sI", _read_exact(handle, 8))
        if tag == b"fmt ":
            body = _read_exact(handle, size + (size & 1))
            code, channels, rate, _, block, bits = struct.unpack("<HHIIHH", body[:16])
            if code == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: the real code leads the GUID
                code = struct.unpack("<H", body[24:26])[0]
            fmt = (code, channels, rate, block, bits)
        elif tag == b"data":
            if fmt is None:
                raise IOError("wave 'data' chunk before 'fmt '")
            code, channels, rate, block, bits = fmt
            if size in (0, 0xFFFFFFFF):  # streamed header: the data runs to the end of the file
                raw = handle.read()
            else:
                raw = _read_exact(handle, size)
                if size & 1:
                    handle.read(1)
            if code == 1 and bits == 16:
                data = np.frombuffer(raw, dtype="<i2").astype(np.float32)
            elif code == 1 and bits == 8:
                data = np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0
            elif code == 1 and bits == 32:
                data = np.frombuffer(raw, dtype="<i4").astype(np.float32)
            elif code == 1 and bits == 24:
                b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
                data = ((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8 >> 8).astype(np.float32)
            elif code == 3 and bits == 32:
                data = np.frombuffer(raw, dtype="<f4").astype(np.float32)
            elif code == 3 and bits == 64:
                data = np.frombuffer(raw, dtype="<f8").astype(np.float32)
            else:
                raise IOError(f"unsupported wave encoding (format {code}, {bits} bits)")
            if channels < 1 or len(data) % channels:
                raise IOError("number of channels does not evenly divide the wave samples")
            data = np.ascontiguousarray(data.reshape(-1, channels).T)
            return data, float(rate), data.shape[1] / float(rate)
        else:
            handle.seek(size + (size & 1), io.SEEK_CUR)


def _read_key(handle) -> Optional[str]:
    """The whitespace-terminated key that starts an archive entry; None at the end of the archive"""
    chars = []
    while True:
        ch = handle.read(1)
        if not ch:
            if chars:
                raise IOError("archive ends inside a key")
            return None
        if ch in b" \t":
            if chars:
                return b"".join(chars).decode()
            continue
        if ch in b"\r\n":
            if chars:
                raise IOError("newline inside an archive key")
            continue
        chars.append(ch)


class WaveTableReader:
    """Sequential reader of a wave table: iterate ``(key, (data, rate, duration))``"""

    def __init__(self, rspecifier: str):
        kinds, _, paths = parse_specifier(rspecifier)
        if len(kinds) != 1:
            raise IOError(f"'{rspecifier}': a read specifier names one table")
        self._kind, self._path = kinds[0], paths[0]
        self._handle = open(self._path, "rb")  # raises IOError for unreadable tables, like kaldi_open

    def items(self) -> Iterator[Tuple[str, Tuple[np.ndarray, float, float]]]:
        if self._kind == "scp":
            for line_no, line in enumerate(self._handle):
                line = line.decode().strip()
                if not line:
                    continue
                fields = line.split(None, 1)
                if len(fields) != 2:
                    raise IOError(f"{self._path}:{line_no + 1}: expected '<key> <path>'")
                with open(fields[1].strip(), "rb") as wav:
                    yield fields[0], _read_wave(wav)
        else:
            while True:
                key = _read_key(self._handle)
                if key is None:
                    return
                yield key, _read_wave(self._handle)

    def __iter__(self):
        return (value for _, value in self.items())

    def close(self) -> None:
        self._handle.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- float matrices ----------------------------------------------------------------------------
class MatrixTableWriter:
    """Writer of a Kaldi float-matrix table (``BaseFloat`` = float32 unless ``double=True``)"""

    is_double = False

    def __init__(self, wspecifier: str, double: bool = False):
        kinds, options, paths = parse_specifier(wspecifier)
        self._text = "t" in options
        self.is_double = bool(double)
        by_kind = dict(zip(kinds, paths))
        if "ark" not in by_kind:
            raise IOError(f"'{wspecifier}': writing needs an archive ('ark:' or 'ark,scp:')")
        self._ark_path = by_kind["ark"]
        self._ark = open(self._ark_path, "wb")
        self._scp = open(by_kind["scp"], "w") if "scp" in by_kind else None

    def write(self, key: str, matrix: np.ndarray) -> None:
        if not key or any(c.isspace() for c in key):
            raise ValueError(f"invalid table key '{key}'")
        dtype = np.float64 if self.is_double else np.float32
        matrix = np.ascontiguousarray(np.atleast_2d(np.asarray(matrix)), dtype=dtype)
        if matrix.ndim != 2:
            raise ValueError("expected a matrix")
        self._ark.write(key.encode() + b" ")
        if self._scp is not None:
            self._scp.write(f"{key} {self._ark_path}:{self._ark.tell()}\n")
        if self._text:
            rows = ["  ".join(repr(float(v)) for v in row) for row in matrix]
            body = " [" + ("\n  " + "\n  ".join(rows) if rows else "") + " ]\n"
            self._ark.write(body.encode())
        else:
            self._ark.write(b"\0B" + (b"DM " if self.is_double else b"FM "))
            self._ark.write(b"\4" + struct.pack("<i", matrix.shape[0]) + b"\4" + struct.pack("<i", matrix.shape[1]))
            self._ark.write(matrix.astype("<f8" if self.is_double else "<f4", copy=False).tobytes())

    def close(self) -> None:
        self._ark.close()
        if self._scp is not None:
            self._scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _read_matrix(handle) -> np.ndarray:
    lead = handle.read(2)
    if lead == b"\0B":
        token = _read_exact(handle, 3)
        if token not in (b"FM ", b"DM "):
            raise IOError(f"unsupported matrix token {token!r} (compressed matrices are not handled)")
        if _read_exact(handle, 1) != b"\4":
            raise IOError("malformed matrix header")
        rows = struct.unpack("<i", _read_exact(handle, 4))[0]
        if _read_exact(handle, 1) != b"\4":
            raise IOError("malformed matrix header")
        cols = struct.unpack("<i", _read_exact(handle, 4))[0]
        dtype = "<f4" if token == b"FM " else "<f8"
        data = np.frombuffer(_read_exact(handle, rows * cols * np.dtype(dtype).itemsize), dtype=dtype)
        return data.reshape(rows, cols).copy()
    # text mode: " [ r00 r01 ...\n  r10 ... ]"
    text = lead
    while not text.rstrip().endswith(b"]"):
        ch = handle.read(1)
        if not ch:
            raise IOError("archive ends inside a text matrix")
        text += ch
    handle.readline()
    body = text.decode().strip()
    if not body.startswith("[") or not body.endswith("]"):
        raise IOError("malformed text matrix")
    rows = [r.split() for r in body[1:-1].strip().split("\n") if r.strip()]
    if not rows:
        return np.zeros((0, 0), dtype=np.float32)
    return np.array([[float(v) for v in r] for r in rows], dtype=np.float32)


def read_matrix_table(rspecifier: str) -> Iterator[Tuple[str, np.ndarray]]:
    """``(key, matrix)`` pairs of an ``ark:`` or ``scp:`` float-matrix table (used by the tests)"""
    kinds, _, paths = parse_specifier(rspecifier)
    if kinds == ["ark"]:
        with open(paths[0], "rb") as handle:
            while True:
                key = _read_key(handle)
                if key is None:
                    return
                yield key, _read_matrix(handle)
    elif kinds == ["scp"]:
        with open(paths[0]) as scp:
            for line in scp:
                if not line.strip():
                    continue
                key, where = line.split(None, 1)
                path, offset = where.strip().rsplit(":", 1)
                with open(path, "rb") as handle:
                    handle.seek(int(offset))
                    yield key, _read_matrix(handle)
    else:
        raise IOError(f"'{rspecifier}': a read specifier names one table")
