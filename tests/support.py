"""Test-only helpers: the CPU oracle (oracle/liboracle.so), the reference CPU build (oracle/_ref), corpora."""

from __future__ import annotations

import ctypes as ct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SOURCE = ROOT / "oracle" / "sz_oracle.c"
ORACLE_LIBRARY = ROOT / "oracle" / "liboracle.so"
REFERENCE_LIBRARY = ROOT / "oracle" / "_ref" / "libstringzillas_ref.so"
GOLDEN_DIR = ROOT / "tests" / "golden"

DNA = b"ACGT"
PROTEIN = b"ARNDCQEGHILKMFPSTWYV"


def build_oracle() -> Path:
    if not ORACLE_LIBRARY.exists() or ORACLE_LIBRARY.stat().st_mtime < ORACLE_SOURCE.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-fvisibility=hidden",
             "-o", str(ORACLE_LIBRARY), str(ORACLE_SOURCE)], check=True)  # fmt: skip
    return ORACLE_LIBRARY


_oracle = None


def oracle() -> ct.CDLL:
    global _oracle
    if _oracle is None:
        lib = ct.CDLL(str(build_oracle()))
        vp, u64, i8 = ct.c_void_p, ct.c_uint64, ct.c_int8
        lib.oracle_levenshtein.restype = ct.c_uint64
        lib.oracle_levenshtein.argtypes = [vp, u64, vp, u64, i8, i8, i8, i8]
        lib.oracle_weighted.restype = ct.c_int64
        lib.oracle_weighted.argtypes = [vp, u64, vp, u64, vp, vp, i8, i8, ct.c_int]
        lib.oracle_levenshtein_cross.argtypes = [vp, vp, u64, vp, vp, u64, i8, i8, i8, i8, vp, u64]
        lib.oracle_weighted_cross.argtypes = [vp, vp, u64, vp, vp, u64, vp, vp, i8, i8, ct.c_int, vp, u64]
        lib.oracle_levenshtein_utf8_cross.argtypes = [vp, vp, u64, vp, vp, u64, i8, i8, i8, i8, vp, u64]
        _oracle = lib
    return _oracle


def _tape(strings):
    from stringzilla_b200 import Tape

    return strings if isinstance(strings, Tape) else Tape(strings)


def oracle_levenshtein(queries, candidates=None, match=0, mismatch=1, open=1, extend=1) -> np.ndarray:
    q = _tape(queries)
    c = _tape(candidates) if candidates is not None else None
    cols = len(c) if c is not None else len(q)
    out = np.zeros((len(q), cols), dtype=np.uint64)
    oracle().oracle_levenshtein_cross(
        q.data.ctypes.data, q.offsets.ctypes.data, len(q),
        c.data.ctypes.data if c is not None else None, c.offsets.ctypes.data if c is not None else None, cols,
        match, mismatch, open, extend, out.ctypes.data, max(cols, 1))  # fmt: skip
    return out


def oracle_levenshtein_utf8(queries, candidates=None, match=0, mismatch=1, open=1, extend=1) -> np.ndarray:
    q = _tape(queries)
    c = _tape(candidates) if candidates is not None else None
    cols = len(c) if c is not None else len(q)
    out = np.zeros((len(q), cols), dtype=np.uint64)
    oracle().oracle_levenshtein_utf8_cross(
        q.data.ctypes.data, q.offsets.ctypes.data, len(q),
        c.data.ctypes.data if c is not None else None, c.offsets.ctypes.data if c is not None else None, cols,
        match, mismatch, open, extend, out.ctypes.data, max(cols, 1))  # fmt: skip
    return out


def oracle_weighted(queries, candidates, byte_to_class, costs, open, extend, local: bool) -> np.ndarray:
    q = _tape(queries)
    c = _tape(candidates) if candidates is not None else None
    cols = len(c) if c is not None else len(q)
    out = np.zeros((len(q), cols), dtype=np.int64)
    byte_to_class = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
    costs = np.ascontiguousarray(costs, dtype=np.int8)
    oracle().oracle_weighted_cross(
        q.data.ctypes.data, q.offsets.ctypes.data, len(q),
        c.data.ctypes.data if c is not None else None, c.offsets.ctypes.data if c is not None else None, cols,
        byte_to_class.ctypes.data, costs.ctypes.data, open, extend, int(local), out.ctypes.data, max(cols, 1))  # fmt: skip
    return out


def random_strings(rng: np.random.Generator, count: int, low: int, high: int, alphabet: bytes) -> list[bytes]:
    """`count` strings with lengths uniform on [low, high] over `alphabet`."""
    symbols = np.frombuffer(alphabet, dtype=np.uint8)
    lengths = rng.integers(low, high + 1, size=count)
    return [symbols[rng.integers(0, len(symbols), size=int(n))].tobytes() for n in lengths]


def mutated_copies(rng: np.random.Generator, strings: list[bytes], rate: float, alphabet: bytes) -> list[bytes]:
    """Related strings: substitutions, insertions and deletions at `rate` per character."""
    out = []
    for s in strings:
        buffer = bytearray()
        for ch in s:
            roll = rng.random()
            if roll < rate / 3:
                continue
            if roll < 2 * rate / 3:
                buffer.append(alphabet[int(rng.integers(0, len(alphabet)))])
            elif roll < rate:
                buffer.append(ch)
                buffer.append(alphabet[int(rng.integers(0, len(alphabet)))])
                continue
            else:
                buffer.append(ch)
        out.append(bytes(buffer))
    return out


def have_reference() -> bool:
    return REFERENCE_LIBRARY.exists()


class DeviceBuffer:
    """A copy of a numpy array in DEVICE memory: a torch CUDA tensor on a GPU box, memory of the emulated device
    (tests/emulator, `simt_device_alloc`) when the tests run against the emulated library."""

    def __init__(self, array: np.ndarray):
        from stringzilla_b200 import _abi

        array = np.ascontiguousarray(array)
        self.shape, self.dtype, self.nbytes = array.shape, array.dtype, max(array.nbytes, 1)
        self._lib = _abi.bind()
        if hasattr(self._lib, "simt_device_alloc"):
            self._lib.simt_device_alloc.restype, self._lib.simt_device_alloc.argtypes = ct.c_void_p, [ct.c_uint64]
            self._lib.simt_device_free.restype, self._lib.simt_device_free.argtypes = None, [ct.c_void_p]
            self._tensor, self.pointer = None, self._lib.simt_device_alloc(self.nbytes)
            ct.memmove(self.pointer, array.ctypes.data, array.nbytes)
        else:
            import torch

            self._tensor = torch.from_numpy(array.view(np.uint8).reshape(-1).copy()).cuda()
            self.pointer = self._tensor.data_ptr()

    def numpy(self) -> np.ndarray:
        if self._tensor is not None:
            return self._tensor.cpu().numpy().view(self.dtype).reshape(self.shape)
        return np.frombuffer(ct.string_at(self.pointer, int(np.prod(self.shape)) * self.dtype.itemsize), dtype=self.dtype).reshape(self.shape).copy()

    def __del__(self):
        if getattr(self, "_tensor", None) is None and getattr(self, "pointer", None):
            self._lib.simt_device_free(self.pointer)
            self.pointer = None

