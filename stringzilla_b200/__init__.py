"""stringzilla_b200 — B200-native batched string similarity behind StringZilla's `szs_*` C ABI.

Host-side mirror of the reference's Python operator interface for this hot path
(/root/reference/python/stringzillas/similarities.c, device_scope.c): the same class names, argument names and
result shapes — `LevenshteinDistances`, `NeedlemanWunschScores`, `SmithWatermanScores`, `DeviceScope` — implemented as
a thin ctypes layer over `lib/libstringzillas_b200.so`. All arithmetic happens in the CUDA library; this module only
packs strings into Arrow-style tapes and allocates the result matrix. There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as ct
from typing import Iterable, Sequence as SequenceType

import numpy as np

from . import _abi
from ._abi import CallStats, SimilarityError  # noqa: F401  (re-exported)

__all__ = [
    "Communicator",
    "DeviceGroup",
    "DeviceScope",
    "LevenshteinDistances",
    "LevenshteinDistancesUTF8",
    "NeedlemanWunschScores",
    "SmithWatermanScores",
    "Tape",
    "SimilarityError",
    "capabilities",
]


class Tape:
    """Arrow-style string collection: one contiguous byte blob + N+1 64-bit offsets (reference stringzillas.h:88-92).

    Plays the role of `stringzilla.Strs` in the reference's Python API.
    """

    __slots__ = ("data", "offsets", "count", "_struct")

    def __init__(self, strings: "Iterable[bytes | str] | Tape | tuple[np.ndarray, np.ndarray]"):
        if isinstance(strings, Tape):
            self.data, self.offsets = strings.data, strings.offsets
        elif _is_arrow_array(strings):
            # Apache Arrow string / binary arrays ARE tapes (reference test/similarities.py:975-1005): the value buffer
            # is used in place, the offsets (sliced arrays keep a non-zero first offset) are widened to 64 bits.
            self.data, self.offsets = _arrow_buffers(strings)
        elif isinstance(strings, tuple) and len(strings) == 2 and isinstance(strings[0], np.ndarray):
            self.data = np.ascontiguousarray(strings[0], dtype=np.uint8)
            self.offsets = np.ascontiguousarray(strings[1], dtype=np.uint64)
        else:
            encoded = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
            lengths = np.fromiter((len(e) for e in encoded), dtype=np.uint64, count=len(encoded))
            self.offsets = np.zeros(len(encoded) + 1, dtype=np.uint64)
            np.cumsum(lengths, out=self.offsets[1:])
            self.data = np.frombuffer(b"".join(encoded), dtype=np.uint8).copy()
        if self.data.size == 0:
            self.data = np.zeros(1, dtype=np.uint8)  # keep a valid pointer for empty blobs
        self.count = int(self.offsets.size - 1)
        self._struct = _abi.SequenceU64Tape(self.data.ctypes.data, self.offsets.ctypes.data, self.count)

    def __len__(self) -> int:
        return self.count

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets).astype(np.int64)

    def __getitem__(self, index: int) -> bytes:
        return self.data[int(self.offsets[index]) : int(self.offsets[index + 1])].tobytes()

    def pointer(self):
        return ct.byref(self._struct)


def _is_arrow_array(value) -> bool:
    module = type(value).__module__ or ""
    return module.startswith("pyarrow") and hasattr(value, "buffers") and hasattr(value, "type")


def _arrow_buffers(array):
    import pyarrow as pa

    if isinstance(array, pa.ChunkedArray):
        array = array.combine_chunks()
    if array.null_count:
        raise ValueError("Arrow arrays with nulls are not supported: tapes hold non-NULL strings (stringzillas.h:75)")
    wide = pa.types.is_large_string(array.type) or pa.types.is_large_binary(array.type)
    if not (wide or pa.types.is_string(array.type) or pa.types.is_binary(array.type)):
        raise TypeError(f"expected a string or binary Arrow array, got {array.type}")
    _, offsets_buffer, data_buffer = array.buffers()
    offsets = np.frombuffer(offsets_buffer, dtype=np.int64 if wide else np.int32)
    offsets = offsets[array.offset : array.offset + len(array) + 1].astype(np.uint64)
    data = np.frombuffer(data_buffer, dtype=np.uint8) if data_buffer is not None else np.zeros(1, dtype=np.uint8)
    return data, offsets


_libraries: dict = {}


def _library(path=None) -> ct.CDLL:
    key = str(path) if path is not None else None
    if key not in _libraries:
        _libraries[key] = _abi.bind(path)
    return _libraries[key]


def capabilities(library_path=None) -> int:
    """`szs_capabilities()` of the bound library."""
    return int(_library(library_path).szs_capabilities())


class DeviceScope:
    """Mirror of `stringzillas.DeviceScope(cpu_cores=..., gpu_device=...)` (python/stringzillas/device_scope.c:31)."""

    def __init__(self, cpu_cores: int | None = None, gpu_device: int | None = None, library_path=None):
        self._lib = _library(library_path)
        self._handle = ct.c_void_p()
        message = ct.c_char_p()
        if cpu_cores is not None and gpu_device is not None:
            raise ValueError("cpu_cores and gpu_device are mutually exclusive")
        if gpu_device is not None:
            status = self._lib.szs_device_scope_init_gpu_device(int(gpu_device), ct.byref(self._handle), ct.byref(message))
        elif cpu_cores is not None:
            status = self._lib.szs_device_scope_init_cpu_cores(int(cpu_cores), ct.byref(self._handle), ct.byref(message))
        else:
            status = self._lib.szs_device_scope_init_default(ct.byref(self._handle), ct.byref(message))
        _abi.check(status, message)

    @property
    def handle(self) -> ct.c_void_p:
        return self._handle

    @property
    def stream(self) -> int:
        """The `cudaStream_t` (as an integer) every call on this GPU scope enqueues its work on."""
        value, message = ct.c_void_p(), ct.c_char_p()
        _abi.check(self._lib.szs_b200_device_scope_stream(self._handle, ct.byref(value), ct.byref(message)), message)
        return int(value.value or 0)

    @property
    def cpu_cores(self) -> int:
        value, message = ct.c_uint64(), ct.c_char_p()
        _abi.check(self._lib.szs_device_scope_get_cpu_cores(self._handle, ct.byref(value), ct.byref(message)), message)
        return int(value.value)

    @property
    def gpu_device(self) -> int:
        value, message = ct.c_uint64(), ct.c_char_p()
        _abi.check(self._lib.szs_device_scope_get_gpu_device(self._handle, ct.byref(value), ct.byref(message)), message)
        return int(value.value)

    @property
    def capabilities(self) -> int:
        value, message = ct.c_uint32(), ct.c_char_p()
        _abi.check(self._lib.szs_device_scope_get_capabilities(self._handle, ct.byref(value), ct.byref(message)), message)
        return int(value.value)

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            self._lib.szs_device_scope_free(handle)


class DeviceGroup:
    """Several GPUs of one box driven from THIS process (`szs_b200_device_group_*`): one scope and one parked host
    thread per GPU. `gpu_devices=None` takes every visible device. A B200 addition — the reference has no multi-GPU."""

    def __init__(self, gpu_devices: "SequenceType[int] | None" = None, library_path=None):
        self._lib = _library(library_path)
        self._handle = ct.c_void_p()
        message = ct.c_char_p()
        if gpu_devices is None:
            status = self._lib.szs_b200_device_group_init(None, 0, ct.byref(self._handle), ct.byref(message))
        else:
            ids = (ct.c_uint64 * len(gpu_devices))(*[int(d) for d in gpu_devices])
            status = self._lib.szs_b200_device_group_init(ids, len(gpu_devices), ct.byref(self._handle), ct.byref(message))
        _abi.check(status, message)

    @property
    def handle(self) -> ct.c_void_p:
        return self._handle

    def __len__(self) -> int:
        count = ct.c_uint64()
        self._lib.szs_b200_device_group_size(self._handle, ct.byref(count))
        return int(count.value)

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            self._lib.szs_b200_device_group_free(handle)


def _prefer_bundled_nccl() -> None:
    """The library resolves NCCL at run time (the copy already in the process, `SZS_B200_NCCL_LIBRARY`, or the system's
    libnccl.so.2). A process that will also import torch must not pull in an OLDER system NCCL first — the two share a
    soname and the first one loaded wins — so point the library at the copy torch bundles when there is one."""
    import importlib.util
    import os
    from pathlib import Path

    if "SZS_B200_NCCL_LIBRARY" in os.environ:
        return
    spec = importlib.util.find_spec("nvidia")
    for root in (spec.submodule_search_locations if spec is not None else []):
        candidate = Path(root) / "nccl" / "lib" / "libnccl.so.2"
        if candidate.exists():
            os.environ["SZS_B200_NCCL_LIBRARY"] = str(candidate)
            return


class Communicator:
    """One rank of a one-process-per-GPU job (`szs_b200_comm_*`): wraps an NCCL communicator created from C. Rank 0
    draws `Communicator.unique_id()` and passes the 128 bytes to the other ranks out of band (e.g. a
    `torch.distributed` broadcast or a file)."""

    def __init__(self, unique_id: bytes, world_size: int, rank: int, gpu_device: int, library_path=None):
        if len(unique_id) != _abi.COMM_ID_BYTES:
            raise ValueError(f"unique_id must be {_abi.COMM_ID_BYTES} bytes")
        _prefer_bundled_nccl()
        self._lib = _library(library_path)
        self._handle = ct.c_void_p()
        self.world_size, self.rank = int(world_size), int(rank)
        message = ct.c_char_p()
        buffer = ct.create_string_buffer(bytes(unique_id), _abi.COMM_ID_BYTES)
        _abi.check(self._lib.szs_b200_comm_init(buffer, self.world_size, self.rank, int(gpu_device), ct.byref(self._handle),
                                                ct.byref(message)), message)  # fmt: skip

    @staticmethod
    def unique_id(library_path=None) -> bytes:
        _prefer_bundled_nccl()
        buffer, message = ct.create_string_buffer(_abi.COMM_ID_BYTES), ct.c_char_p()
        _abi.check(_library(library_path).szs_b200_comm_unique_id(buffer, ct.byref(message)), message)
        return buffer.raw

    @property
    def handle(self) -> ct.c_void_p:
        return self._handle

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            self._lib.szs_b200_comm_free(handle)


def _resolve_capabilities(value, lib) -> int:
    if value is None:
        return int(lib.szs_capabilities())
    if isinstance(value, DeviceScope):
        return value.capabilities
    return int(value)


class _Engine:
    _family = ""
    _family_index = -1  # `family` argument of szs_b200_scores_u64tape_rows
    _result_dtype = np.int64

    def __init__(self, library_path=None):
        self._library_path = library_path
        self._lib = _library(library_path)
        self._handle = ct.c_void_p()
        self._default_scope = None

    def _call(self, queries, candidates, device, out, rows=None) -> np.ndarray:
        queries = queries if isinstance(queries, Tape) else Tape(queries)
        if candidates is not None and not isinstance(candidates, Tape):
            candidates = Tape(candidates)
        n_rows, n_cols = len(queries), len(candidates) if candidates is not None else len(queries)
        if out is None:
            out = np.zeros((n_rows, n_cols), dtype=self._result_dtype)
        else:
            if out.shape != (n_rows, n_cols) or out.dtype != self._result_dtype or out.strides[1] != out.itemsize:
                raise ValueError("`out` must be a (queries x candidates) matrix of the engine's dtype with unit column stride")
        stride = out.strides[0] // out.itemsize if n_rows > 1 else max(n_cols, 1)
        if device is None:
            if self._default_scope is None:
                self._default_scope = DeviceScope(library_path=self._library_path)
            device = self._default_scope
        message = ct.c_char_p()
        candidates_pointer = candidates.pointer() if candidates is not None else None
        if rows is None:
            function = getattr(self._lib, f"szs_{self._family}_u64tape")
            status = function(self._handle, device.handle, queries.pointer(), candidates_pointer,
                              out.ctypes.data_as(ct.c_void_p), stride, ct.byref(message))  # fmt: skip
        else:
            status = self._lib.szs_b200_scores_u64tape_rows(
                self._family_index, self._handle, device.handle,
                queries.pointer(), candidates_pointer, int(rows[0]), int(rows[1]),
                out.ctypes.data_as(ct.c_void_p), stride, ct.byref(message))  # fmt: skip
        _abi.check(status, message)
        return out

    def __call__(self, queries, candidates=None, device: DeviceScope | None = None, out: np.ndarray | None = None):
        """Scores every query against every candidate; `candidates=None` requests symmetric self-similarity."""
        return self._call(queries, candidates, device, out)

    def score_rows(self, queries, candidates, row_begin: int, row_end: int, device=None, out=None) -> np.ndarray:
        """Shard primitive: fills only rows [row_begin, row_end) of the full-shape matrix `out`."""
        return self._call(queries, candidates, device, out, rows=(row_begin, row_end))

    def sharded(self, queries, candidates=None, group: "DeviceGroup | None" = None, out: np.ndarray | None = None):
        """The same cross product row-sharded over the GPUs of `group` from this process
        (`szs_b200_scores_u64tape_sharded`): no collective, every GPU reads the inputs and writes its rows itself."""
        queries = queries if isinstance(queries, Tape) else Tape(queries)
        if candidates is not None and not isinstance(candidates, Tape):
            candidates = Tape(candidates)
        n_rows, n_cols = len(queries), len(candidates) if candidates is not None else len(queries)
        if out is None:
            out = np.zeros((n_rows, n_cols), dtype=self._result_dtype)
        elif out.shape != (n_rows, n_cols) or out.dtype != self._result_dtype or out.strides[1] != out.itemsize:
            raise ValueError("`out` must be a (queries x candidates) matrix of the engine's dtype with unit column stride")
        stride = out.strides[0] // out.itemsize if n_rows > 1 else max(n_cols, 1)
        message = ct.c_char_p()
        status = self._lib.szs_b200_scores_u64tape_sharded(
            self._family_index, self._handle, group.handle, queries.pointer(),
            candidates.pointer() if candidates is not None else None, out.ctypes.data_as(ct.c_void_p), stride,
            ct.byref(message))  # fmt: skip
        _abi.check(status, message)
        return out

    def communicate(self, communicator: "Communicator", device: DeviceScope, queries=None, candidates=None,
                    out: np.ndarray | None = None):
        """Collective over a `Communicator` (`szs_b200_comm_scores_u64tape`): rank 0 passes the strings and receives the
        whole matrix; the other ranks pass nothing and return None."""
        message = ct.c_char_p()
        if communicator.rank == 0:
            queries = queries if isinstance(queries, Tape) else Tape(queries)
            if candidates is not None and not isinstance(candidates, Tape):
                candidates = Tape(candidates)
            n_rows, n_cols = len(queries), len(candidates) if candidates is not None else len(queries)
            if out is None:
                out = np.zeros((n_rows, n_cols), dtype=self._result_dtype)
            stride = out.strides[0] // out.itemsize if n_rows > 1 else max(n_cols, 1)
            status = self._lib.szs_b200_comm_scores_u64tape(
                self._family_index, self._handle, device.handle, communicator.handle, queries.pointer(),
                candidates.pointer() if candidates is not None else None, out.ctypes.data_as(ct.c_void_p), stride,
                ct.byref(message))  # fmt: skip
        else:
            out = None
            status = self._lib.szs_b200_comm_scores_u64tape(self._family_index, self._handle, device.handle,
                                                            communicator.handle, None, None, None, 0, ct.byref(message))
        _abi.check(status, message)
        return out

    @property
    def last_call_stats(self) -> CallStats:
        stats = CallStats()
        self._lib.szs_b200_last_call_stats(self._handle, ct.byref(stats))
        return stats

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            getattr(self._lib, f"szs_{self._family}_free")(handle)


class LevenshteinDistances(_Engine):
    """Mirror of `stringzillas.LevenshteinDistances(match=0, mismatch=1, open=1, extend=1, capabilities=...)`
    (python/stringzillas/similarities.c:63-91). Returns a `uint64` matrix.

    `bit_parallel=False` (a B200 addition) keeps unit-cost engines on the int16 DPX wavefront tier."""

    _family = "levenshtein_distances"
    _family_index = 0
    _result_dtype = np.uint64

    def __init__(self, match: int = 0, mismatch: int = 1, open: int = 1, extend: int = 1, capabilities=None,
                 library_path=None, bit_parallel: bool = True):  # fmt: skip
        super().__init__(library_path)
        message = ct.c_char_p()
        status = self._lib.szs_levenshtein_distances_init(
            match, mismatch, open, extend, None, _resolve_capabilities(capabilities, self._lib),
            ct.byref(self._handle), ct.byref(message))  # fmt: skip
        _abi.check(status, message)
        if not bit_parallel:
            self._lib.szs_b200_set_bit_parallel(self._handle, 0)


class LevenshteinDistancesUTF8(_Engine):
    """Mirror of `stringzillas.LevenshteinDistancesUTF8(match=0, mismatch=1, open=1, extend=1, capabilities=...)`
    (python/stringzillas/similarities.c:519-547): edit distance over UTF-8 runes instead of bytes."""

    _family = "levenshtein_distances_utf8"
    _family_index = 0  # a Levenshtein engine as far as the shard primitives are concerned
    _result_dtype = np.uint64

    def __init__(self, match: int = 0, mismatch: int = 1, open: int = 1, extend: int = 1, capabilities=None,
                 library_path=None):  # fmt: skip
        super().__init__(library_path)
        message = ct.c_char_p()
        status = self._lib.szs_levenshtein_distances_utf8_init(
            match, mismatch, open, extend, None, _resolve_capabilities(capabilities, self._lib),
            ct.byref(self._handle), ct.byref(message))  # fmt: skip
        _abi.check(status, message)


class _Weighted(_Engine):
    def __init__(self, byte_to_class, class_substitution_costs, open: int = -1, extend: int = -1, capabilities=None,
                 library_path=None):  # fmt: skip
        super().__init__(library_path)
        table = np.asarray(class_substitution_costs)
        if table.shape != (32, 32):
            raise ValueError("class_substitution_costs must be a 32x32 array")
        if table.dtype != np.int8:
            raise TypeError("class_substitution_costs must have int8 dtype")
        classes = np.asarray(byte_to_class)
        if classes.shape != (256,) or classes.dtype != np.uint8:
            raise TypeError("byte_to_class must be 256 uint8 values")
        self._table = np.ascontiguousarray(table)
        self._classes = np.ascontiguousarray(classes)
        message = ct.c_char_p()
        status = getattr(self._lib, f"szs_{self._family}_init")(
            self._classes.ctypes.data, self._table.ctypes.data, open, extend, None,
            _resolve_capabilities(capabilities, self._lib), ct.byref(self._handle), ct.byref(message))  # fmt: skip
        _abi.check(status, message)


class NeedlemanWunschScores(_Weighted):
    """Mirror of `stringzillas.NeedlemanWunschScores(byte_to_class, class_substitution_costs, open, extend, ...)`
    (python/stringzillas/similarities.c:912-941). Returns an `int64` matrix of global alignment scores."""

    _family = "needleman_wunsch_scores"
    _family_index = 1


class SmithWatermanScores(_Weighted):
    """Mirror of `stringzillas.SmithWatermanScores(...)` (python/stringzillas/similarities.c:1340-1369)."""

    _family = "smith_waterman_scores"
    _family_index = 2
