"""ctypes view of the `szs_*` C ABI (include/stringzillas_b200.h == reference include/stringzillas/stringzillas.h).

The same prototypes bind ANY library exporting that ABI. The product path always binds the in-tree
`stringzilla_b200/lib/libstringzillas_b200.so` and fails loudly when it is missing; tests additionally bind the
unmodified reference CPU build (`oracle/_ref/libstringzillas_ref.so`) through `bind(path)` to prove the drop-in claim.
"""

from __future__ import annotations

import ctypes as ct
import os
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
DEFAULT_LIBRARY = PACKAGE_DIR / "lib" / "libstringzillas_b200.so"

# reference include/stringzilla/types.h:810-833
STATUS_NAMES = {
    0: "success",
    -10: "bad_alloc",
    -12: "invalid_utf8",
    -13: "contains_duplicates",
    -14: "overflow_risk",
    -15: "unexpected_dimensions",
    -16: "missing_gpu",
    -17: "device_code_mismatch",
    -18: "device_memory_mismatch",
    -19: "authentication_failed",
    -1: "unknown",
}

# reference include/stringzilla/types.h:839-888
CAP_SERIAL = 1
CAP_PARALLEL = 1 << 2
CAP_HASWELL = 1 << 5
CAP_SKYLAKE = 1 << 6
CAP_ICELAKE = 1 << 7
CAP_CUDA = 1 << 21
CAP_KEPLER = 1 << 22
CAP_HOPPER = 1 << 23
CAPS_CUDA = CAP_CUDA | CAP_KEPLER | CAP_HOPPER
CAP_ANY = 0x7FFFFFFF


class SequenceU32Tape(ct.Structure):
    """reference stringzillas.h:77-81"""

    _fields_ = [("data", ct.c_void_p), ("offsets", ct.c_void_p), ("count", ct.c_uint64)]


class SequenceU64Tape(ct.Structure):
    """reference stringzillas.h:88-92"""

    _fields_ = [("data", ct.c_void_p), ("offsets", ct.c_void_p), ("count", ct.c_uint64)]


SEQUENCE_START = ct.CFUNCTYPE(ct.c_void_p, ct.c_void_p, ct.c_uint64)
SEQUENCE_LENGTH = ct.CFUNCTYPE(ct.c_uint64, ct.c_void_p, ct.c_uint64)


class Sequence(ct.Structure):
    """reference types.h:1393-1398"""

    _fields_ = [
        ("handle", ct.c_void_p),
        ("count", ct.c_uint64),
        ("get_start", SEQUENCE_START),
        ("get_length", SEQUENCE_LENGTH),
    ]


class CallStats(ct.Structure):
    """include/stringzillas_b200.h: szs_b200_call_stats_t"""

    _fields_ = [
        ("kernel_milliseconds", ct.c_double),
        ("total_milliseconds", ct.c_double),
        ("cells", ct.c_uint64),
        ("kernel_launches", ct.c_uint64),
        ("h2d_bytes", ct.c_uint64),
        ("d2h_bytes", ct.c_uint64),
        ("pairs_u16", ct.c_uint64),
        ("pairs_s32", ct.c_uint64),
        ("pairs_bit_parallel", ct.c_uint64),
    ]


class SimilarityError(RuntimeError):
    def __init__(self, status: int, message: str | None):
        self.status = status
        self.status_name = STATUS_NAMES.get(status, str(status))
        super().__init__(f"sz_status_t {status} ({self.status_name}): {message or ''}")


# Every symbol include/stringzillas_b200.h declares that the reference header declares too.
REFERENCE_SYMBOLS = [
    "szs_version_major", "szs_version_minor", "szs_version_patch",
    "szs_capabilities", "szs_capabilities_comptime", "szs_capabilities_runtime",
    "sz_memory_allocator_init_unified", "szs_unified_alloc", "szs_unified_free",
    "szs_device_scope_init_default", "szs_device_scope_init_cpu_cores", "szs_device_scope_init_gpu_device",
    "szs_device_scope_get_cpu_cores", "szs_device_scope_get_gpu_device", "szs_device_scope_get_capabilities",
    "szs_device_scope_free",
    "szs_levenshtein_distances_init", "szs_levenshtein_distances", "szs_levenshtein_distances_u32tape",
    "szs_levenshtein_distances_u64tape", "szs_levenshtein_distances_free",
    "szs_levenshtein_distances_utf8_init", "szs_levenshtein_distances_utf8",
    "szs_levenshtein_distances_utf8_u32tape", "szs_levenshtein_distances_utf8_u64tape",
    "szs_levenshtein_distances_utf8_free",
    "szs_needleman_wunsch_scores_init", "szs_needleman_wunsch_scores", "szs_needleman_wunsch_scores_u32tape",
    "szs_needleman_wunsch_scores_u64tape", "szs_needleman_wunsch_scores_free",
    "szs_smith_waterman_scores_init", "szs_smith_waterman_scores", "szs_smith_waterman_scores_u32tape",
    "szs_smith_waterman_scores_u64tape", "szs_smith_waterman_scores_free",
    "szs_fingerprints_init", "szs_fingerprints_sequence", "szs_fingerprints_u32tape", "szs_fingerprints_u64tape",
    "szs_fingerprints_free",
]  # fmt: skip
B200_SYMBOLS = ["szs_b200_last_call_stats", "szs_b200_scores_u64tape_rows", "szs_b200_set_bit_parallel",
                "szs_b200_debug_schedule", "szs_b200_device_scope_stream", "szs_b200_partition_rows", "szs_b200_device_group_init",
                "szs_b200_device_group_size", "szs_b200_device_group_free", "szs_b200_scores_u64tape_sharded",
                "szs_b200_comm_unique_id", "szs_b200_comm_init", "szs_b200_comm_free", "szs_b200_comm_scores_u64tape"]
COMM_ID_BYTES = 128


def bind(path: os.PathLike | str | None = None) -> ct.CDLL:
    """Loads a library exporting the `szs_*` ABI and attaches prototypes. No fallback: a missing file raises."""
    if path is None and os.environ.get("SZS_B200_LIBRARY"):  # kernel experiments: a variant build of the same library
        path = os.environ["SZS_B200_LIBRARY"]
    path = Path(path) if path is not None else DEFAULT_LIBRARY
    if not path.exists():
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback in stringzilla_b200)"
        )
    lib = ct.CDLL(str(path))
    err = ct.POINTER(ct.c_char_p)
    vp, u64, i8 = ct.c_void_p, ct.c_uint64, ct.c_int8

    for name in ("szs_version_major", "szs_version_minor", "szs_version_patch"):
        getattr(lib, name).restype = ct.c_int
    for name in ("szs_capabilities", "szs_capabilities_comptime", "szs_capabilities_runtime"):
        getattr(lib, name).restype = ct.c_uint32

    lib.szs_unified_alloc.restype = vp
    lib.szs_unified_alloc.argtypes = [u64]
    lib.szs_unified_free.restype = None
    lib.szs_unified_free.argtypes = [vp, u64]

    lib.szs_device_scope_init_default.argtypes = [ct.POINTER(vp), err]
    lib.szs_device_scope_init_cpu_cores.argtypes = [u64, ct.POINTER(vp), err]
    lib.szs_device_scope_init_gpu_device.argtypes = [u64, ct.POINTER(vp), err]
    lib.szs_device_scope_get_cpu_cores.argtypes = [vp, ct.POINTER(u64), err]
    lib.szs_device_scope_get_gpu_device.argtypes = [vp, ct.POINTER(u64), err]
    lib.szs_device_scope_get_capabilities.argtypes = [vp, ct.POINTER(ct.c_uint32), err]
    lib.szs_device_scope_free.argtypes = [vp]
    lib.szs_device_scope_free.restype = None

    lib.szs_levenshtein_distances_init.argtypes = [i8, i8, i8, i8, vp, ct.c_uint32, ct.POINTER(vp), err]
    lib.szs_levenshtein_distances_utf8_init.argtypes = [i8, i8, i8, i8, vp, ct.c_uint32, ct.POINTER(vp), err]
    for family in ("szs_needleman_wunsch_scores", "szs_smith_waterman_scores"):
        getattr(lib, family + "_init").argtypes = [vp, vp, i8, i8, vp, ct.c_uint32, ct.POINTER(vp), err]
    for family in ("szs_levenshtein_distances", "szs_levenshtein_distances_utf8", "szs_needleman_wunsch_scores",
                   "szs_smith_waterman_scores"):
        for suffix in ("", "_u32tape", "_u64tape"):
            getattr(lib, family + suffix).argtypes = [vp, vp, vp, vp, vp, u64, err]
            getattr(lib, family + suffix).restype = ct.c_int
        getattr(lib, family + "_init").restype = ct.c_int
        getattr(lib, family + "_free").argtypes = [vp]
        getattr(lib, family + "_free").restype = None

    if hasattr(lib, "szs_b200_last_call_stats"):
        lib.szs_b200_last_call_stats.argtypes = [vp, ct.POINTER(CallStats)]
        lib.szs_b200_last_call_stats.restype = ct.c_int
        lib.szs_b200_scores_u64tape_rows.argtypes = [ct.c_int, vp, vp, vp, vp, u64, u64, vp, u64, err]
        lib.szs_b200_scores_u64tape_rows.restype = ct.c_int
        lib.szs_b200_set_bit_parallel.argtypes = [vp, ct.c_int]
        lib.szs_b200_set_bit_parallel.restype = ct.c_int
        lib.szs_b200_debug_schedule.argtypes = [vp, vp, u64, vp, u64, ct.c_int, ct.c_int, vp, vp, vp, vp, u64, ct.POINTER(u64)]
        lib.szs_b200_debug_schedule.restype = ct.c_int
        lib.szs_b200_device_scope_stream.argtypes = [vp, ct.POINTER(vp), err]
        lib.szs_b200_device_scope_stream.restype = ct.c_int
        lib.szs_b200_partition_rows.argtypes = [vp, u64, u64, ct.c_int, vp]
        lib.szs_b200_partition_rows.restype = ct.c_int
        lib.szs_b200_device_group_init.argtypes = [vp, u64, ct.POINTER(vp), err]
        lib.szs_b200_device_group_init.restype = ct.c_int
        lib.szs_b200_device_group_size.argtypes = [vp, ct.POINTER(u64)]
        lib.szs_b200_device_group_size.restype = ct.c_int
        lib.szs_b200_device_group_free.argtypes = [vp]
        lib.szs_b200_device_group_free.restype = None
        lib.szs_b200_scores_u64tape_sharded.argtypes = [ct.c_int, vp, vp, vp, vp, vp, u64, err]
        lib.szs_b200_scores_u64tape_sharded.restype = ct.c_int
        lib.szs_b200_comm_unique_id.argtypes = [vp, err]
        lib.szs_b200_comm_unique_id.restype = ct.c_int
        lib.szs_b200_comm_init.argtypes = [vp, ct.c_int, ct.c_int, u64, ct.POINTER(vp), err]
        lib.szs_b200_comm_init.restype = ct.c_int
        lib.szs_b200_comm_free.argtypes = [vp]
        lib.szs_b200_comm_free.restype = None
        lib.szs_b200_comm_scores_u64tape.argtypes = [ct.c_int, vp, vp, vp, vp, vp, vp, u64, err]
        lib.szs_b200_comm_scores_u64tape.restype = ct.c_int
    return lib


def check(status: int, message: ct.c_char_p) -> None:
    if status != 0:
        text = message.value.decode(errors="replace") if message and message.value else None
        raise SimilarityError(status, text)
