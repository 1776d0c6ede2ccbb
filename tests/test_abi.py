"""CPU suite, part 2: the C-ABI library builds, loads and exports every symbol include/stringzillas_b200.h declares;
host-only entry points behave like the reference's; scoring calls refuse loudly without a GPU (no CPU fallback)."""

import ctypes as ct
import re

import numpy as np
import pytest

import stringzilla_b200 as szb
from stringzilla_b200 import _abi
from tests import support

HEADER = support.ROOT / "include" / "stringzillas_b200.h"


def test_library_exports_every_declared_symbol(library_built):
    lib = _abi.bind()
    declared = set(re.findall(r"\b(szs?_[a-z0-9_]+)\s*\(", HEADER.read_text()))
    declared = {name for name in declared if not name.endswith("_t")}
    assert set(_abi.REFERENCE_SYMBOLS) | set(_abi.B200_SYMBOLS) == declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref not built")
def test_reference_exports_the_same_abi(library_built):
    """Drop-in check: every reference-declared symbol we export is exported by the reference build too."""
    reference = ct.CDLL(str(support.REFERENCE_LIBRARY))
    for name in _abi.REFERENCE_SYMBOLS:
        if "fingerprints" in name:
            continue  # the oracle build skips the fingerprint TUs (out of scope)
        assert hasattr(reference, name), name


def test_version_and_capabilities(library_built):
    lib = _abi.bind()
    assert (lib.szs_version_major(), lib.szs_version_minor(), lib.szs_version_patch()) == (5, 1, 2)
    assert lib.szs_capabilities_comptime() & _abi.CAPS_CUDA == _abi.CAPS_CUDA
    assert lib.szs_capabilities() & _abi.CAP_SERIAL


def test_device_scopes_host_side(library_built):
    default = szb.DeviceScope()
    assert default.cpu_cores == 1
    four = szb.DeviceScope(cpu_cores=4)
    assert four.cpu_cores == 4
    with pytest.raises(szb.SimilarityError):
        four.gpu_device
    folded = szb.DeviceScope(cpu_cores=1)  # folds onto the default scope, reference runtime.cuh:93-94
    assert folded.cpu_cores == 1


def test_engine_lifecycle_without_gpu(library_built):
    """Engines are plain host objects until the first call: init/free work on a CPU-only box."""
    from stringzilla_b200 import scoring

    engine = szb.LevenshteinDistances(0, 1, 1, 1)
    assert engine._handle
    del engine
    byte_to_class, costs = scoring.blosum62()
    for engine_type in (szb.NeedlemanWunschScores, szb.SmithWatermanScores):
        assert engine_type(byte_to_class, costs, open=-4, extend=-1)._handle
    with pytest.raises(ValueError):
        szb.NeedlemanWunschScores(byte_to_class, costs[:8], open=-4, extend=-1)
    # A class index outside the 32x32 table (e.g. the two tables passed in each other's place) is refused at construction.
    out_of_range = byte_to_class.copy()
    out_of_range[ord("Q")] = 32
    with pytest.raises(szb.SimilarityError) as refusal:
        szb.SmithWatermanScores(out_of_range, costs, open=-4, extend=-1)
    assert refusal.value.status_name == "unexpected_dimensions"


def test_no_cpu_fallback(library_built):
    """A CPU scope must be refused; and without a GPU every scoring call fails instead of computing on the host."""
    engine = szb.LevenshteinDistances()
    with pytest.raises(szb.SimilarityError) as failure:
        engine([b"abc"], [b"abd"], device=szb.DeviceScope(cpu_cores=4))
    assert failure.value.status_name == "device_code_mismatch"
    lib = _abi.bind()
    if not (lib.szs_capabilities_runtime() & _abi.CAP_CUDA):
        with pytest.raises(szb.SimilarityError) as failure:
            engine([b"abc"], [b"abd"])
        assert failure.value.status_name == "missing_gpu"


@pytest.mark.skipif(not (support.ROOT.parent / "reference" / "include").is_dir() and not __import__("pathlib").Path("/root/reference/include").is_dir(),
                    reason="the reference header tree exists only in the build container")
def test_reference_header_compiles_and_links_against_our_library(library_built, tmp_path):
    """The drop-in claim at source level: a C program written against the REFERENCE's stringzillas.h links against
    libstringzillas_b200.so unchanged; without a GPU the scoring call must refuse (no CPU fallback)."""
    import subprocess

    binary = tmp_path / "dropin"
    library_dir = support.ROOT / "stringzilla_b200" / "lib"
    subprocess.run(["gcc", "-DSZ_DYNAMIC_DISPATCH=1", "-I/root/reference/include", str(support.ROOT / "tests" / "c_dropin_probe.c"),
                    "-o", str(binary), f"-L{library_dir}", "-lstringzillas_b200", f"-Wl,-rpath,{library_dir}"], check=True)
    ran = subprocess.run([str(binary)], capture_output=True, text=True)
    assert ran.returncode == 0, ran.stdout + ran.stderr
    assert "version 5.1.2" in ran.stdout


def test_tape_from_arrow_arrays():
    """Arrow string arrays are tapes already: value buffer used in place, slices keep their non-zero first offset."""
    pa = pytest.importorskip("pyarrow")
    words = ["", "alpha", "βήτα", "gamma", "", "delta"]
    for array in (pa.array(words), pa.array(words, type=pa.large_string()), pa.array([w.encode() for w in words], type=pa.binary())):
        tape = szb.Tape(array)
        assert [tape[i] for i in range(len(tape))] == [w.encode() for w in words]
        window = szb.Tape(array.slice(2, 3))
        assert [window[i] for i in range(len(window))] == [w.encode() for w in words[2:5]]
        assert int(window.offsets[0]) == len("alpha".encode()) and window.data.ctypes.data == tape.data.ctypes.data
    with pytest.raises(ValueError):
        szb.Tape(pa.array(["a", None]))


def test_tape_packing():
    tape = szb.Tape([b"", b"abc", "é"])
    assert len(tape) == 3 and tape.offsets.tolist() == [0, 0, 3, 5]
    assert tape[1] == b"abc" and tape[2] == "é".encode()
    assert tape.lengths.tolist() == [0, 3, 2]
