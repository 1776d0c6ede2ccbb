#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: builds tests/emulator/_build/libstringzillas_b200_emulated.so — the UNCHANGED sources of
`stringzilla_b200/csrc/` (host driver, C ABI, every kernel) compiled with g++ against the SIMT emulator of this directory.

The sources are not copied: `_build/tree/` mirrors the repo layout with symbolic links, except that `ptx.cuh` (the one
header holding inline PTX and the launch syntax) is the emulator's own. Only tests load the result; nothing in
`stringzilla_b200/` knows it exists.

    python tests/emulator/build.py          # builds if a source is newer than the library; prints its path
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "stringzilla_b200" / "csrc"
BUILD = HERE / "_build"
LIBRARY = BUILD / "libstringzillas_b200_emulated.so"
FAMILIES = range(9)
FLAGS = ["-std=c++17", "-O1", "-g0", "-fPIC", "-fvisibility=hidden", "-w", "-x", "c++", f"-I{HERE / 'include'}"]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu*")) + sorted(CSRC.glob("*.inc")) + sorted(CSRC.glob("*.?pp")) + sorted((HERE / "include").iterdir()) + \
        [HERE / "simt.cpp", ROOT / "include" / "stringzillas_b200.h", Path(__file__)]


def link_tree() -> Path:
    """`_build/tree/stringzilla_b200/csrc/*` -> the product sources (but the emulator's ptx.cuh), `_build/tree/include`."""
    tree = BUILD / "tree"
    csrc = tree / "stringzilla_b200" / "csrc"
    csrc.mkdir(parents=True, exist_ok=True)
    (tree / "include").mkdir(exist_ok=True)
    wanted = {path.name: path for path in CSRC.iterdir() if path.suffix in (".cu", ".cuh", ".inc", ".cpp", ".hpp")}
    wanted["ptx.cuh"] = HERE / "include" / "ptx.cuh"
    wanted.pop("carry_chain.inc", None)  # inline PTX; the emulator's ptx.cuh states the same addition in C++
    for stale in csrc.iterdir():
        stale.unlink()
    for name, target in wanted.items():
        (csrc / name).symlink_to(target)
    header = tree / "include" / "stringzillas_b200.h"
    if header.is_symlink() or header.exists():
        header.unlink()
    header.symlink_to(ROOT / "include" / "stringzillas_b200.h")
    return csrc


def build(verbose: bool = False, sanitize: bool = False) -> Path:
    """`sanitize=True` builds `libstringzillas_b200_emulated_asan.so` with AddressSanitizer: exact-size device allocations,
    the unrequested part of shared memory poisoned (run the tests with the sanitizer's runtime preloaded, see
    tools/emulator_memcheck.sh)."""
    library = LIBRARY.with_name("libstringzillas_b200_emulated_asan.so") if sanitize else LIBRARY
    if library.exists() and library.stat().st_mtime >= max(path.stat().st_mtime for path in sources()):
        return library
    BUILD.mkdir(exist_ok=True)
    csrc = link_tree()
    objects = BUILD / ("obj_asan" if sanitize else "obj")
    flags = FLAGS + (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-g1", "-DSIMT_SANITIZE=1"] if sanitize else [])
    objects.mkdir(exist_ok=True)
    jobs = [(csrc / "engine.cu", objects / "engine.o", []), (csrc / "myers.cu", objects / "myers.o", []),
            (csrc / "host_scan.cpp", objects / "host_scan.o", ["-O3"]), (HERE / "simt.cpp", objects / "simt.o", [])]
    jobs += [(csrc / "family.cu", objects / f"family{family}.o", [f"-DSZB_FAMILY={family}"]) for family in FAMILIES]

    def compile_one(job):
        source, output, extra = job
        command = ["g++", *flags, *extra, "-c", str(source), "-o", str(output)]
        done = subprocess.run(command, capture_output=True, text=True)
        if done.returncode != 0:
            raise RuntimeError(f"{' '.join(command)}\n{done.stderr[-6000:]}")
        if verbose:
            print(f"compiled {output.name}", file=sys.stderr)
        return output

    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as pool:
        compiled = list(pool.map(compile_one, jobs))
    temporary = library.with_suffix(".so.tmp")
    subprocess.run(["g++", "-shared", *(["-fsanitize=address,undefined"] if sanitize else []), "-o", str(temporary), *map(str, compiled),
                    "-lpthread", "-ldl"], check=True)
    temporary.replace(library)
    return library


def build_nccl() -> Path:
    """`libnccl_emulated.so`: NCCL's entry points over host threads of one process (nccl_threads.cpp), found by the emulated
    library through SZS_B200_NCCL_LIBRARY — so that the multi-rank exchange runs on the CPU tier with ranks as threads."""
    library, source = BUILD / "libnccl_emulated.so", HERE / "nccl_threads.cpp"
    if library.exists() and library.stat().st_mtime >= max(source.stat().st_mtime, (HERE / "include" / "nccl.h").stat().st_mtime):
        return library
    BUILD.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-fvisibility=hidden", "-shared", f"-I{HERE / 'include'}", str(source),
                    "-o", str(library), "-lpthread"], check=True)
    return library


if __name__ == "__main__":
    build_nccl()
    print(build(verbose=True, sanitize="--sanitize" in sys.argv))
