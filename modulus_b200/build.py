"""Build recipe for libmgn_b200.so (the C-ABI CUDA library) -- sm_100a only.

`python -m modulus_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU.  The shared object is written in-tree (modulus_b200/lib/) so it travels with the repo
snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libmgn_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


# extra nvcc flags for A/B builds of kernel variants (e.g. MGN_NVCC_EXTRA="-DMGN_WAIT_HINT=20000"); part of the digest
NVCC_FLAGS += os.environ.get("MGN_NVCC_EXTRA", "").split()


def sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    """sha256 over every source the library is made of and the compiler flags.  It is compiled INTO the library
    (mgn_build_digest()); `_lib.load()` recomputes it from the sources next to it and refuses a library that was built
    from anything else, so a stale binary can never be bound to prototypes parsed from a newer header."""
    h = hashlib.sha256()
    inc = ROOT.parent / "include"
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [inc / "mgn_b200.h", inc / "mgn_b200_debug.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    LIBDIR.mkdir(exist_ok=True)
    dig = _digest()
    if not force and LIB.exists() and embedded_digest() == dig:
        return LIB
    if not os.path.exists(nvcc):
        if LIB.exists():  # GPU box without a toolchain mismatch: use the prebuilt library
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libmgn_b200.so present")
    objs = []
    procs = []
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if src.name == "mgn_misc.cu":
            cmd.insert(1, f'-DMGN_BUILD_DIGEST="MGNDIGEST:{dig}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src.name} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs)]
    subprocess.run(cmd, check=True)
    for stale in objdir.glob("*.o"):  # objects of sources that no longer exist
        if stale not in objs:
            stale.unlink()
    return LIB


def embedded_digest() -> str:
    """digest string compiled into the built library ('' when absent).  Read from the file's bytes (the string is stored
    behind the marker "MGNDIGEST:"), not through dlopen: a handle opened here would shadow the rebuilt library."""
    import re

    try:
        m = re.search(rb"MGNDIGEST:([0-9a-f]{64})", LIB.read_bytes())
    except OSError:
        return ""
    return m.group(1).decode() if m else ""


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
