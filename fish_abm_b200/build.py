"""Builds libfishabm.so (the sm_100a C-ABI library) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfishabm.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fishabm.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


LAST = {"decision": None, "nvcc": None}   # what build_lib did in this process (bench.py prints it)


def nvcc_version() -> str:
    try:
        out = subprocess.run([nvcc_path(), "--version"], capture_output=True, text=True).stdout
        return next((l.split("release")[-1].strip() for l in out.splitlines() if "release" in l), "unknown")
    except Exception:  # noqa: BLE001
        return "unavailable"


def build_lib(force: bool = False, verbose: bool = False, extra=(), out: str | None = None) -> str:
    """out: build a variant (extra flags, e.g. -DFISHABM_V3_MIN_CTAS=3) to another path, for A/B runs with FISHABM_LIB"""
    if out:
        cmd = [nvcc_path(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-ccbin", shutil.which("g++") or "g++",
               "-I", os.path.join(HERE, "..", "include"), *extra, *sources(), "-o", out, "-lcudart"]
        subprocess.run(cmd, check=True)
        return out
    if not force and not needs_build():
        LAST.update(decision="reused (libfishabm.so newer than every source)", nvcc=None)
        return LIB
    LAST.update(decision="rebuilt", nvcc=nvcc_version())
    cmd = [nvcc_path(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
           "-ccbin", shutil.which("g++") or "g++", "-I", os.path.join(HERE, "..", "include"), *extra,
           *sources(), "-o", LIB, "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfishabm.so")
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose="-v" in sys.argv))
