"""Builds libequilib_b200.so in-tree with nvcc for sm_100a.

    python -m equilib_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/equilib_b200.h); it does not
link against torch.  Objects go to equilib_b200/csrc/build/, the .so next to this
file so that it travels with the repo snapshot.
"""

from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libequilib_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--threads", "2"]
# experiments: extra -D switches (knock-outs, buffer sizes); part of the source hash
NVCC_FLAGS += os.environ.get("EQB_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _dep_files():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
             if f.endswith((".cu", ".cuh", ".inl", ".h"))]
    files.append(os.path.join(os.path.dirname(PKG), "include", "equilib_b200.h"))
    return sorted(files)


def source_hash() -> str:
    """sha256 over every source / header the library is built from (names + contents) and
    the compiler flags: the identity of a build.  It is compiled into the library
    (eqb_source_hash) so that a shipped .so can be checked against the tree it sits in --
    modification times say nothing after a checkout, a copy or a snapshot."""
    h = hashlib.sha256()
    for f in _dep_files():
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def built_hash() -> str:
    """The source hash compiled into the existing library ('' if there is none / it predates
    the hash).  Read from the file, not by loading it."""
    if not os.path.exists(LIB):
        return ""
    with open(LIB, "rb") as fh:
        blob = fh.read()
    tag = b"EQB_SRC_HASH="
    i = blob.find(tag)
    return blob[i + len(tag):i + len(tag) + 16].decode("ascii", "replace") if i >= 0 else ""


def is_fresh() -> bool:
    return built_hash() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    digest = source_hash()
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        tag = [f"-DEQB_SRC_HASH={digest}"] if os.path.basename(src) == "capi.cu" else []
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *extra, *tag, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + ".tmp"
    cmd = [nvcc, *ARCH, "-shared", "-o", tmp, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
