"""Builds libruvgnn.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libruvgnn.so"
SOURCES = ["ffi.cu", "gather.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tc3.cu", "gemm_upd_fused.cu", "rowops.cu", "search.cu", "dist.cu", "store.cu", "json.cc", "packer.cc", "hnswio.cc"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive a snapshot)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    files = sorted(CSRC.glob("*")) + [PKG.parent / "include" / "ruvgnn.h"]
    for f in files:
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source to objects (in parallel) and link the shared library."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    stamp = PKG / "libruvgnn.so.stamp"
    want = source_hash()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return LIB
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "ruvgnn.h"]
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = objdir / (src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(CSRC / src)]
            log = open(objdir / (src + ".log"), "w")
            procs.append((src, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log))
    for src, p, log in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + (objdir / (src + ".log")).read_text()[-4000:])
        if verbose:
            print((objdir / (src + ".log")).read_text())
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs), "-ldl"]
        subprocess.run(cmd, check=True)
    stamp.write_text(want)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
