"""Build libmptg.so (CUDA kernels + C ABI) for sm_100a, in-tree.

    python -m mpt_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The resulting mpt_b200/_lib/libmptg.so is git-ignored but travels
to the GPU box with the repo snapshot.  Flags that matter for parity with the CPU oracle:
--fmad=false (no implicit fused multiply-add; explicit fma calls in the sources are kept) and
-Xcompiler -ffp-contract=off for the host side.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "_lib"
LIB = LIBDIR / "libmptg.so"
NVCC = os.environ.get("MPTG_NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "--fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-fopenmp",
    "-ccbin", "/usr/bin/g++",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-warn-spills",
] + os.environ.get("MPTG_NVCC_EXTRA", "").split()  # kernel experiments: extra -D flags


def sources():
    return sorted(CSRC.glob("*.cu"))


def fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + list((ROOT.parent / "include" / "mptg").glob("*.h"))):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / "libmptg.stamp"
    fp = fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    objs = []
    procs = []
    for src in sources():
        obj = LIBDIR / (src.stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0 or "warning" in out or "spill" in out.lower()):
            print(f"--- {src.name}\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmptg.so")
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-ccbin", "/usr/bin/g++", "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lgomp"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise RuntimeError("link of libmptg.so failed")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
