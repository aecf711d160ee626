"""Build the CUDA shared library in-tree (rminc_b200/librminc_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The library links only the CUDA runtime (static);
it has no torch or R dependency -- that is the drop-in C ABI of include/rminc_b200.h.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librminc_b200.so")

SOURCES = ["capi.cu", "lm_kernels.cu", "lm_cov.cu", "lm_packed.cu", "fdr.cu", "tfce.cu", "randomize.cu", "dataset.cu", "perm_gemm.cu", "design.cpp", "minc2_reader.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; rminc_b200 has no CPU fallback and cannot be built without CUDA")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return os.path.getmtime(os.path.join(HERE, "..", "include", "rminc_b200.h")) > t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    import re

    def deps_time(path, seen=None):
        """Newest mtime of `path` and of every header it includes (quoted includes, recursively)."""
        seen = seen if seen is not None else set()
        path = os.path.normpath(path)
        if path in seen or not os.path.exists(path):
            return 0.0
        seen.add(path)
        t = os.path.getmtime(path)
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
            t = max(t, deps_time(os.path.join(os.path.dirname(path), inc), seen))
        return t

    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(build_dir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(deps_time(path), os.path.getmtime(__file__)):
            continue
        cmd = [nvcc(), *NVCC_FLAGS, *os.environ.get("RMG_NVCC_EXTRA", "").split(), "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose and out:
            print(out, file=sys.stderr)
    link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lpthread", "-ldl", "-lz"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
