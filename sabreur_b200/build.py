"""Builds libsabreur_gpu.so (sm_100a only) in-tree with nvcc.  `python -m sabreur_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsabreur_gpu.so")
SOURCES = ["ctx.cu", "k_scan.cu", "k_scan_fast.cu", "k_match.cu", "k_partition.cu", "k_synth.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math", "-diag-suppress", "177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: list[str] | None = None) -> str:
    """variant: an experimental build with extra -D flags, kept apart in lib_<variant>/ (select it at run time with
    SABREUR_GPU_LIB=<path>); the product is the plain build in lib/"""
    global LIBDIR, LIB
    if variant:
        LIBDIR = os.path.join(HERE, "lib_" + variant)
        LIB = os.path.join(LIBDIR, "libsabreur_gpu.so")
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sabreur_gpu.h"))
    objs = []
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + list(defines or []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    var = ""
    if "--variant" in sys.argv:
        var = sys.argv[sys.argv.index("--variant") + 1]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
