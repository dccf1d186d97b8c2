"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, and
refuses loudly to run without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sabreur_b200 import build
    build.build()
    from sabreur_b200 import _ffi
    return _ffi.load()


def test_header_symbols_are_exported(lib):
    from sabreur_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "sabreur_gpu.h")).read()
    declared = set(re.findall(r"\b(sbr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_ffi.EXPORTS), declared ^ set(_ffi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sabreur_gpu.h but not exported"
    assert lib.sbr_abi_version() == 2


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sabreur_b200 import demux
    assert lib.sbr_device_count() < 0 or lib.sbr_device_count() == 0
    with pytest.raises(demux.SabreurGpuError):
        demux.Demux([b"ACGT"], 0)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sabreur_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b", src) and "no CPU fallback" not in src[:0]:
                    if re.search(r"(import|include|from)\s+.*oracle", src):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_synth_table_and_host_generator(lib):
    import numpy as np
    from sabreur_b200 import synth
    for w in synth.WORKLOADS.values():
        tab = synth.table(w)
        d = (tab[:, None, :] != tab[None, :, :]).sum(-1)
        np.fill_diagonal(d, 99)
        assert d.min() >= 2 * w.mismatch + 1  # unambiguous under the reference's first-match rule
        text = synth.reads_host(w, tab, 5, 64).tobytes()
        assert len(text) == 64 * w.rec_bytes
        from oracle import oracle as orc
        recs, fmt, consumed = orc.parse(text)
        assert len(recs) == 64 and consumed == len(text) and fmt == w.fmt
        assert not (recs["flags"] & 1).any()  # canonical records
    # position independence: a slice equals the same records of a longer run
    w = synth.CFG3
    tab = synth.table(w)
    a = synth.reads_host(w, tab, 0, 50)
    b = synth.reads_host(w, tab, 20, 10)
    assert (a[20 * w.rec_bytes:30 * w.rec_bytes] == b).all()
    assert w.rec_bytes == 321 and synth.CFG5.rec_bytes == 266


def test_plain_c_caller_compiles_and_links(lib, tmp_path):
    """include/sabreur_gpu.h is C99 and a plain C program links against the library (no C++/torch types in the ABI)"""
    import subprocess
    from sabreur_b200 import _ffi
    hdr = os.path.join(ROOT, "include", "sabreur_gpu.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c", hdr])
    exe = os.path.join(ROOT, "sabreur_b200", "bin", "abi_smoke")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "abi_smoke.c"),
                           "-L" + libdir, "-lsabreur_gpu", "-Wl,-rpath," + libdir])
    assert os.path.exists(exe)
