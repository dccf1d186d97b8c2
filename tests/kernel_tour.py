#!/usr/bin/env python
"""A small tour of every kernel variant for compute-sanitizer (memcheck): HBM-resident batches through the two-stream
pipeline (plain and lean match kernel), streamed batches through the region-parallel and the look-back scan, a batch that
is handed over, FASTA and FASTQ.  Everything is checked against the oracle, so the plain run is a test of its own.

    python tests/kernel_tour.py && compute-sanitizer --tool memcheck python tests/kernel_tour.py

(Round 2: compute-sanitizer is closed on this GPU pool -- "runs under it have left GPUs needing a reset" -- so only the
plain run exists; tests/test_gpu_tour.py keeps it in the GPU suite.  Test infrastructure: the oracle is the checker.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from sabreur_b200 import _ffi, demux, synth  # noqa: E402


def resident(name, n, nb, flags, force_lean=False):
    if force_lean:
        os.environ["SBR_FORCE_LEAN"] = "1"
    else:
        os.environ.pop("SBR_FORCE_LEAN", None)
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec = w.rec_bytes
    dev = torch.empty(n * nb * rec + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 1234, n * nb, dev.data_ptr())
    host = dev[: n * nb * rec].cpu().numpy()
    st = torch.cuda.Stream()
    with demux.Demux(bcs, w.mismatch, fmt=w.fmt, batch_bytes=n * rec + 4096, n_slots=2, max_records=n + 16, flags=flags) as dm:
        for i in range(nb):
            dm.demux_device(i % 2, dev.data_ptr() + i * n * rec, n * rec, final=True, stream=st.cuda_stream)
            if i >= 1:
                b = dm.wait((i - 1) % 2)
                exp = orc.demux(host[(i - 1) * n * rec:i * n * rec], bcs, w.mismatch)
                assert np.array_equal(b.assign, exp["assign"]) and np.array_equal(b.order, exp["order"]), (name, flags, i)
        dm.wait((nb - 1) % 2)
    print("resident", name, hex(flags), "lean" if force_lean else "", "ok", flush=True)


def streamed(text, bcs, m, flags):
    exp = orc.demux(text, bcs, m)
    got = demux.demux_bytes(text, bcs, m, batch_bytes=1 << 18, flags=flags)
    assert got["assign"].tolist() == exp["assign"].tolist(), flags
    for b in range(len(bcs) + 1):
        assert got["per_bin"][b].tolist() == exp["order"][exp["bin_start"][b]:exp["bin_start"][b + 1]].tolist()
    print("streamed", len(text), hex(flags), "ok", flush=True)


def main():
    _ffi.load()
    for name in ("cfg3", "cfg4", "cfg5"):
        resident(name, 20000, 3, _ffi.CFG_OVERLAP)
    resident("cfg3", 20000, 3, _ffi.CFG_OVERLAP, force_lean=True)
    resident("cfg3", 20000, 2, 0)
    resident("cfg5", 20000, 2, _ffi.CFG_NO_FUSE)
    resident("cfg3", 20000, 2, _ffi.CFG_NO_LUT | _ffi.CFG_OVERLAP)
    os.environ.pop("SBR_FORCE_LEAN", None)
    import random
    rng = random.Random(4)
    bcs = [bytes(rng.choices(b"ACGT", k=7)) for _ in range(30)]
    fq = b"".join(b"@r%d x\n" % i + rng.choice(bcs) + bytes(rng.choices(b"ACGTN", k=rng.randint(5, 90))) + b"\n+\n" for i in range(0))
    recs = []
    for i in range(6000):
        s_ = rng.choice(bcs) + bytes(rng.choices(b"ACGTN", k=rng.randint(5, 90)))
        nl = b"\r\n" if i % 97 == 0 else b"\n"
        recs.append(b"@r%d x" % i + nl + s_ + nl + (b"+r%d" % i if i % 53 == 0 else b"+") + nl + bytes(rng.choices(range(33, 74), k=len(s_))) + nl)
    fq = b"".join(recs)
    fa = b"".join(b">s%d\n" % i + rng.choice(bcs) + bytes(rng.choices(b"ACGT", k=rng.randint(1, 200))) + b"\n" for i in range(6000))
    short = b"".join(b"@r\nACGTACG\n+\nIIIIIII\n" for _ in range(20000))  # lines too dense for the region-parallel scan
    for text in (fq, fa, short):
        for flags in (0, _ffi.CFG_EXACT_SCAN, _ffi.CFG_GENERIC, _ffi.CFG_NO_FUSE):
            streamed(text, bcs, 1, flags)
    print("kernel_tour: all ok")


if __name__ == "__main__":
    main()
