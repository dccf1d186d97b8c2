#!/usr/bin/env python
"""Writes tests/golden/synth_goldens.json: for a fixed slice of each synthetic workload (cfg3, cfg4 both
mates, cfg5) the digests of what the ORACLE says -- sha256 of the text, of the per-read assignments (u16 little endian),
of the per-bin index lists (u32, bins in order) and the per-bin counts.  Both the oracle (tests/test_oracle.py) and the GPU
path (tests/test_gpu_parity.py) are compared with the committed file, so the two cannot drift together unnoticed.

    python tests/golden/make_synth_goldens.py        (CPU only; run it again only when the workload definition changes)"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from sabreur_b200 import synth  # noqa: E402

FIRST, N = 424_242, 20_000


def digest(name: str, mate: int) -> dict:
    w = synth.WORKLOADS[name]
    tab = orc.synth_table(w.cfg_id, w.n_barcodes, w.bc_len, w.min_dist)
    text = orc.synth_reads(w.fmt, w.read_len, mate, synth.SEED_READS + w.cfg_id, tab, FIRST, N)
    exp = orc.demux(text, [bytes(r) for r in tab], w.mismatch)
    return {"workload": name, "mate": mate, "first": FIRST, "n": N, "mismatch": w.mismatch,
            "text_sha256": hashlib.sha256(text.tobytes()).hexdigest(),
            "table_sha256": hashlib.sha256(tab.tobytes()).hexdigest(),
            "assign_sha256": hashlib.sha256(exp["assign"].astype("<u2").tobytes()).hexdigest(),
            "order_sha256": hashlib.sha256(exp["order"].astype("<u4").tobytes()).hexdigest(),
            "unknown": int(exp["bin_count"][-1]), "bin_count_head": [int(x) for x in exp["bin_count"][:8]]}


if __name__ == "__main__":
    out = [digest("cfg3", 0), digest("cfg4", 0), digest("cfg4", 1), digest("cfg5", 0)]
    with open(os.path.join(HERE, "synth_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    for d in out:
        print(d["workload"], d["mate"], d["assign_sha256"][:16], d["unknown"])
