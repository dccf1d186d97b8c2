"""Randomised parity hunt on the GPU box: random FASTQ/FASTA shapes x tables x batch sizes x kernel variants, the CUDA
path (C ABI) against the CPU oracle, bit for bit.  Test infrastructure (the oracle is the checker).

    python tests/fuzz_gpu.py [--seconds 150] [--seed 1] [--out gpurun_out/fuzz]

Every failing case is written to <out>/case_<seed>.bin (+ .json with the parameters) and the run exits 1."""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from sabreur_b200 import _ffi, demux as dm  # noqa: E402

# kernel variants every case runs through: the look-back scan, the region-parallel scan with the fused and the separate
# partition, and (CFG_OVERLAP + SBR_FORCE_LEAN) the one-CTA-per-SM grid with the lean 16-bit-counter match kernel
os.environ.setdefault("SBR_FORCE_LEAN", "1")
VARIANTS = ((_ffi.CFG_EXACT_SCAN, 0), (0, 0), (0, _ffi.CFG_NO_FUSE), (0, _ffi.CFG_OVERLAP))

BASES = b"ACGT"
ODD = b"NnacgtRYKM-."


def seq_bytes(rng: random.Random, L: int, p_odd: float) -> bytes:
    s = bytearray(rng.choices(BASES, k=L))
    if p_odd:
        for _ in range(np.random.binomial(L, p_odd) if L > 50 else sum(rng.random() < p_odd for _ in range(L))):
            s[rng.randrange(L)] = rng.choice(ODD)
    return bytes(s)


def make_case(seed: int):
    rng = random.Random(seed)
    np.random.seed(seed & 0x7FFFFFFF)
    fastq = rng.random() < 0.6
    k = rng.choice((3, 4, 5, 6, 7, 8, 8, 10, 12, 13, 16))
    m = rng.choice((0, 0, 1, 1, 2, 3))
    m = min(m, k - 1)
    n_bc = rng.choice((1, 2, 5, 14, 48, 96, 300))
    shape = rng.choice(("short", "mid", "mid", "illumina", "tiny", "long", "mixed"))
    lo, hi = {"short": (k, k + 40), "mid": (k, 400), "illumina": (150, 150), "tiny": (k, k + 2), "long": (20000, 90000),
              "mixed": (k, 400)}[shape]
    target = rng.choice((2_000, 60_000, 400_000, 3_000_000, 12_000_000))
    if shape == "long":
        target = max(target, 400_000)
    p_crlf = rng.choice((0.0, 0.0, 0.02, 1.0))
    p_plus = rng.choice((0.0, 0.0, 0.05, 1.0))
    p_odd = rng.choice((0.0, 0.0, 0.01, 0.15))
    p_qat = rng.choice((0.0, 0.3, 1.0))  # quality / header bytes that look like record markers
    width = rng.choice((0, 0, k - 1, k, k + 1, 3, 60, 70)) if not fastq else 0
    bcs = []
    while len(bcs) < n_bc:
        if bcs and rng.random() < 0.4:  # a near neighbour of an earlier row: first-match ties
            b = bytearray(rng.choice(bcs))
            b[rng.randrange(k)] = rng.choice(BASES)
            b = bytes(b)
        else:
            b = bytes(rng.choices(BASES, k=k))
        if b in bcs and rng.random() < 0.9 and len(bcs) < 4 ** k:
            continue
        bcs.append(b)
    out = bytearray()
    max_rec = 0
    i = 0
    starts = []
    while len(out) < target:
        L = rng.randint(lo, hi)
        if shape == "mixed" and rng.random() < 0.002:
            L = rng.randint(40000, 80000)
        s = seq_bytes(rng, L, p_odd)
        if rng.random() < 0.75:  # tagged read: a table row with a few substitutions / odd bases
            p = bytearray(rng.choice(bcs))
            for _ in range(rng.choice((0, 0, 1, 1, 2, 3))):
                p[rng.randrange(k)] = rng.choice(BASES)
            if p_odd and rng.random() < 0.2:
                p[rng.randrange(k)] = rng.choice(ODD)
            s = bytes(p) + s[k:]
        nl = b"\r\n" if rng.random() < p_crlf else b"\n"
        hdr = (b"r%d" % i) + (b" @x >y +z" if rng.random() < p_qat else b"")
        at = len(out)
        starts.append(at)
        if fastq:
            q = bytearray(rng.choices(range(33, 74), k=L))
            if rng.random() < p_qat:
                q[0] = rng.choice(b"@+>")
            out += b"@" + hdr + nl + s + nl + (b"+" + hdr if rng.random() < p_plus else b"+") + nl + bytes(q) + nl
        else:
            out += b">" + hdr + nl
            w = width if width > 0 else L
            for j in range(0, L, w):
                out += s[j:j + w] + nl
        max_rec = max(max_rec, len(out) - at)
        i += 1
    tail = rng.choice(("nl", "nl", "none", "blank1", "blank3"))
    text = bytes(out)
    inject = None
    if rng.random() < 0.3 and i >= 3:
        # one violation somewhere: the GPU must name the same record as the oracle and keep everything before it
        j = rng.randrange(1, i)
        a, e = starts[j], (starts[j + 1] if j + 1 < i else len(text))
        rec = text[a:e]
        lines = rec.split(b"\n")
        kind = rng.choice(("start", "sep", "unequal", "short", "trunc") if fastq else ("short", "short_last", "empty", "empty"))
        if kind == "start":
            rec = b"X" + rec[1:]
        elif kind == "sep":
            lines[2] = b"-" + lines[2][1:]
            rec = b"\n".join(lines)
        elif kind == "unequal":
            cr = lines[3].endswith(b"\r")
            lines[3] = lines[3][:-2] + b"\r" if cr else lines[3][:-1]
            rec = b"\n".join(lines)
        elif kind == "short":
            rec = (b"@s\n" + b"A" * (k - 1) + b"\n+\n" + b"I" * (k - 1) + b"\n") if fastq else (b">s\n" + b"A" * (k - 1) + b"\n")
        elif kind == "empty":  # header directly followed by the next record's '>' (or by a blank line)
            rec = lines[0] + b"\n" + rng.choice((b"", b"", b"\n", b"\r\n"))
        elif kind == "short_last":
            a = e = len(text)
            starts.append(a)
            rec = b">s\n" + b"A" * (k - 1) + b"\n"
        elif kind == "trunc":
            a, e = starts[i - 1], len(text)
            lines = text[a:e].split(b"\n")
            rec = b"\n".join(lines[:rng.choice((1, 2, 3))]) + rng.choice((b"", b"\n"))
            if rec.count(b"\n") == 3:
                rec = rec[:-1]
        text = text[:a] + rec + text[e:]
        inject = kind
        tail = "nl"
    if tail == "none":
        text = text.rstrip(b"\r\n")
    elif tail == "blank1":
        text += b"\n"
    elif tail == "blank3" and fastq:
        text += b"\n\n\n"
    batch = rng.choice((1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24))
    batch = max(batch, 4 * max_rec + 4096)
    chunk = rng.choice((4099, 50_001, 1 << 20))
    user_flags = rng.choice((0, 0, _ffi.CFG_NO_LUT, _ffi.CFG_GENERIC))
    params = dict(seed=seed, fastq=fastq, k=k, m=m, n_bc=len(bcs), shape=shape, nbytes=len(text), n_rec=i, p_crlf=p_crlf,
                  p_plus=p_plus, p_odd=p_odd, p_qat=p_qat, width=width, tail=tail, batch=batch, chunk=chunk,
                  user_flags=user_flags, inject=inject)
    return text, bcs, m, batch, chunk, user_flags, params, starts


CODE_BITS = {orc.E_INVALID_START: _ffi.ST_INVALID_START, orc.E_INVALID_SEP: _ffi.ST_INVALID_SEP,
             orc.E_UNEQUAL_LEN: _ffi.ST_UNEQUAL_LEN, orc.E_TRUNCATED: _ffi.ST_TRUNCATED, orc.E_SHORT_READ: _ffi.ST_SHORT_READ}


def check_violation(text, bcs, m, batch, chunk, user_flags, starts):
    try:
        orc.demux(text, bcs, m)
        return "oracle accepted the damaged input (generator bug)"
    except orc.OracleError as e:
        err = e
    if err.record > len(starts):
        return f"oracle names record {err.record} of {len(starts)}"
    cut = starts[err.record] if err.record < len(starts) else None
    exp = orc.demux(text[:cut], bcs, m) if (cut is None or cut > 0) and err.record > 0 else None
    for scan_flag, extra in VARIANTS:
        got = dm.demux_bytes(text, bcs, m, batch_bytes=batch, chunk=chunk, flags=user_flags | scan_flag | extra)
        tag = f"flags={user_flags | scan_flag | extra:#x}"
        if got["bad"] is None:
            return f"{tag}: no violation reported (oracle: code {err.code} record {err.record})"
        if got["bad"][0] != err.record:
            return f"{tag}: violation at record {got['bad'][0]}, oracle says {err.record} (code {err.code})"
        if not (got["bad"][1] & CODE_BITS.get(err.code, 0)):
            return f"{tag}: violation bits {got['bad'][1]:#x}, oracle code {err.code}"
        if exp is not None and got["assign"].tolist() != exp["assign"].tolist():
            return f"{tag}: records before the violation differ"
    return None


def check(text, bcs, m, batch, chunk, user_flags):
    exp = orc.demux(text, bcs, m)
    for scan_flag, extra in VARIANTS:
        got = dm.demux_bytes(text, bcs, m, batch_bytes=batch, chunk=chunk, flags=user_flags | scan_flag | extra)
        tag = f"flags={user_flags | scan_flag | extra:#x}"
        if got["bad"] is not None or (got["status"] & _ffi.ST_ERROR_MASK):
            return f"{tag}: error reported: bad={got['bad']} status={got['status']:#x}"
        if got["assign"].tolist() != exp["assign"].tolist():
            d = np.flatnonzero(got["assign"][:len(exp["assign"])] != exp["assign"][:len(got["assign"])])
            return f"{tag}: assign differs (n {len(got['assign'])} vs {len(exp['assign'])}, first at {d[:3].tolist()})"
        if got["bin_count"].tolist() != exp["bin_count"].tolist():
            return f"{tag}: bin_count differs"
        if got["spans"][:, 0].tolist() != exp["recs"]["rec_off"].tolist() or got["spans"][:, 1].tolist() != exp["recs"]["rec_end"].tolist():
            return f"{tag}: record spans differ"
        for b in range(len(bcs) + 1):
            seg = exp["order"][exp["bin_start"][b]:exp["bin_start"][b + 1]]
            if got["per_bin"][b].tolist() != seg.tolist():
                return f"{tag}: order of bin {b} differs"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/fuzz")
    ap.add_argument("--only", type=int, default=None, help="run exactly this case seed")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    _ffi.load()
    t0 = time.time()
    seed = a.seed * 1_000_003
    n = fails = 0
    total = 0
    while (time.time() - t0 < a.seconds) if a.only is None else n == 0:
        s = a.only if a.only is not None else seed + n
        text, bcs, m, batch, chunk, uf, params, starts = make_case(s)
        try:
            why = check_violation(text, bcs, m, batch, chunk, uf, starts) if params["inject"] else check(text, bcs, m, batch, chunk, uf)
        except Exception as e:  # noqa: BLE001 -- a crash is a finding too
            why = f"exception: {e!r}"
        n += 1
        total += len(text)
        if why:
            fails += 1
            params["why"] = why
            params["barcodes"] = [b.decode() for b in bcs]
            with open(os.path.join(a.out, f"case_{s}.json"), "w") as f:
                json.dump(params, f)
            if len(text) <= 4 << 20:
                with open(os.path.join(a.out, f"case_{s}.bin"), "wb") as f:
                    f.write(text)
            print("FAIL", json.dumps(params)[:600], flush=True)
            if fails >= 5:
                break
    print(f"fuzz: {n} cases, {total / 1e6:.1f} MB, {fails} failures, {time.time() - t0:.0f} s", flush=True)
    with open(os.path.join(a.out, "summary.json"), "w") as f:
        json.dump(dict(cases=n, megabytes=total / 1e6, failures=fails, seconds=time.time() - t0, seed=a.seed), f)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
