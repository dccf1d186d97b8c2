"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact: assignments, record offsets, per-bin counts, per-bin index lists, per-bin output bytes.
Run on the B200 box: python -m pytest tests -m gpu"""
import hashlib
import random

import numpy as np
import pytest

from conftest import fixture_bytes, read_barcode_table
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (there is no CPU fallback to test)")
    from sabreur_b200 import _ffi, demux, synth
    _ffi.load()
    return dict(torch=torch, ffi=_ffi, demux=demux, synth=synth)


def _bcs(name):
    return [r[0].encode() for r in read_barcode_table(name)]


def _check_against_oracle(gpu, text: bytes, barcodes, m, **kw):
    """stream `text` through the GPU and compare everything with the oracle"""
    dm = gpu["demux"]
    exp = orc.demux(text, barcodes, m)
    user_flags = kw.pop("flags", 0)
    # both record scans: the region-parallel one (default) and the look-back one (forced); match + partition as one
    # cooperative launch (default) and as separate launches
    ffi = gpu["ffi"]
    for scan_flag, extra in ((ffi.CFG_EXACT_SCAN, 0), (0, 0), (0, ffi.CFG_NO_FUSE)):
        got = dm.demux_bytes(text, barcodes, m, flags=user_flags | scan_flag | extra, **kw)
        assert got["bad"] is None and not (got["status"] & gpu["ffi"].ST_ERROR_MASK)
        assert got["assign"].tolist() == exp["assign"].tolist()
        assert got["bin_count"].tolist() == exp["bin_count"].tolist()
        assert got["spans"][:, 0].tolist() == exp["recs"]["rec_off"].tolist()
        assert got["spans"][:, 1].tolist() == exp["recs"]["rec_end"].tolist()
        for b in range(len(barcodes) + 1):
            seg = exp["order"][exp["bin_start"][b]:exp["bin_start"][b + 1]]
            assert got["per_bin"][b].tolist() == seg.tolist()
        if scan_flag:
            assert not any(got["scan_paths"])
    return got, exp


# ---- generator: device bytes == host bytes ---------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg5"])
def test_synth_device_equals_host(gpu, name):
    torch, synth = gpu["torch"], gpu["synth"]
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    n, first = 5000, 123456789
    for mate in ((0, 1) if w.paired else (0,)):
        host = synth.reads_host(w, tab, first, n, mate)
        dev = torch.empty(n * w.rec_bytes + 64, dtype=torch.uint8, device="cuda")
        synth.reads_device(w, tab, first, n, dev.data_ptr(), mate=mate)
        assert np.array_equal(dev[: n * w.rec_bytes].cpu().numpy(), host)


# ---- the first-match LUT is the oracle's find() over every possible prefix -------------------------
@pytest.mark.parametrize("k,n,m", [(4, 7, 1), (6, 40, 2), (8, 96, 1), (5, 9, 0)])
def test_lut_equals_first_match_over_all_prefixes(gpu, k, n, m):
    rng = random.Random(k * 1000 + n)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(k)) for _ in range(n)]
    bcs[-1] = bcs[0]  # a duplicate row: the earlier row must win
    with gpu["demux"].Demux(bcs, m, fmt=2, batch_bytes=1 << 20) as dm:
        lut = dm.lut()
    assert lut is not None and lut.size == 4 ** k
    idx = np.arange(4 ** k, dtype=np.uint32)
    code2base = np.frombuffer(b"ACTG", dtype=np.uint8)  # code = (ascii >> 1) & 3
    pref = np.stack([code2base[(idx >> (2 * j)) & 3] for j in range(k)], 1)
    tab = np.frombuffer(b"".join(bcs), dtype=np.uint8).reshape(n, k)
    mism = (pref[:, None, :] != tab[None, :, :]).sum(-1)
    ok = mism <= m
    exp = np.where(ok.any(1), ok.argmax(1), n).astype(np.uint16)
    assert np.array_equal(lut, exp)


# ---- synthetic slices of the BASELINE shapes ---------------------------------------------------------
@pytest.mark.parametrize("name,n", [("cfg3", 60000), ("cfg4", 40000), ("cfg5", 50000)])
@pytest.mark.parametrize("flags", [0, 1])  # LUT and the plain XOR+popc table walk
def test_synthetic_slice_parity(gpu, name, n, flags):
    synth = gpu["synth"]
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    for mate in ((0, 1) if w.paired else (0,)):
        text = synth.reads_host(w, tab, 777, n, mate).tobytes()
        got, exp = _check_against_oracle(gpu, text, bcs, w.mismatch, batch_bytes=4 << 20, flags=flags)
        assert got["bin_count"].sum() == n
        assert all(got["scan_paths"]), "clean synthetic batches must finish in the region-parallel scan"


def _synth_goldens():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_goldens.json")))


@pytest.mark.parametrize("g", _synth_goldens(), ids=lambda g: f"{g['workload']}-mate{g['mate']}")
def test_gpu_reproduces_the_committed_synthetic_goldens(gpu, g):
    """the CUDA path against digests COMMITTED in tests/golden/synth_goldens.json (made from the oracle by
    make_synth_goldens.py): device generator, scan, match and partition, without the oracle in the loop"""
    torch, synth, dmx, ffi = gpu["torch"], gpu["synth"], gpu["demux"], gpu["ffi"]
    w = synth.WORKLOADS[g["workload"]]
    tab = synth.table(w)
    assert hashlib.sha256(tab.tobytes()).hexdigest() == g["table_sha256"]
    n, rec = g["n"], w.rec_bytes
    dev = torch.empty(n * rec + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, g["first"], n, dev.data_ptr(), mate=g["mate"])
    assert hashlib.sha256(dev[: n * rec].cpu().numpy().tobytes()).hexdigest() == g["text_sha256"]
    for flags in (0, ffi.CFG_OVERLAP, ffi.CFG_NO_FUSE | ffi.CFG_NO_LUT):
        with dmx.Demux([bytes(r) for r in tab], g["mismatch"], fmt=w.fmt, batch_bytes=n * rec + 4096, flags=flags) as dm:
            dm.demux_device(0, dev.data_ptr(), n * rec, final=True)
            b = dm.wait(0)
        assert b.n_records == n and b.status == 0
        assert hashlib.sha256(b.assign.astype("<u2").tobytes()).hexdigest() == g["assign_sha256"]
        assert hashlib.sha256(b.order.astype("<u4").tobytes()).hexdigest() == g["order_sha256"]
        assert int(b.bin_count[-1]) == g["unknown"] and [int(x) for x in b.bin_count[:8]] == g["bin_count_head"]


def test_device_resident_batch_equals_oracle(gpu):
    """sbr_demux_device on HBM-resident text (the benchmark path)"""
    torch, synth, dmx = gpu["torch"], gpu["synth"], gpu["demux"]
    w = synth.CFG3
    tab = synth.table(w)
    n = 200_000
    dev = torch.empty(n * w.rec_bytes + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 0, n, dev.data_ptr())
    text = dev[: n * w.rec_bytes].cpu().numpy()
    bcs = [bytes(r) for r in tab]
    with dmx.Demux(bcs, w.mismatch, fmt=2, batch_bytes=n * w.rec_bytes + 4096) as dm:
        dm.demux_device(0, dev.data_ptr(), n * w.rec_bytes, final=True, stream=torch.cuda.current_stream().cuda_stream)
        b = dm.wait(0)
    exp = orc.demux(text, bcs, w.mismatch)
    assert b.n_records == n and b.consumed == n * w.rec_bytes and b.status == 0 and b.scan_path == 1
    assert np.array_equal(b.assign, exp["assign"])
    assert np.array_equal(b.bin_count, exp["bin_count"])
    assert np.array_equal(b.bin_start, exp["bin_start"])
    assert np.array_equal(b.order, exp["order"])
    assert np.array_equal(b.rec_off[:-1].astype(np.uint64), exp["recs"]["rec_off"])


@pytest.mark.parametrize("name,n,nb", [("cfg3", 30_000, 7), ("cfg4", 20_000, 5), ("cfg5", 25_000, 6)])
@pytest.mark.parametrize("extra", [0, 0x10])  # fused cooperative launch / separate launches
def test_overlapped_pipeline_equals_oracle(gpu, name, n, nb, extra):
    """SBR_CFG_OVERLAP (the benchmark's HBM-resident loop): batches are launched back to back on one stream and rotate
    through three device workspaces; match + partition of a batch run on the slot's own stream beside the next scan.
    Every batch must come out exactly as the oracle's, whatever ran concurrently."""
    torch, synth, dmx, ffi = gpu["torch"], gpu["synth"], gpu["demux"], gpu["ffi"]
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec = w.rec_bytes
    dev = torch.empty(n * nb * rec + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 5_000_000, n * nb, dev.data_ptr())
    host = dev[: n * nb * rec].cpu().numpy()
    st = torch.cuda.Stream()
    S = 3
    with dmx.Demux(bcs, w.mismatch, fmt=w.fmt, batch_bytes=n * rec + 4096, n_slots=S, max_records=n + 16,
                   flags=ffi.CFG_OVERLAP | ffi.CFG_TIMING | extra) as dm:
        for rnd in range(2):  # the second round reuses every slot without a host-side wait in between
            for i in range(nb):
                dm.demux_device(i % S, dev.data_ptr() + i * n * rec, n * rec, final=True, stream=st.cuda_stream)
        dm.join(st.cuda_stream)
        st.synchronize()
        assert dm.scan_stats() == (2 * nb, 0)
        for i in range(nb - S, nb):  # the slots hold the last three batches
            b = dm.wait(i % S)
            exp = orc.demux(host[i * n * rec:(i + 1) * n * rec], bcs, w.mismatch)
            assert b.n_records == n and b.status == 0 and b.scan_path == 1 and b.consumed == n * rec
            assert np.array_equal(b.assign, exp["assign"])
            assert np.array_equal(b.bin_count, exp["bin_count"]) and np.array_equal(b.bin_start, exp["bin_start"])
            assert np.array_equal(b.order, exp["order"])
            assert np.array_equal(b.rec_off[:-1].astype(np.uint64), exp["recs"]["rec_off"])
        ms, nbatches, launches = dm.timing()
        assert nbatches == 2 * nb and all(t > 0 for t in ms[:2])
        # one at a time as well (submit, wait, submit ...), every batch checked
        for i in range(nb):
            dm.demux_device(i % S, dev.data_ptr() + i * n * rec, n * rec, final=True, stream=st.cuda_stream)
            b = dm.wait(i % S)
            exp = orc.demux(host[i * n * rec:(i + 1) * n * rec], bcs, w.mismatch)
            assert np.array_equal(b.assign, exp["assign"]) and np.array_equal(b.order, exp["order"])


@pytest.mark.parametrize("name,flags", [("cfg3", 0), ("cfg3", 1), ("cfg4", 0), ("cfg5", 0), ("cfg5", 0x10)])
def test_lean_match_kernel_equals_oracle(gpu, name, flags, monkeypatch):
    """the match kernel's shared-memory diet (16-bit per-warp counters, CTA base + 16-bit offsets as cursors, table and
    region prefix read from global memory) -- what big tables (cfg5: 1537 bins) use beside the scan -- forced onto every
    table size, LUT and table-walk matching"""
    torch, synth, dmx, ffi = gpu["torch"], gpu["synth"], gpu["demux"], gpu["ffi"]
    monkeypatch.setenv("SBR_FORCE_LEAN", "1")
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec = w.rec_bytes
    n, nb = 40_000, 4
    dev = torch.empty(n * nb * rec + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 31_000_000, n * nb, dev.data_ptr())
    host = dev[: n * nb * rec].cpu().numpy()
    st = torch.cuda.Stream()
    with dmx.Demux(bcs, w.mismatch, fmt=w.fmt, batch_bytes=n * rec + 4096, n_slots=2, max_records=n + 16,
                   flags=ffi.CFG_OVERLAP | flags) as dm:
        for i in range(nb):
            dm.demux_device(i % 2, dev.data_ptr() + i * n * rec, n * rec, final=True, stream=st.cuda_stream)
            if i >= 1:
                b = dm.wait((i - 1) % 2)
                exp = orc.demux(host[(i - 1) * n * rec:i * n * rec], bcs, w.mismatch)
                assert b.n_records == n and b.status == 0
                assert np.array_equal(b.assign, exp["assign"]) and np.array_equal(b.bin_count, exp["bin_count"])
                assert np.array_equal(b.bin_start, exp["bin_start"]) and np.array_equal(b.order, exp["order"])
        dm.wait((nb - 1) % 2)


@pytest.mark.parametrize("lean", [False, True])
def test_every_read_in_one_bin(gpu, lean, monkeypatch):
    """the counters' worst case: all records of every warp slice and every CTA fall into ONE bin (here: unknown, and then
    the first barcode), with the plain and with the lean (16-bit) match kernel, one CTA per SM"""
    torch, dmx, ffi = gpu["torch"], gpu["demux"], gpu["ffi"]
    if lean:
        monkeypatch.setenv("SBR_FORCE_LEAN", "1")
    rng = random.Random(8)
    bcs = [b"ACGTACGTAC"] + [bytes(rng.choices(b"ACGT", k=10)) for _ in range(1500)]
    n = 400_000
    for pre, want_bin in ((b"NNNNNNNNNN", len(bcs)), (b"ACGTACGTAC", 0)):
        text = b"".join(b">r%07d\n" % i + pre + b"ACGTTGCATTGACCA" * 4 + b"\n" for i in range(n))
        t = torch.frombuffer(bytearray(text + b"\0" * 64), dtype=torch.uint8).cuda()
        with dmx.Demux(bcs, 1, fmt=1, batch_bytes=len(text) + 4096, max_records=n + 16, flags=ffi.CFG_OVERLAP) as dm:
            dm.demux_device(0, t.data_ptr(), len(text), final=True)
            b = dm.wait(0)
        assert b.n_records == n and b.status == 0 and b.scan_path == 1
        assert int(b.bin_count[want_bin]) == n and int(b.bin_count.sum()) == n
        assert np.array_equal(b.order, np.arange(n, dtype=np.uint32))
        assert (b.assign == want_bin).all()


def test_scan_stats_count_handed_over_batches(gpu):
    """the device-side counters bench.py reads: a malformed batch is handed to the exact scan and counted"""
    torch, dmx, ffi = gpu["torch"], gpu["demux"], gpu["ffi"]
    seq = b"ACGTACGTAC" * 6  # (lines shorter than ~16 bytes on average are the exact scan's job by design)
    recs = [b"@r%d\n" % i + seq + b"\n+\n" + b"I" * len(seq) + b"\n" for i in range(5000)]
    good = b"".join(recs)
    bad = b"".join(recs[:1000]) + b"@x\nACGT\n+\nIII\n" + b"".join(recs[1000:])
    bcs = [b"ACGT", b"TTTT"]
    with dmx.Demux(bcs, 0, fmt=2, batch_bytes=1 << 20) as dm:
        for text, fails in ((good, 0), (bad, 1)):
            t = torch.frombuffer(bytearray(text + b"\0" * 64), dtype=torch.uint8).cuda()
            dm.scan_stats(True)
            dm.demux_device(0, t.data_ptr(), len(text), final=True)
            fin, handed = dm.scan_stats()
            assert (fin, handed) == (1 - fails, fails)
            b = dm.wait(0)
            assert b.scan_path == (0 if fails else 1)
            assert bool(b.status & ffi.ST_ERROR_MASK) == bool(fails)


def test_missing_final_newline_ships_every_scan_word(gpu):
    """when only the host knows that the batch is non-canonical (it supplied the stream's final newline),
    rec_key must still hold the scan word of EVERY record, not just the patched last one"""
    dmx, ffi = gpu["demux"], gpu["ffi"]
    recs = [(b"r%d" % i, bytes(b"ACGT"[(i >> (2 * j)) & 3] for j in range(10))) for i in range(3000)]
    text = b"".join(b"@" + h + b"\n" + s + b"\n+\n" + b"I" * 10 + b"\n" for h, s in recs)[:-1]
    bcs = [b"ACGTAC", b"TTTTTT"]
    with dmx.Demux(bcs, 0, batch_bytes=1 << 20) as dm:
        buf = dm.slot_buffer(0)
        buf[:len(text)] = np.frombuffer(text, dtype=np.uint8)
        dm.submit(0, 0, len(text), 0, True)
        b = dm.wait(0)
    assert b.n_records == len(recs) and b.status & ffi.ST_NONCANONICAL and b.rec_key is not None
    code = {65: 0, 67: 1, 84: 2, 71: 3}
    for i in (0, 1, 17, 1500, len(recs) - 2, len(recs) - 1):
        want = sum(code[c] << (2 * j) for j, c in enumerate(recs[i][1][:6]))
        assert int(b.rec_key[i]) & 0xFFF == want, i
        assert bool(int(b.rec_key[i]) & ffi.KEY_NONCANON) == (i == len(recs) - 1)
    assert b.consumed == len(text)


# ---- the in-tree fixtures (SURVEY 8c goldens) -------------------------------------------------------
R1_GOLD = ["90d9764b8b69432c", "dd06fd51837741d1", "98a7be33c3ac7170", "5af261ae384d0e8d", "cf6171f02f1974ce",
           "8a8208327d1ba4cc", "d88301f0531c93a4", "9e490c81cc82b11e", "88d17addc88d6666", "5a32eafb431d020f"]


def _outputs(gpu, text, got, n_bins):
    outs = []
    for b in range(n_bins):
        outs.append(b"".join(text[int(s):int(e)] for s, e in got["spans"][got["per_bin"][b].astype(np.int64)]))
    return outs


@pytest.mark.parametrize("m", [0, 1, 2])
def test_reads_1_fixture(gpu, reads_1, m):
    bcs = _bcs("bc_se.txt")
    got, exp = _check_against_oracle(gpu, reads_1, bcs, m, batch_bytes=1 << 20)
    outs = _outputs(gpu, reads_1, got, len(bcs) + 1)
    assert outs == orc.demux_outputs(reads_1, bcs, m)
    if m == 0:
        assert [hashlib.sha256(o).hexdigest()[:16] for o in outs[:-1]] == R1_GOLD
        assert got["bin_count"].tolist() == [1000] + [999] * 9 + [0]


def test_test_fq_fixture(gpu):
    text = fixture_bytes("test.fq")
    got, _ = _check_against_oracle(gpu, text, _bcs("bc_se.txt"), 0, batch_bytes=1 << 16)
    outs = _outputs(gpu, text, got, 11)
    assert all(o == b"" for o in outs[:10]) and outs[10] == text
    got, _ = _check_against_oracle(gpu, text, [b"ACCGTA", b"ATTGTT"], 1, batch_bytes=1 << 16)
    assert got["assign"].tolist() == [0, 1, 2]


# ---- parser edge cases --------------------------------------------------------------------------------
def _rand_fastq(rng, n, crlf=False, plus_id=False, final_nl=True, lo=6, hi=60, alphabet=b"ACGTN"):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i in range(n):
        L = rng.randint(lo, hi)
        s = bytes(rng.choice(alphabet) for _ in range(L))
        q = bytes(rng.randint(33, 73) for _ in range(L))  # '@' '+' '>' all occur in qualities
        out += b"@r%d x" % i + nl + s + nl + (b"+r%d" % i if plus_id else b"+") + nl + q + nl
    if not final_nl:
        out = out[:-len(nl)]
    return bytes(out)


def _rand_fasta(rng, n, width=0, crlf=False, lo=6, hi=90):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i in range(n):
        L = rng.randint(lo, hi)
        s = bytes(rng.choice(b"ACGT") for _ in range(L))
        out += b">s%d desc" % i + nl
        if width:
            for j in range(0, L, width):
                out += s[j:j + width] + nl
        else:
            out += s + nl
    return bytes(out)


CASES = {
    "fq": lambda r: _rand_fastq(r, 3000),
    "fq_crlf": lambda r: _rand_fastq(r, 3000, crlf=True),
    "fq_plus": lambda r: _rand_fastq(r, 3000, plus_id=True),
    "fq_nofinal": lambda r: _rand_fastq(r, 3000, final_nl=False),
    "fq_blank_tail": lambda r: _rand_fastq(r, 500) + b"\n\n\n",
    "fq_tiny_reads": lambda r: _rand_fastq(r, 20000, lo=4, hi=5),     # > 2048 newlines per 12 KB tile
    "fq_long_reads": lambda r: _rand_fastq(r, 40, lo=20000, hi=60000),  # lines longer than a tile
    "fq_lower": lambda r: _rand_fastq(r, 2000, alphabet=b"ACGTacgtNRY"),
    "fa": lambda r: _rand_fasta(r, 4000),
    "fa_w7": lambda r: _rand_fasta(r, 4000, width=7),
    "fa_w3": lambda r: _rand_fasta(r, 4000, width=3),                  # the match window spans lines
    "fa_crlf": lambda r: _rand_fasta(r, 3000, width=11, crlf=True),
    "fa_long": lambda r: _rand_fasta(r, 30, lo=30000, hi=70000, width=60),
    "fa_blank_tail": lambda r: _rand_fasta(r, 300) + b"\n\n",
}


@pytest.mark.parametrize("kind", sorted(CASES))
@pytest.mark.parametrize("m,big", [(0, False), (1, True)])
def test_edge_inputs(gpu, kind, m, big):
    rng = random.Random(sum(kind.encode()) * 7 + m)
    text = CASES[kind](rng)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(4)) for _ in range(14)]
    # big: one multi-region batch (several CTAs of the region-parallel scan); small: many single-region batches
    batch = (1 << 20) if ("long" in kind or big) else (1 << 16)
    got, exp = _check_against_oracle(gpu, text, bcs, m, batch_bytes=batch, chunk=50_001)
    # per-bin output bytes through the host writer's rule: verbatim unless flagged non-canonical
    outs_exp = orc.demux_outputs(text, bcs, m)
    assert sum(len(o) for o in outs_exp) > 0


@pytest.mark.parametrize("batch", [1 << 12, 4608, 12288, 12288 + 16, 3 * 12288 - 1, 36864, 36864 + 16, 2 * 36864 - 1, 50_000,
                                   4 * 36864 + 48])
def test_batch_and_tile_boundaries(gpu, batch):
    """records straddle warp slices (4.5 KB), tiles (36 KB; 12 KB in the look-back scan) and batches at every alignment"""
    rng = random.Random(batch)
    text = _rand_fastq(rng, 2500, lo=30, hi=200)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(6)) for _ in range(20)]
    _check_against_oracle(gpu, text, bcs, 1, batch_bytes=batch, chunk=777)
    text = _rand_fasta(rng, 2500, lo=30, hi=200, width=50)
    _check_against_oracle(gpu, text, bcs, 1, batch_bytes=batch, chunk=777)


def test_ragged_and_non_acgt_tables(gpu):
    rng = random.Random(5)
    text = _rand_fastq(rng, 3000, alphabet=b"ACGTN")
    # ragged lengths: zip() truncation (packed path with care masks)
    bcs = [b"ACGTAC", b"ACG", b"TTTTTTTT", b"GA", b"CCCCC"]
    _check_against_oracle(gpu, text, bcs, 1, batch_bytes=1 << 16)
    # non-ACGT barcode bytes: an 'N' in the table only matches an 'N' in the read (byte-exact kernel)
    bcs = [b"ACNT", b"NNNN", b"acgt", b"TTTT"]
    _check_against_oracle(gpu, text, bcs, 1, batch_bytes=1 << 16)
    # long barcodes (> 16 nt) take the byte-exact kernel too
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(24)) for _ in range(10)]
    text = _rand_fastq(rng, 2000, lo=24, hi=60)
    _check_against_oracle(gpu, text, bcs, 3, batch_bytes=1 << 16)
    text = _rand_fasta(rng, 2000, lo=24, hi=90, width=10)
    _check_against_oracle(gpu, text, bcs, 3, batch_bytes=1 << 16)


@pytest.mark.parametrize("text,code", [
    (b"@a\nACGT\n+\nIII\n", "unequal"),
    (b"@a\nACGT\n-\nIIII\n", "sep"),
    (b"@a\nACGT\n+\nIIII\nXb\nACGT\n+\nIIII\n", "start"),
    (b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nII\n", "short"),
    (b"@a\nACGT\n+\nIIII\n@b\nACGT\n+", "trunc"),
    (b"@a\nACGT\n+\nIIII\n\n\n\n\n", "start"),
    (b"ACGT\n", "start"),
    (b">a\nAC\n>b\nACGT\n", "short"),
    (b">a\n>b\nACGT\n", "short"),              # an empty sequence: the next record starts right behind the header
    (b">x\nACGT\n>a\n>bbbbbbbbbbbbbbbbbb\nACGTACGT\n>c\nACGT\n", "short"),
    (b">a\n\n>b\nACGT\n", "short"),
])
def test_parse_violations_are_reported_like_the_oracle(gpu, text, code):
    ffi, dm = gpu["ffi"], gpu["demux"]
    bits = {"unequal": ffi.ST_UNEQUAL_LEN, "sep": ffi.ST_INVALID_SEP, "start": ffi.ST_INVALID_START,
            "short": ffi.ST_SHORT_READ, "trunc": ffi.ST_TRUNCATED}[code]
    with pytest.raises(orc.OracleError) as e:
        orc.demux(text, [b"ACGT", b"TTTT"], 0)
    for flags in (0, ffi.CFG_EXACT_SCAN, ffi.CFG_GENERIC):
        got = dm.demux_bytes(text, [b"ACGT", b"TTTT"], 0, batch_bytes=1 << 16, flags=flags)
        assert got["bad"] is not None and got["bad"][1] & bits
        if code != "start" or text[:1] in b"@>":
            assert got["bad"][0] == e.value.record


def test_fasta_empty_sequences_inside_a_large_batch(gpu):
    """records whose header is directly followed by the next '>' (ADVICE r1): a short read at exactly that record,
    wherever it falls in a region / slice of the region-parallel scan"""
    rng = random.Random(5)
    ffi, dm = gpu["ffi"], gpu["demux"]
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(6)) for _ in range(12)]
    for bad_at in (1, 777, 20011, 39998):
        out = bytearray()
        for i in range(40000):
            L = rng.randint(20, 70)
            out += b">r%d\n" % i
            if i != bad_at:
                out += bytes(rng.choice(b"ACGT") for _ in range(L)) + b"\n"
        text = bytes(out)
        with pytest.raises(orc.OracleError) as e:
            orc.demux(text, bcs, 1)
        assert e.value.record == bad_at
        for flags in (0, ffi.CFG_EXACT_SCAN):
            got = dm.demux_bytes(text, bcs, 1, batch_bytes=4 << 20, flags=flags)
            assert got["bad"] == (bad_at, ffi.ST_SHORT_READ)


def test_adversarial_phase_detection(gpu):
    """quality strings made of '@' and headers/sequences that mimic other lines: whatever phase a region
    guesses, the result must be the exact scan's (the scan's epilogue verifies, the look-back scan re-runs)"""
    rng = random.Random(77)
    out = bytearray()
    for i in range(40000):
        L = rng.randint(4, 9)
        s = bytes(rng.choice(b"ACGT") for _ in range(L))
        q = b"@" * L if rng.random() < 0.5 else b"+" * L
        out += b"@" + b"@" * rng.randint(0, 3) + b"\n" + s + b"\n+" + (b"@x" if rng.random() < 0.3 else b"") + b"\n" + q + b"\n"
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(4)) for _ in range(9)]
    _check_against_oracle(gpu, bytes(out), bcs, 1, batch_bytes=1 << 20, chunk=333_333)


def test_records_larger_than_the_reserved_room_in_front_of_a_slot(gpu):
    """Demux.stream reads the next batch's fresh bytes behind a reserved quarter of the slot while the GPU works; a tail
    (incomplete record) larger than that room makes it slide the fresh bytes back and hand the overflow to the reader"""
    rng = random.Random(21)
    bcs = [bytes(rng.choices(b"ACGT", k=8)) for _ in range(20)]
    out = bytearray()
    for i in range(400):
        L = rng.choice((40, 200, 9000, 13000))  # 26 KB records against 64 KB batches (16 KB reserved)
        s_ = rng.choice(bcs) + bytes(rng.choices(b"ACGT", k=L))
        out += b"@r%d\n" % i + s_ + b"\n+\n" + bytes(rng.choices(range(35, 74), k=len(s_))) + b"\n"
    _check_against_oracle(gpu, bytes(out), bcs, 1, batch_bytes=1 << 16, chunk=7001)
    fa = bytearray()
    for i in range(300):
        fa += b">r%d\n" % i + rng.choice(bcs) + bytes(rng.choices(b"ACGT", k=rng.choice((30, 20000, 27000)))) + b"\n"
    _check_against_oracle(gpu, bytes(fa), bcs, 0, batch_bytes=1 << 16, chunk=100_000)


def test_record_capacity_overflow_is_chunked_not_lost(gpu):
    rng = random.Random(9)
    text = _rand_fastq(rng, 5000, lo=8, hi=12)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(4)) for _ in range(8)]
    _check_against_oracle(gpu, text, bcs, 0, batch_bytes=1 << 16, max_records=300)


# ---- size-independent properties at a large size ------------------------------------------------------
def test_large_batch_properties(gpu):
    torch, synth, dmx = gpu["torch"], gpu["synth"], gpu["demux"]
    w = synth.CFG3
    tab = synth.table(w)
    n = 2_000_000  # 642 MB of text, one batch
    nbytes = n * w.rec_bytes
    dev = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 10_000_000, n, dev.data_ptr())
    bcs = [bytes(r) for r in tab]
    with dmx.Demux(bcs, w.mismatch, fmt=2, batch_bytes=nbytes + 4096, max_records=n + 16) as dm:
        dm.demux_device(0, dev.data_ptr(), nbytes, final=True)
        b = dm.wait(0)
    assert b.n_records == n and b.consumed == nbytes and b.status == 0 and b.scan_path == 1
    assert np.array_equal(b.rec_off, np.arange(n + 1, dtype=np.uint64) * w.rec_bytes)  # fixed-size records
    assert b.bin_count.sum() == n and np.array_equal(np.bincount(b.assign, minlength=97), b.bin_count)
    assert np.array_equal(np.cumsum(np.concatenate([[0], b.bin_count])), b.bin_start)
    # order is the stable counting sort of assign
    assert np.array_equal(b.order, np.argsort(b.assign, kind="stable"))
    # a slice against the oracle
    k = 50_000
    text = dev[: k * w.rec_bytes].cpu().numpy()
    assert np.array_equal(b.assign[:k], orc.demux(text, bcs, w.mismatch)["assign"])


# ---- reads with non-ACGT bases in the match window: the (m-1)-table shortcut and the table walk -----------------
@pytest.mark.parametrize("k,n_bc,m", [(6, 40, 1), (6, 40, 2), (8, 96, 1), (5, 30, 0), (10, 200, 2)])
def test_reads_with_invalid_bases_in_the_window(gpu, k, n_bc, m):
    rng = random.Random(k * 100 + n_bc + m)
    bcs = []
    while len(bcs) < n_bc:  # an adversarial table: near neighbours on purpose, no distance constraint
        if bcs and rng.random() < 0.5:
            b = bytearray(rng.choice(bcs))
            b[rng.randrange(k)] = rng.choice(b"ACGT")
            b = bytes(b)
        else:
            b = bytes(rng.choice(b"ACGT") for _ in range(k))
        if b not in bcs:
            bcs.append(b)
    out = bytearray()
    for i in range(6000):
        pre = bytearray(rng.choice(bcs))
        for _ in range(rng.choice((0, 0, 1, 1, 2, 3))):          # substitutions
            pre[rng.randrange(k)] = rng.choice(b"ACGT")
        for _ in range(rng.choice((0, 1, 1, 1, 2, 3))):          # invalid bases: N, lower case, IUPAC
            pre[rng.randrange(k)] = rng.choice(b"NnacgtRY")
        s = bytes(pre) + bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 30)))
        out += b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n"
    for flags in (0, gpu["ffi"].CFG_NO_LUT):
        _check_against_oracle(gpu, bytes(out), bcs, m, batch_bytes=1 << 20, flags=flags)


# ---- irregular records across many scan regions ------------------------------------------------------------------
@pytest.mark.parametrize("fmt,seed", [("fastq", 1), ("fastq", 2), ("fasta", 3)])
def test_irregular_records_many_regions(gpu, fmt, seed):
    """~25 MB in ONE batch (hundreds of 36 KB tiles, a region per CTA): read lengths 30-400, a sprinkle of CRLF records,
    '+id' separator lines, N / lower-case bases, multi-line FASTA -- region starts land in every kind of line"""
    rng = random.Random(seed)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(7)) for _ in range(60)]
    parts = []
    size = 0
    i = 0
    while size < 25_000_000:
        L = rng.randint(30, 400)
        pre = bytearray(rng.choice(bcs)) if rng.random() < 0.8 else bytearray(rng.choice(b"ACGT") for _ in range(7))
        if rng.random() < 0.3:
            pre[rng.randrange(7)] = rng.choice(b"ACGTNa")
        s = bytes(pre) + bytes(rng.choice(b"ACGT") for _ in range(L))
        odd = rng.random() < 0.01
        nl = b"\r\n" if odd and rng.random() < 0.5 else b"\n"
        if fmt == "fastq":
            q = bytes(rng.randint(33, 73) for _ in range(len(s)))
            sep = (b"+r%d" % i) if odd else b"+"
            rec = b"@r%d len=%d" % (i, L) + nl + s + nl + sep + nl + q + nl
        else:
            if odd:
                w = rng.randint(5, 60)
                body = nl.join(s[j:j + w] for j in range(0, len(s), w))
            else:
                body = s
            rec = b">r%d len=%d" % (i, L) + nl + body + nl
        parts.append(rec)
        size += len(rec)
        i += 1
    text = b"".join(parts)
    _check_against_oracle(gpu, text, bcs, 1, batch_bytes=32 << 20, chunk=4 << 20)


@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("width_delta", [-2, -1, 0, 1, 2])
def test_fasta_line_width_around_the_match_window(gpu, crlf, width_delta):
    """sequence lines exactly as wide as the match window, one less, one more ...: with CR LF the CR sits right where a
    base of the window would be"""
    rng = random.Random(100 + width_delta + (7 if crlf else 0))
    k = 9
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(k)) for _ in range(50)]
    nl = b"\r\n" if crlf else b"\n"
    w = k + width_delta
    out = bytearray()
    for i in range(30000):
        pre = bytearray(rng.choice(bcs))
        if rng.random() < 0.4:
            pre[rng.randrange(k)] = rng.choice(b"ACGTN")
        s = bytes(pre) + bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 40)))
        wi = w if rng.random() < 0.7 else rng.randint(3, 50)
        out += b">s%d" % i + nl + nl.join(s[j:j + wi] for j in range(0, len(s), wi)) + nl
    _check_against_oracle(gpu, bytes(out), bcs, 1, batch_bytes=4 << 20, chunk=1 << 20)


def test_fastq_many_odd_records(gpu):
    """every tenth record needs the general evaluation (CR LF, '+id', N): the out-of-line path next to the fast one"""
    rng = random.Random(77)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(8)) for _ in range(30)]
    out = bytearray()
    for i in range(60000):
        L = rng.randint(8, 200)
        pre = bytearray(rng.choice(bcs))
        if rng.random() < 0.3:
            pre[rng.randrange(8)] = rng.choice(b"ACGTNn")
        s = bytes(pre) + bytes(rng.choice(b"ACGT") for _ in range(L - 8))
        q = bytes(rng.randint(33, 73) for _ in range(L))
        odd = rng.random() < 0.1
        nl = b"\r\n" if odd and rng.random() < 0.5 else b"\n"
        out += b"@q%d" % i + nl + s + nl + ((b"+q%d" % i) if odd else b"+") + nl + q + nl
    _check_against_oracle(gpu, bytes(out), bcs, 2, batch_bytes=8 << 20, chunk=1 << 20)


def test_plain_c_caller(gpu):
    """tests/abi_smoke.c: a C99 program over the C ABI (built by the CPU test or here)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sabreur_b200", "bin", "abi_smoke")
    libdir = os.path.join(root, "sabreur_b200", "lib")
    src = os.path.join(root, "tests", "abi_smoke.c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["gcc", "-std=c99", "-o", exe, os.path.join(root, "tests", "abi_smoke.c"), "-L" + libdir,
                               "-lsabreur_gpu", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    # r1 ACGT exact, r2 ACGA one mismatch -> row 0; r3 TTTT -> unknown; r4 GGGG -> row 1
    assert out.stdout.strip() == "ok 4 2 1 1 | 0 0 2 1"
