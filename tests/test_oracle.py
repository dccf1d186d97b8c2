"""CPU tests: pin the oracle against the reference's own known-answer tests, the goldens derived
from the in-tree fixtures and an independent pure-Python twin."""
import hashlib
import random

import numpy as np
import pytest

from conftest import fixture_bytes, read_barcode_table
from oracle import oracle as orc


# --- the reference's own bc_cmp truth table: /root/reference/src/utils.rs:177-205 ---------------
@pytest.mark.parametrize("bc,seq,m,want", [
    (b"ATCG", b"ATCGATCGATCG", 0, True),    # test_bc_cmp_ok
    (b"TGCA", b"ATCGATCGATCG", 0, False),   # test_bc_cmp_not_ok
    (b"AACG", b"ATCGATCGATCG", 1, True),    # test_bc_cmp_mismatch_ok
    (b"AACG", b"ATCGATCGATCG", 0, False),   # test_bc_cmp_mismatch_not_ok
])
def test_bc_cmp_reference_truth_table(bc, seq, m, want):
    assert orc.bc_cmp(bc, seq, m) is want
    assert orc.py_bc_cmp(bc, seq, m) is want


def test_bc_cmp_criterion_pair():
    # benches/my_benchmark.rs:22-32 compares ATCGATGTGC / ATCGATGTGA with mismatch = 1
    assert orc.bc_cmp(b"ATCGATGTGC", b"ATCGATGTGA", 1)
    assert not orc.bc_cmp(b"ATCGATGTGC", b"ATCGATGTGA", 0)


def test_bc_cmp_u8_wrap():
    # sum::<u8>() with overflow-checks off (Cargo.toml:29-35): 256 mismatches wrap to 0
    assert orc.bc_cmp(b"A" * 256, b"C" * 256, 0)
    assert not orc.bc_cmp(b"A" * 255, b"C" * 255, 0)
    assert orc.py_bc_cmp(b"A" * 256, b"C" * 256, 0)


def test_bc_cmp_zip_truncation():
    assert orc.bc_cmp(b"ACGTTTTT", b"ACG", 0)      # only 3 bytes compared
    assert orc.bc_cmp(b"XXX", b"ACGTACGT", 3)      # the sentinel matches anything at m >= 3 (H4)
    assert not orc.bc_cmp(b"XXX", b"ACGTACGT", 2)  # ... and never at m <= 2 (H3)


# --- derived goldens, SURVEY.md 8c ---------------------------------------------------------------
def _bcs(name):
    return [r[0].encode() for r in read_barcode_table(name)]


def test_cfg1_test_fq_bc_se_m0():
    text = fixture_bytes("test.fq")
    assert hashlib.sha256(text).hexdigest().startswith("4b0535c8a1fe434f")
    outs = orc.demux_outputs(text, _bcs("bc_se.txt"), 0)
    assert all(o == b"" for o in outs[:-1])   # ten empty barcode files
    assert outs[-1] == text                   # unknown.fa is byte-identical to tests/test.fq


@pytest.mark.parametrize("m", [0, 1, 2])
def test_test_fq_intest_barcodes(m):
    # src/demux.rs:236-238: in-test barcodes ACCGTA / ATTGTT on tests/test.fq(.gz)
    text = fixture_bytes("test.fq.gz")
    d = orc.demux(text, [b"ACCGTA", b"ATTGTT"], m)
    assert d["assign"].tolist() == [0, 1, 2]
    ids = [text[int(r["id_off"]):int(r["id_off"] + r["id_len"])] for r in d["recs"]]
    assert ids == [b"id desc", b"id2", b"id3"]


def test_test_fa_gz_m0_unknown():
    text = fixture_bytes("test.fa.gz")
    d = orc.demux(text, [b"ACCGTA"], 0)
    assert d["fmt"] == orc.FMT_FASTA and d["assign"].tolist() == [1]
    assert orc.seq(text, d["recs"][0], d["fmt"]) == b"ATCGATCGATCGATC"


R1_GOLD = [(82312, "90d9764b8b69432c"), (83351, "dd06fd51837741d1"), (83359, "98a7be33c3ac7170"),
           (83325, "5af261ae384d0e8d"), (83320, "cf6171f02f1974ce"), (83354, "8a8208327d1ba4cc"),
           (83302, "d88301f0531c93a4"), (83347, "9e490c81cc82b11e"), (83316, "88d17addc88d6666"),
           (83327, "5a32eafb431d020f")]
R2_GOLD = ["3686079bca97996c", "7fc5a66ffdfe6101", "8293662d5bd9b6b4", "98e65e4c22cb35df", "1cc04b3230073339",
           "d4e226dd2f2c9917", "5e5fc12f2bbaebba", "4a4c0bd82078af82", "87f8a9ac4c20bd39", "c7969e3ae26f42ac"]


def test_reads_1_bc_se_m0(reads_1):
    outs = orc.demux_outputs(reads_1, _bcs("bc_se.txt"), 0)
    got = [(len(o), hashlib.sha256(o).hexdigest()[:16]) for o in outs[:-1]]
    assert got == R1_GOLD
    assert outs[-1] == b""
    d = orc.demux(reads_1, _bcs("bc_se.txt"), 0)
    assert d["bin_count"].tolist() == [1000] + [999] * 9 + [0]


def test_reads_2_bc_pe_m0(reads_2):
    outs = orc.demux_outputs(reads_2, _bcs("bc_pe_fa.txt"), 0)
    assert [hashlib.sha256(o).hexdigest()[:16] for o in outs[:-1]] == R2_GOLD
    assert len(outs[0]) == 82312 and outs[-1] == b""


def test_reads_1_file_order_shadowing(reads_1):
    # m=1: row 4 (TGTACATG) shadows row 9 (TGTACAAG); m=2: row 6 (TGTCAGTA) also shadows row 10
    bcs = _bcs("bc_se.txt")
    c1 = orc.demux(reads_1, bcs, 1)["bin_count"].tolist()
    assert c1[3] == 1998 and c1[8] == 0
    c2 = orc.demux(reads_1, bcs, 2)["bin_count"].tolist()
    assert c2[3] == 1998 and c2[8] == 0 and c2[5] == 1998 and c2[9] == 0


# --- C oracle vs the independent Python twin -----------------------------------------------------
def _rand_fastq(rng, n, crlf=False, plus_id=False, final_nl=True):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i in range(n):
        L = rng.randint(6, 40)
        s = bytes(rng.choice(b"ACGTN") for _ in range(L))
        q = bytes(rng.randint(33, 73) for _ in range(L))
        out += b"@r%d x" % i + nl + s + nl + (b"+r%d" % i if plus_id else b"+") + nl + q + nl
    if not final_nl:
        out = out[:-len(nl)]
    return bytes(out)


def _rand_fasta(rng, n, width=0, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i in range(n):
        L = rng.randint(6, 70)
        s = bytes(rng.choice(b"ACGT") for _ in range(L))
        out += b">s%d" % i + nl
        if width:
            for j in range(0, L, width):
                out += s[j:j + width] + nl
        else:
            out += s + nl
    return bytes(out)


@pytest.mark.parametrize("kind", ["fq", "fq_crlf", "fq_plus", "fq_nofinal", "fa", "fa_w7", "fa_crlf"])
@pytest.mark.parametrize("m", [0, 1, 2])
def test_c_oracle_matches_python_twin(kind, m):
    rng = random.Random(hash(kind) & 0xFFFF)
    text = {
        "fq": lambda: _rand_fastq(rng, 60),
        "fq_crlf": lambda: _rand_fastq(rng, 60, crlf=True),
        "fq_plus": lambda: _rand_fastq(rng, 60, plus_id=True),
        "fq_nofinal": lambda: _rand_fastq(rng, 60, final_nl=False),
        "fa": lambda: _rand_fasta(rng, 60),
        "fa_w7": lambda: _rand_fasta(rng, 60, width=7),
        "fa_crlf": lambda: _rand_fasta(rng, 60, width=9, crlf=True),
    }[kind]()
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(4)) for _ in range(12)]
    asg, outs = orc.py_demux(text, bcs, m)
    d = orc.demux(text, bcs, m)
    assert d["assign"].tolist() == asg
    assert orc.demux_outputs(text, bcs, m) == outs
    # partition is a stable counting sort of the assignments
    order = d["order"]
    assert sorted(order.tolist()) == list(range(len(asg)))
    for b in range(len(bcs) + 1):
        seg = order[d["bin_start"][b]:d["bin_start"][b + 1]]
        assert all(asg[i] == b for i in seg) and list(seg) == sorted(seg)
    # the streaming (reference-faithful) loop agrees with the batch path
    got, cnt, sb, sh = orc.demux_stream(text, bcs, m, want_hash=True)
    assert got == len(asg) and cnt.tolist() == d["bin_count"].tolist()
    assert sb.tolist() == [len(o) for o in outs]
    assert [int(h) for h in sh] == [orc.fnv1a(o) if o else 0 for o in outs]


def test_parse_errors_and_tail():
    with pytest.raises(orc.OracleError) as e:
        orc.parse(b"@a\nACGT\n+\nIII\n")
    assert e.value.code == orc.E_UNEQUAL_LEN
    with pytest.raises(orc.OracleError) as e:
        orc.parse(b"@a\nACGT\n-\nIIII\n")
    assert e.value.code == orc.E_INVALID_SEP
    with pytest.raises(orc.OracleError) as e:
        orc.parse(b"@a\nACGT\n+\nIIII\nXb\nAC\n+\nII\n")
    assert e.value.code == orc.E_INVALID_START and e.value.record == 1
    with pytest.raises(orc.OracleError) as e:
        orc.parse(b"@a\nACGT\n+")
    assert e.value.code == orc.E_TRUNCATED
    with pytest.raises(orc.OracleError) as e:
        orc.parse(b"")
    assert e.value.code == orc.E_EMPTY
    with pytest.raises(orc.OracleError):
        orc.parse(b"ACGT\n")
    # up to three trailing blank lines are tolerated, four make a bogus record
    recs, _, consumed = orc.parse(b"@a\nACGT\n+\nIIII\n\n\n\n")
    assert len(recs) == 1 and consumed == 18
    with pytest.raises(orc.OracleError):
        orc.parse(b"@a\nACGT\n+\nIIII\n\n\n\n\n")
    # non-final batch: the incomplete tail is not consumed
    recs, _, consumed = orc.parse(b"@a\nACGT\n+\nIIII\n@b\nAC", final=False)
    assert len(recs) == 1 and consumed == 15
    recs, _, consumed = orc.parse(b">a\nACGT\n>b\nAC\n", final=False)
    assert len(recs) == 1 and consumed == 8
    # short read: the reference panics on the slice (src/demux.rs:54)
    with pytest.raises(orc.OracleError) as e:
        orc.demux(b"@a\nAC\n+\nII\n", [b"ACGT"], 0)
    assert e.value.code == orc.E_SHORT_READ


def test_ragged_barcode_lengths():
    # bc_len comes from row 0; a longer row is truncated by zip, a shorter row compares fewer bytes
    text = b">a\nACGTAAAA\n>b\nACCCAAAA\n"
    d = orc.demux(text, [b"TTTT", b"ACGTAC", b"AC"], 0)
    assert d["assign"].tolist() == [1, 2]


# --- the oracle-side synthetic generator (oracle/synth_ref.c) against the product's ------------
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg5"])
def test_synth_ref_equals_product_generator(name):
    """bench.py --impl reference builds its input with the oracle's restatement of the generator (so that the process
    timing the CPU baseline never maps libsabreur_gpu.so); it must be the same workload byte for byte.  sbr_synth_host
    and sbr_synth_table are plain host functions: no GPU is needed."""
    from sabreur_b200 import synth
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    ref_tab = orc.synth_table(w.cfg_id, w.n_barcodes, w.bc_len, w.min_dist)
    assert np.array_equal(tab, ref_tab)
    seed = synth.SEED_READS + w.cfg_id
    for mate in ((0, 1) if w.paired else (0,)):
        for first, n in ((0, 3000), (999_999_999_000, 700)):
            want = synth.reads_host(w, tab, first, n, mate)
            got = orc.synth_reads(w.fmt, w.read_len, mate, seed, ref_tab, first, n, threads=3 if first == 0 else 1)
            assert got.size == n * w.rec_bytes
            assert np.array_equal(got, want)


# --- committed digests of the synthetic workloads (tests/golden/synth_goldens.json, made by make_synth_goldens.py) -------
def _synth_goldens():
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return json.load(open(os.path.join(here, "golden", "synth_goldens.json")))


@pytest.mark.parametrize("g", _synth_goldens(), ids=lambda g: f"{g['workload']}-mate{g['mate']}")
def test_oracle_reproduces_the_committed_synthetic_goldens(g):
    """the oracle today against the digests committed when the workloads were defined: generator, parser, first-match
    rule and partition cannot drift (together with the GPU path, which is compared with the same file) unnoticed"""
    from sabreur_b200 import synth
    w = synth.WORKLOADS[g["workload"]]
    tab = orc.synth_table(w.cfg_id, w.n_barcodes, w.bc_len, w.min_dist)
    text = orc.synth_reads(w.fmt, w.read_len, g["mate"], synth.SEED_READS + w.cfg_id, tab, g["first"], g["n"])
    assert hashlib.sha256(text.tobytes()).hexdigest() == g["text_sha256"]
    assert hashlib.sha256(tab.tobytes()).hexdigest() == g["table_sha256"]
    exp = orc.demux(text, [bytes(r) for r in tab], g["mismatch"])
    assert hashlib.sha256(exp["assign"].astype("<u2").tobytes()).hexdigest() == g["assign_sha256"]
    assert hashlib.sha256(exp["order"].astype("<u4").tobytes()).hexdigest() == g["order_sha256"]
    assert int(exp["bin_count"][-1]) == g["unknown"] and [int(x) for x in exp["bin_count"][:8]] == g["bin_count_head"]
