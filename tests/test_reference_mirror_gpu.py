"""The reference's own tests of the hot path (src/demux.rs:135-297: test_se_demux_1 ... test_se_demux_m4), restated
against the Python mirror of its interface (sabreur_b200.demux.se_demux / pe_demux: same names, argument order and
return values).  The reference only asserts `.is_ok()`; here the outputs are checked too (SURVEY.md 8c goldens)."""
import gzip
import hashlib
import io
import os

import pytest

from conftest import FIXTURES, fixture_bytes, read_barcode_table

pytestmark = pytest.mark.gpu

GZIP, NO = ".gz", ""   # niffler::send::compression::Format::{Gzip, No} as the mirror spells them
LEVEL_ONE = 1          # niffler::Level::One


@pytest.fixture(scope="module")
def demux():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    from sabreur_b200 import demux as d
    return d


def _fx(name):
    return os.path.join(FIXTURES, name)


def _gunzip(buf: io.BytesIO) -> bytes:
    data = buf.getvalue()
    return gzip.decompress(data) if data else b""


def _bc_data(*barcodes, files=1):
    """Barcode = HashMap<&[u8], Vec<File>> with the "XXX" entry (tempfile() handles -> BytesIO)"""
    d = {b: [io.BytesIO() for _ in range(files)] for b in barcodes}
    d[b"XXX"] = [io.BytesIO() for _ in range(files)]
    return d


def test_se_demux_1(demux):
    bc_data, nb_records = _bc_data(b"ACCGTA"), {}
    res, is_unk_empty = demux.se_demux(_fx("test.fa.gz"), GZIP, LEVEL_ONE, bc_data, 0, nb_records)
    # the one 15-nt read ATCGATCGATCGATC does not start with ACCGTA: it goes to the unknown file
    assert res == {} and is_unk_empty is False
    assert _gunzip(bc_data[b"XXX"][0]) == fixture_bytes("test.fa.gz") and bc_data[b"ACCGTA"][0].getvalue() == b""


def test_se_demux_trim(demux):
    bc_data, nb_records = _bc_data(b"ACCGTA"), {}
    res, is_unk_empty = demux.se_demux(_fx("test.fa.gz"), GZIP, LEVEL_ONE, bc_data, 0, nb_records)
    assert res == {} and is_unk_empty is False


@pytest.mark.parametrize("name,m", [("test_se_demux_m1", 1), ("test_se_demux_m2", 2)])
def test_se_demux_fa_mismatches(demux, name, m):
    bc_data, nb_records = _bc_data(b"ACCGTA", b"ATTGTT"), {}
    res, is_unk_empty = demux.se_demux(_fx("test.fa.gz"), GZIP, LEVEL_ONE, bc_data, m, nb_records)
    # ATCGAT vs ACCGTA: 3 mismatches, vs ATTGTT: 2 mismatches -> unknown at m = 1, ATTGTT at m = 2
    if m == 1:
        assert res == {} and is_unk_empty is False
    else:
        assert res == {b"ATTGTT": 1} and is_unk_empty is True
        assert _gunzip(bc_data[b"ATTGTT"][0]) == fixture_bytes("test.fa.gz")


@pytest.mark.parametrize("name,m", [("test_se_demux_2", 0), ("test_se_demux_m3", 1), ("test_se_demux_m4", 2)])
def test_se_demux_fq(demux, name, m):
    bc_data, nb_records = _bc_data(b"ACCGTA", b"ATTGTT"), {}
    res, is_unk_empty = demux.se_demux(_fx("test.fq.gz"), GZIP, LEVEL_ONE, bc_data, m, nb_records)
    # `id desc` -> ACCGTA, `id2` -> ATTGTT, `id3` -> unknown (distances 4 and 5) at every m <= 2
    assert res == {b"ACCGTA": 1, b"ATTGTT": 1} and is_unk_empty is False
    text = fixture_bytes("test.fq.gz")
    recs = text.split(b"@id")[1:]
    assert _gunzip(bc_data[b"ACCGTA"][0]) == b"@id" + recs[0]
    assert _gunzip(bc_data[b"ATTGTT"][0]) == b"@id" + recs[1]
    assert _gunzip(bc_data[b"XXX"][0]) == b"@id" + recs[2]


def test_se_demux_errors_like_the_reference(demux):
    with pytest.raises(ValueError, match="Barcode data is empty"):                      # demux.rs:31-33
        demux.se_demux(_fx("test.fq.gz"), GZIP, LEVEL_ONE, {}, 0, {})
    with pytest.raises(ValueError, match="Missing 'XXX' fallback barcode entry"):        # demux.rs:43-46
        demux.se_demux(_fx("test.fq.gz"), GZIP, LEVEL_ONE, {b"ACCGTA": [io.BytesIO()]}, 0, {})


R1_GOLD = ["90d9764b8b69432c", "dd06fd51837741d1", "98a7be33c3ac7170", "5af261ae384d0e8d", "cf6171f02f1974ce",
           "8a8208327d1ba4cc", "d88301f0531c93a4", "9e490c81cc82b11e", "88d17addc88d6666", "5a32eafb431d020f"]
R2_GOLD = ["3686079bca97996c", "7fc5a66ffdfe6101", "8293662d5bd9b6b4", "98e65e4c22cb35df", "1cc04b3230073339",
           "d4e226dd2f2c9917", "5e5fc12f2bbaebba", "4a4c0bd82078af82", "87f8a9ac4c20bd39", "c7969e3ae26f42ac"]


def test_pe_demux_fixture(demux):
    """tests/reads_{1,2}.fa.gz x tests/bc_pe_fa.txt, --mismatch 0: format No keeps the forward file's gzip"""
    rows = read_barcode_table("bc_pe_fa.txt")
    bc_data = _bc_data(*[r[0].encode() for r in rows], files=2)
    res, unk = demux.pe_demux(_fx("reads_1.fa.gz"), _fx("reads_2.fa.gz"), NO, LEVEL_ONE, bc_data, 0, {})
    assert unk == "truetrue"
    assert res[b"GTCTGATG"] == 2000 and sorted(res.values())[:9] == [1998] * 9
    for i, r in enumerate(rows):
        f1, f2 = bc_data[r[0].encode()]
        assert hashlib.sha256(_gunzip(f1)).hexdigest()[:16] == R1_GOLD[i]
        assert hashlib.sha256(_gunzip(f2)).hexdigest()[:16] == R2_GOLD[i]


def test_pe_demux_refuses_a_mate_that_is_not_a_sequence_file(demux, tmp_path):
    """demux.rs:85-86: both parsers are built (`?`) before the first record is processed, so an empty or
    unrecognisable REVERSE mate fails the call before anything is written -- the C++ host and this mirror agree"""
    import gzip as gz
    rows = read_barcode_table("bc_pe_fa.txt")
    junk = tmp_path / "junk.fa"
    junk.write_bytes(b"not a sequence file\n")
    empty = tmp_path / "empty.fa.gz"
    empty.write_bytes(gz.compress(b""))
    for bad in (str(junk), str(empty)):
        for fwd, rev in ((_fx("reads_1.fa.gz"), bad), (bad, _fx("reads_2.fa.gz"))):
            bc_data = _bc_data(*[r[0].encode() for r in rows], files=2)
            with pytest.raises(demux.ParseError):
                demux.pe_demux(fwd, rev, NO, LEVEL_ONE, bc_data, 0, {})
            assert all(f.tell() == 0 for fs in bc_data.values() for f in fs)


# ---- src/utils.rs:177-205: the bc_cmp known-answer tests, asked of the GPU match kernel --------------------------
def _gpu_bc_cmp(demux, bc: bytes, seq: bytes, mismatch: int, **kw) -> bool:
    """bc_cmp(bc, seq, mismatch) as the hot loop uses it: does a read starting with `seq` go to barcode `bc`?"""
    text = b"@r\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n"
    got = demux.demux_bytes(text, [bc], mismatch, **kw)
    return int(got["assign"][0]) == 0


@pytest.mark.parametrize("flags", [0, 1, 2])  # first-match table, XOR+popc table walk, byte-exact kernel
def test_bc_cmp_known_answers(demux, flags):
    seq = b"ATCGATCGATCG"
    assert _gpu_bc_cmp(demux, b"ATCG", seq, 0, flags=flags)          # test_bc_cmp_ok
    assert not _gpu_bc_cmp(demux, b"TGCA", seq, 0, flags=flags)      # test_bc_cmp_not_ok
    assert _gpu_bc_cmp(demux, b"AACG", seq, 1, flags=flags)          # test_bc_cmp_mismatch_ok
    assert not _gpu_bc_cmp(demux, b"AACG", seq, 0, flags=flags)      # test_bc_cmp_mismatch_not_ok


def test_which_format(demux):
    assert demux.which_format(_fx("test.fa.gz")) == ".gz"            # utils.rs:281-298
    assert demux.which_format(_fx("reads_1.fa.bz2")) == ".bz2"
    assert demux.which_format(_fx("reads_1.fa.xz")) == ".xz"
    assert demux.which_format(_fx("test.fq.zst")) == ".zst"
