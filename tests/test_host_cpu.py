"""CPU tests of the C++ host (sabreur_b200/csrc/host): codecs against Python's own gzip/bz2/lzma, the format
sniffer against the reference's which_format fixtures (src/utils.rs:281-298), and the command line's argument /
exit-code behaviour (src/cli.rs, src/main.rs:49-56).  No GPU is needed: the demux itself must FAIL here."""
import bz2
import gzip
import hashlib
import lzma
import os
import subprocess

import pytest

from conftest import FIXTURES, ROOT

BIN = os.path.join(ROOT, "sabreur_b200", "bin")
SABREUR = os.path.join(BIN, "sabreur")
SELFTEST = os.path.join(BIN, "codec_selftest")


@pytest.fixture(scope="module", autouse=True)
def built():
    from sabreur_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sabreur_b200", "csrc", "host")])
    assert os.path.exists(SABREUR) and os.path.exists(SELFTEST)


def _fx(name):
    return os.path.join(FIXTURES, name)


@pytest.mark.parametrize("name,ext", [("reads_1.fa.gz", ".gz"), ("reads_1.fa.bz2", ".bz2"), ("reads_1.fa.xz", ".xz"),
                                      ("test.fq.zst", ".zst"), ("test.fq", "")])
def test_which_format(name, ext):
    out = subprocess.run([SELFTEST, "sniff", _fx(name)], capture_output=True, text=True, check=True).stdout
    assert out.rstrip("\n") == ext


@pytest.mark.parametrize("name", ["reads_1.fa.gz", "reads_1.fa.bz2", "reads_1.fa.xz"])
@pytest.mark.parametrize("chunk", [1, 4097, 65536])
def test_reader_decodes_like_python(name, chunk):
    opener = {".gz": gzip.open, ".bz2": bz2.open, ".xz": lzma.open}[os.path.splitext(name)[1]]
    want = opener(_fx(name)).read()
    if chunk == 1 and len(want) > 200000:
        want_n = 20000  # byte-at-a-time over the whole file is slow: the head is enough
        got = subprocess.run(f"{SELFTEST} cat {_fx(name)} 1 | head -c {want_n}", shell=True, capture_output=True).stdout
        assert got == want[:want_n]
        return
    got = subprocess.run([SELFTEST, "cat", _fx(name), str(chunk)], capture_output=True, check=True).stdout
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()


def test_zstd_fixture_decodes_to_the_plain_fixture():
    got = subprocess.run([SELFTEST, "cat", _fx("test.fq.zst")], capture_output=True, check=True).stdout
    assert got == open(_fx("test.fq"), "rb").read()


@pytest.mark.parametrize("fmt,opener", [("gz", gzip.open), ("bz2", bz2.open), ("xz", lzma.open)])
def test_members_concatenate_into_a_valid_file(tmp_path, fmt, opener):
    data = gzip.open(_fx("reads_1.fa.gz")).read()
    p = tmp_path / f"out.{fmt}"
    with open(p, "wb") as fh:
        subprocess.run([SELFTEST, "pack", fmt, "1", "70001"], input=data, stdout=fh, check=True)
    assert opener(p).read() == data                      # an independent decoder reads every member
    back = subprocess.run([SELFTEST, "cat", str(p)], capture_output=True, check=True).stdout
    assert back == data                                  # and so does the host's own reader


def test_zstd_members_roundtrip(tmp_path):
    data = gzip.open(_fx("reads_1.fa.gz")).read()
    p = tmp_path / "out.zst"
    with open(p, "wb") as fh:
        subprocess.run([SELFTEST, "pack", "zst", "3", "50000"], input=data, stdout=fh, check=True)
    assert subprocess.run([SELFTEST, "cat", str(p)], capture_output=True, check=True).stdout == data


def _run(args, cwd):
    return subprocess.run([SABREUR] + args, capture_output=True, text=True, cwd=cwd)


def test_cli_argument_errors(tmp_path):
    r = _run([], tmp_path)
    assert r.returncode == 2 and "required arguments" in r.stderr
    r = _run([_fx("bc_se.txt"), str(tmp_path / "nope.fq")], tmp_path)
    assert r.returncode == 2 and "path does not exist" in r.stderr
    r = _run([_fx("bc_se.txt"), _fx("test.fq"), "-f", "rar"], tmp_path)
    assert r.returncode == 2 and "possible values" in r.stderr
    r = _run([_fx("bc_se.txt"), _fx("test.fq"), "-m", "300"], tmp_path)
    assert r.returncode == 2
    assert _run(["--help"], tmp_path).returncode == 0 and _run(["-V"], tmp_path).stdout.startswith("sabreur 0.7")


def test_cli_refuses_existing_output_dir_with_exit_73(tmp_path):
    (tmp_path / "out").mkdir()
    r = _run([_fx("bc_se.txt"), _fx("test.fq"), "-o", "out"], tmp_path)
    assert r.returncode == 73 and "already exists. Use --force to overwrite." in r.stdout
    log = (tmp_path / "sabreur.log").read_text()
    assert "[sabreur][INFO] sabreur v0.7 starting up in single-end mode" in log and "[sabreur][ERROR]" in log


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run([_fx("bc_se.txt"), _fx("test.fq"), "-o", "out"], tmp_path)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_cli_barcode_file_errors(tmp_path):
    """utils.rs:103-112 (no tab -> error, exit 1) and main.rs:148 (a row without enough columns panics, exit 101);
    both happen before any GPU work"""
    bad = tmp_path / "notab.txt"
    bad.write_text("ACGT bc1.fq\n")
    r = _run([str(bad), _fx("test.fq"), "-o", "o1"], tmp_path)
    assert r.returncode == 1 and "Input string is not tab-delimited" in r.stderr
    blank = tmp_path / "blank.txt"
    blank.write_text("ACGT\tbc1.fq\n\nCCGT\tbc2.fq\n")
    r = _run([str(blank), _fx("test.fq"), "-o", "o2"], tmp_path)
    assert r.returncode == 101 and "panicked" in r.stderr
    two = tmp_path / "two.txt"
    two.write_text("ACGT\tbc1_R1.fq\n")
    r = _run([str(two), _fx("test.fq"), _fx("test.fq"), "-o", "o3"], tmp_path)  # paired-end needs three columns
    assert r.returncode == 101


# ---- pure pieces of the host pipeline (host_selftest): newline budget, guessed cuts, canonical re-serialisation --------
HOSTTEST = os.path.join(BIN, "host_selftest")


def _host(args, data: bytes) -> bytes:
    return subprocess.run([HOSTTEST] + [str(a) for a in args], input=data, capture_output=True, check=True).stdout


@pytest.mark.parametrize("n", [0, 1, 15, 16, 63, 64, 65, 4096, 64 * 255, 64 * 255 + 1, 64 * 256 + 17, 1_000_003])
def test_host_newline_count(n):
    import random
    rng = random.Random(n)
    data = bytes(rng.choice(b"\nACGT\n@+I\r") for _ in range(min(n, 70000)))
    data = (data * (n // max(1, len(data)) + 1))[:n]
    assert int(_host(["count"], data)) == data.count(b"\n")
    assert int(_host(["count"], b"\n" * n)) == n  # every byte counter saturating at once


def test_host_guessed_cut_is_a_record_start():
    """the guess may be wrong by design (the GPU verifies it), but on well-formed input -- '@' and '+' opening quality
    lines included -- it must land on a record start inside the window, and never past the data"""
    import random
    rng = random.Random(3)
    for fastq in (True, False):
        out, starts = bytearray(), []
        for i in range(400):
            L = rng.randint(5, 120)
            s = bytes(rng.choices(b"ACGT", k=L))
            starts.append(len(out))
            if fastq:
                q = bytearray(rng.choices(range(33, 74), k=L))
                q[0] = rng.choice(b"@+I")
                out += b"@r%d\n" % i + s + b"\n+\n" + bytes(q) + b"\n"
            else:
                out += b">r%d\n" % i + s[: L // 2] + b"\n" + s[L // 2:] + b"\n"
        text = bytes(out)
        for end in [len(text)] + [rng.randrange(2000, len(text)) for _ in range(60)]:
            cut = int(_host(["cut", 2 if fastq else 1, 1500], text[:end]))
            assert cut <= end
            if cut < end:
                assert cut in starts and cut >= end - 1500, (fastq, end, cut)
    assert int(_host(["cut", 2, 100], b"ACGT" * 100)) == 400  # no line start in sight: everything stays in the batch


def test_host_canonical_record_equals_the_oracle_serialiser():
    """CRLF line ends, '+id' separator lines, multi-line FASTA, a missing final newline: the host's re-serialisation of a
    flagged record == the oracle's restatement of needletail's write_fastq / write_fasta (src/utils.rs:142-154)"""
    import random
    from oracle import oracle as orc
    rng = random.Random(11)
    for fastq in (True, False):
        for _case in range(60):
            nl = rng.choice((b"\n", b"\r\n"))
            L = rng.randint(4, 90)
            s = bytes(rng.choices(b"ACGTN", k=L))
            hdr = b"r%d some description" % _case
            if fastq:
                plus = rng.choice((b"+", b"+" + hdr))
                rec = b"@" + hdr + nl + s + nl + plus + nl + bytes(rng.choices(range(33, 74), k=L)) + nl
            else:
                w = rng.choice((L, 7, 60, 3))
                rec = b">" + hdr + nl + b"".join(s[j:j + w] + nl for j in range(0, L, w))
            if rng.random() < 0.3:
                rec = rec[:-len(nl)]  # the last record of a stream without a final newline
            recs, fmt, _ = orc.parse(rec, final=True)
            assert len(recs) == 1
            assert _host(["canon", fmt], rec) == orc.serialize(rec, recs[0], fmt), rec


# ---- BGZF (bgzip) input: blocks are inflated in parallel, the result must equal the sequential decoders' ---------------
def _bgzf(data: bytes, block: int = 65280, level: int = 6, eof: bool = True, extra_first: bytes = b"") -> bytes:
    """what bgzip writes: gzip members of <= 64 KiB with a 'BC' extra subfield holding the member's size - 1"""
    import struct
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof else [])
    for c in chunks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = co.compress(c) + co.flush()
        xlen = len(extra_first) + 6
        bsize = 12 + xlen + len(raw) + 8 - 1
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", xlen) + extra_first
        out += b"BC\x02\x00" + struct.pack("<H", bsize) + raw + struct.pack("<II", zlib.crc32(c), len(c))
    return bytes(out)


def _cat(path, chunk=None, threads=None, ok=True):
    env = dict(os.environ)
    if threads is not None:
        env["SABREUR_INFLATE_THREADS"] = str(threads)
    r = subprocess.run([SELFTEST, "cat", str(path)] + ([str(chunk)] if chunk else []), capture_output=True, env=env)
    assert (r.returncode == 0) == ok, r.stderr
    return r.stdout


@pytest.mark.parametrize("size", [0, 1, 65279, 65280, 65281, 1_000_000, 40_000_000])
def test_bgzf_input_decodes_in_parallel(tmp_path, size):
    import random
    rng = random.Random(size)
    line = bytes(rng.choices(b"ACGT", k=150))
    data = (b"@r\n" + line + b"\n+\n" + bytes(rng.choices(range(33, 74), k=150)) + b"\n") * (size // 306 + 1)
    data = data[:size]
    p = tmp_path / "in.fq.gz"
    p.write_bytes(_bgzf(data))
    assert gzip.decompress(p.read_bytes()) == data  # the writer above is a valid gzip file
    for threads, chunk in ((None, None), (4, 1000), (1, None)):  # 1 thread = the sequential decoder on the same file
        assert _cat(p, chunk, threads) == data


def test_bgzf_variants_and_mixed_files(tmp_path):
    data = b"".join(b">s%d\nACGTACGTAC\n" % i for i in range(30000))
    p = tmp_path / "x.gz"
    # the 'BC' subfield behind another one; no EOF marker; an ordinary gzip member appended; BGZF after an ordinary member
    p.write_bytes(_bgzf(data, extra_first=b"XY\x03\x00abc"))
    assert _cat(p) == data
    p.write_bytes(_bgzf(data, eof=False))
    assert _cat(p) == data
    p.write_bytes(_bgzf(data, block=10000) + gzip.compress(b"tail1\n") + _bgzf(b"tail2\n"))
    assert _cat(p) == data + b"tail1\ntail2\n"
    p.write_bytes(gzip.compress(b"head\n") + _bgzf(data))
    assert _cat(p) == b"head\n" + data


def test_bgzf_corruption_is_an_error(tmp_path):
    data = b"".join(b">s%d\nACGTACGTAC\n" % i for i in range(30000))
    good = _bgzf(data)
    p = tmp_path / "bad.gz"
    p.write_bytes(good[: len(good) // 2])                       # cut inside a block
    assert b"truncated" in subprocess.run([SELFTEST, "cat", str(p)], capture_output=True).stderr
    bad = bytearray(good)
    bad[len(good) // 3] ^= 0x55                                 # a flipped byte inside some block's deflate data
    p.write_bytes(bytes(bad))
    r = subprocess.run([SELFTEST, "cat", str(p)], capture_output=True)
    assert r.returncode != 0 and b"corrupt" in r.stderr


# ---- round 2: block-parallel decompression of splittable inputs (codec_par.cpp) ----------------------------------
def _text(n_rec=60000, seed=5):
    import random
    rng = random.Random(seed)
    out = bytearray()
    for i in range(n_rec):
        L = rng.randint(40, 120)
        out += b"@r%d\n" % i + bytes(rng.choices(b"ACGTN", k=L)) + b"\n+\n" + bytes(rng.choices(range(33, 74), k=L)) + b"\n"
    return bytes(out)


def _cat_par(path, threads, chunk=None):
    env = dict(os.environ, SABREUR_INFLATE_THREADS=str(threads))
    args = [SELFTEST, "cat", str(path)] + ([str(chunk)] if chunk else [])
    return subprocess.run(args, capture_output=True, env=env)


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_bzip2_blocks_decode_in_parallel(tmp_path, threads):
    """multi-block streams (every block found by its 48-bit magic at any bit offset and decoded as a stream of its own),
    concatenated streams with different levels, an empty stream in between, runs that expand past the block size"""
    text = _text()
    runs = b"A" * 3_000_000 + text[:100_000] + b"\0" * 1_500_000  # RLE1: a 100 k block holds far more than 100 kB of these
    cases = {
        "level1.bz2": bz2.compress(text, 1),            # ~70 blocks of 100 kB
        "level9.bz2": bz2.compress(text, 9),
        "concat.bz2": bz2.compress(text[:1_000_003], 2) + bz2.compress(b"", 9) + bz2.compress(text[1_000_003:], 4),
        "runs.bz2": bz2.compress(runs, 1),
    }
    want = {"level1.bz2": text, "level9.bz2": text, "concat.bz2": text, "runs.bz2": runs}
    for name, blob in cases.items():
        p = tmp_path / name
        p.write_bytes(blob)
        r = _cat_par(p, threads, 50_001)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha256(r.stdout).hexdigest() == hashlib.sha256(want[name]).hexdigest(), name


def test_bzip2_damaged_block_is_an_error_not_garbage(tmp_path):
    blob = bytearray(bz2.compress(_text(20000), 1))
    blob[len(blob) // 2] ^= 0x55
    p = tmp_path / "bad.bz2"
    p.write_bytes(bytes(blob))
    for threads in (1, 4):
        r = _cat_par(p, threads)
        assert r.returncode != 0 and b"bzip2" in r.stderr


@pytest.mark.parametrize("threads", [1, 4])
def test_xz_blocks_and_streams_decode_in_parallel(tmp_path, threads):
    """xz -T style multi-block streams (block sizes come from the stream's index, read from the end of the file
    backwards), concatenated streams, stream padding, different integrity checks"""
    import shutil
    text = _text()
    cases = {}
    cases["streams.xz"] = b"".join(lzma.compress(text[i:i + 700_001], preset=0) for i in range(0, len(text), 700_001))
    cases["padded.xz"] = lzma.compress(text[:2_000_000], preset=0) + b"\0" * 8 + lzma.compress(text[2_000_000:], preset=1,
                                                                                               check=lzma.CHECK_SHA256) + b"\0" * 4
    cases["crc32.xz"] = lzma.compress(text[:999], check=lzma.CHECK_CRC32) + lzma.compress(text[999:], check=lzma.CHECK_NONE)
    if shutil.which("xz"):
        plain = tmp_path / "plain.txt"
        plain.write_bytes(text)
        cases["blocks.xz"] = subprocess.run(["xz", "-0", "-T3", "--block-size=524288", "-c", str(plain)], capture_output=True, check=True).stdout
    for name, blob in cases.items():
        p = tmp_path / name
        p.write_bytes(blob)
        if name != "padded.xz":  # (Python's one-shot decoder stops at stream padding; the format allows it)
            assert lzma.decompress(blob) == text  # the case itself is a valid file
        r = _cat_par(p, threads, 65_537)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha256(r.stdout).hexdigest() == hashlib.sha256(text).hexdigest(), name


@pytest.mark.parametrize("threads", [1, 4])
def test_zstd_frames_decode_in_parallel(tmp_path, threads):
    text = _text()
    for member in (65_536, 4 << 20):
        p = tmp_path / f"m{member}.zst"
        with open(p, "wb") as fh:
            subprocess.run([SELFTEST, "pack", "zst", "2", str(member)], input=text, stdout=fh, check=True)
        r = _cat_par(p, threads, 100_003)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha256(r.stdout).hexdigest() == hashlib.sha256(text).hexdigest()
    # a file cut inside a frame is an error under both readers (frames written by ZSTD_compress carry no checksum, so
    # flipped payload bits are not something either reader could notice)
    blob = open(tmp_path / "m65536.zst", "rb").read()
    p = tmp_path / "cut.zst"
    p.write_bytes(blob[: len(blob) * 2 // 3])
    r = _cat_par(p, threads)
    assert r.returncode != 0 and text.startswith(r.stdout) and len(r.stdout) > len(text) // 2
