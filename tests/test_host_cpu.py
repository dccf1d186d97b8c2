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
