import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIXTURES = os.path.join(ROOT, "tests", "golden", "fixtures")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fixture_bytes(name: str) -> bytes:
    p = os.path.join(FIXTURES, name)
    if name.endswith(".gz"):
        return gzip.open(p).read()
    return open(p, "rb").read()


def read_barcode_table(name: str):
    """rows of the tab-delimited barcode file (src/utils.rs:103-112)"""
    rows = []
    for line in open(os.path.join(FIXTURES, name)).read().splitlines():
        rows.append(line.split("\t"))
    return rows


@pytest.fixture(scope="session")
def reads_1():
    return fixture_bytes("reads_1.fa.gz")


@pytest.fixture(scope="session")
def reads_2():
    return fixture_bytes("reads_2.fa.gz")
