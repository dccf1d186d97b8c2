"""tests/kernel_tour.py -- a short tour of every kernel variant (two-stream pipeline with the plain and the lean match
kernel, one stream, separate launches, table walk, region-parallel / look-back scan, byte-exact match, a stream of lines too
dense for the region-parallel scan) with every result compared with the oracle -- as part of the GPU suite."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_tour_of_every_kernel_variant():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (there is no CPU fallback to test)")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kernel_tour
    kernel_tour.main()
