"""A fixed slice of the randomised parity hunt (tests/fuzz_gpu.py): random record shapes, line endings, tables, batch
sizes and kernel variants, with and without an injected violation, GPU against oracle.  The open-ended run is
`python tests/fuzz_gpu.py --seconds N` on the GPU box (1031 cases / 3.2 GB without a difference over its runs so far)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("first", [3_000_009, 3_000_021, 3_000_033])
def test_fuzz_slice(first):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (there is no CPU fallback to test)")
    import fuzz_gpu as fz
    fz._ffi.load()
    done = 0
    for seed in range(first, first + 12):
        text, bcs, m, batch, chunk, uf, params, starts = fz.make_case(seed)
        if len(text) > 4_000_000:  # the big shapes belong to the open-ended run
            continue
        why = (fz.check_violation(text, bcs, m, batch, chunk, uf, starts) if params["inject"]
               else fz.check(text, bcs, m, batch, chunk, uf))
        assert why is None, (params, why)
        done += 1
    assert done >= 4
