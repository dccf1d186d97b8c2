"""GPU parity at BASELINE's batch sizes: one full 3,000,000-read cfg3 batch compared bit for bit with the C oracle,
and size-independent properties of the result on every batch of a larger resident shard (the domain's invariants:
counts add up, order is a stable permutation grouped by bin, assignments agree with the first-match rule)."""
import numpy as np
import pytest

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


def _first_match_numpy(prefixes: np.ndarray, table: np.ndarray, m: int) -> np.ndarray:
    """src/demux.rs:52-54 over an (n, k) byte matrix: first row with <= m byte mismatches, else n_barcodes"""
    n = prefixes.shape[0]
    out = np.full(n, table.shape[0], dtype=np.uint16)
    todo = np.ones(n, dtype=bool)
    for b in range(table.shape[0]):
        hit = todo & ((prefixes != table[b]).sum(1) <= m)
        out[hit] = b
        todo &= ~hit
    return out


def _check_properties(b, n, n_bins, rec):
    assert b.n_records == n and b.status == 0 and b.consumed == n * rec and b.scan_path == 1
    assert int(b.bin_count.sum()) == n
    assert np.array_equal(b.bin_start, np.concatenate([[0], np.cumsum(b.bin_count)]).astype(np.uint32))
    assert np.array_equal(b.rec_off, (np.arange(n + 1, dtype=np.uint64) * rec).astype(np.uint32))
    order = b.order.astype(np.int64)
    seen = np.zeros(n, dtype=bool)
    seen[order] = True
    assert seen.all()                                                    # a permutation
    assert np.array_equal(b.assign[order], np.repeat(np.arange(n_bins, dtype=np.uint16), b.bin_count))  # grouped by bin
    d = np.diff(order)
    inner = np.ones(n - 1, dtype=bool)
    inner[(b.bin_start[1:-1][b.bin_start[1:-1] < n] - 1).astype(np.int64)[b.bin_start[1:-1][b.bin_start[1:-1] < n] > 0]] = False
    assert (d[inner] > 0).all()                                          # stable: ascending inside every bin


@pytest.mark.parametrize("name,n_batch,n_batches", [("cfg3", 3_000_000, 2), ("cfg4", 1_000_000, 1), ("cfg5", 600_000, 1)])
def test_full_size_batches(gpu, name, n_batch, n_batches):
    torch, synth, dmx = gpu["torch"], gpu["synth"], gpu["demux"]
    w = synth.WORKLOADS[name]
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec = w.rec_bytes
    n = n_batch * n_batches
    dev = torch.empty(n * rec + 4096, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 10_000_000, n, dev.data_ptr())
    k = w.bc_len
    seq_off = 17 if w.fmt == 2 else 15  # header bytes of the synthetic records
    with dmx.Demux(bcs, w.mismatch, fmt=w.fmt, batch_bytes=n_batch * rec + 4096, max_records=n_batch + 16) as dm:
        for bi in range(n_batches):
            dm.demux_device(0, dev.data_ptr() + bi * n_batch * rec, n_batch * rec, final=True,
                            stream=torch.cuda.current_stream().cuda_stream)
            b = dm.wait(0)
            _check_properties(b, n_batch, len(bcs) + 1, rec)
            # every assignment against the first-match rule restated in numpy on the raw prefix bytes
            text = dev[bi * n_batch * rec:(bi + 1) * n_batch * rec].cpu().numpy().reshape(n_batch, rec)
            exp = _first_match_numpy(text[:, seq_off:seq_off + k], tab, w.mismatch)
            assert np.array_equal(b.assign, exp)
            if bi == 0:  # one whole batch through the C oracle: parser, assignments, counts, stable order
                o = orc.demux(text.reshape(-1), bcs, w.mismatch)
                assert np.array_equal(b.assign, o["assign"]) and np.array_equal(b.bin_count, o["bin_count"])
                assert np.array_equal(b.order, o["order"]) and np.array_equal(b.bin_start, o["bin_start"])
                assert np.array_equal(b.rec_off[:-1].astype(np.uint64), o["recs"]["rec_off"])


def test_maximum_batch_one_gib(gpu):
    """the largest batch the u32 offsets allow (1 GiB of text): the scan's region bookkeeping at its limits"""
    torch, synth, dmx = gpu["torch"], gpu["synth"], gpu["demux"]
    w = synth.CFG3
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec = w.rec_bytes
    n = ((1 << 30) - 4096) // rec
    dev = torch.empty(n * rec + 4096, dtype=torch.uint8, device="cuda")
    synth.reads_device(w, tab, 0, n, dev.data_ptr())
    with dmx.Demux(bcs, w.mismatch, fmt=w.fmt, batch_bytes=n * rec + 4096, max_records=n + 16) as dm:
        dm.demux_device(0, dev.data_ptr(), n * rec, final=True, stream=torch.cuda.current_stream().cuda_stream)
        b = dm.wait(0)
    _check_properties(b, n, len(bcs) + 1, rec)
    text = dev[: n * rec].cpu().numpy().reshape(n, rec)
    assert np.array_equal(b.assign, _first_match_numpy(text[:, 17:25], tab, w.mismatch))
