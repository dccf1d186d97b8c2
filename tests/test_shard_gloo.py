"""The N>1 host path on CPU: world_size-2 `gloo` process group.  Batches are dealt round-robin to two
ranks, each rank produces its batches' results (the oracle stands in for the GPU -- this test is about
the dealing/merging logic in sabreur_b200/shard.py, which never computes a match), counts are merged with
an all-reduce and the index lists are re-ordered on rank 0; the merged partition must equal the oracle's
sequential run over the whole stream."""
import os
import random
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sabreur_b200 import shard  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fastq(rng, n, bcs):
    out, cuts = bytearray(), [0]
    for i in range(n):
        L = rng.randint(12, 40)
        pre = bytearray(rng.choice(bcs)) if rng.random() < 0.8 else bytearray(rng.choice(b"ACGT") for _ in range(8))
        if rng.random() < 0.3:
            pre[rng.randrange(len(pre))] = rng.choice(b"ACGTN")
        s = bytes(pre) + bytes(rng.choice(b"ACGT") for _ in range(L))
        q = bytes(rng.randint(33, 73) for _ in range(len(s)))
        out += b"@r%d\n" % i + s + b"\n+\n" + q + b"\n"
        cuts.append(len(out))
    return bytes(out), cuts


def _make_case(seed=11, n=600, per_batch=37):
    rng = random.Random(seed)
    bcs = [bytes(rng.choice(b"ACGT") for _ in range(8)) for _ in range(12)]
    text, cuts = _fastq(rng, n, bcs)
    batches = [(cuts[i], cuts[min(i + per_batch, n)]) for i in range(0, n, per_batch)]  # record-aligned byte ranges
    return bcs, text, batches


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        bcs, text, batches = _make_case()
        m = 1

        def run_batch(seq, span):
            d = orc.demux(text[span[0]:span[1]], bcs, m, fmt=orc.FMT_FASTQ)
            return shard.BatchResult(seq=seq, n_records=len(d["assign"]), bin_count=d["bin_count"].astype(np.uint32),
                                     bin_start=d["bin_start"].astype(np.uint32), order=d["order"].astype(np.uint32))

        mine = shard.run_sharded(batches, world, rank, run_batch)
        assert [r.seq for r in mine] == list(shard.batches_of_rank(len(batches), world, rank))
        local = np.zeros(len(bcs) + 1, dtype=np.uint64)
        for r in mine:
            local += r.bin_count
        total = shard.merge_counts(local)
        # deliver out of order on purpose: reversed per rank
        allr = shard.gather_results(list(reversed(mine)), dst=0)
        if rank == 0:
            merged = shard.merge_in_order(allr, len(bcs) + 1)
            exp = orc.demux(text, bcs, m, fmt=orc.FMT_FASTQ)
            ok = (merged.n_records == len(exp["assign"]) and total.tolist() == exp["bin_count"].tolist()
                  and merged.bin_count.tolist() == exp["bin_count"].tolist())
            for b, lst in enumerate(merged.lists()):
                ok = ok and lst.tolist() == exp["order"][exp["bin_start"][b]:exp["bin_start"][b + 1]].tolist()
            q.put(("ok" if ok else "mismatch", int(total.sum())))
        else:
            assert allr is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_merge_equals_sequential_oracle():
    mp = pytest.importorskip("torch.multiprocessing")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    status, n = q.get(timeout=5)
    assert status == "ok" and n == 600


def test_shard_ranges_cover_the_index_space():
    for total, world in [(100_000_000, 8), (10, 3), (7, 8), (1_000_000_000, 4)]:
        spans = [shard.shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_reorder_buffer_and_error_truncation():
    rb = shard.ReorderBuffer()
    mk = lambda s, n, **kw: shard.BatchResult(seq=s, n_records=n, bin_count=np.array([n, 0], np.uint32),
                                              bin_start=np.array([0, n, n], np.uint32), order=np.arange(n, dtype=np.uint32), **kw)
    out = []
    for r in (mk(2, 3), mk(0, 2), mk(1, 4, status=0x02, first_bad_record=1)):
        out += [x.seq for x in rb.push(r)]
    assert out == [0, 1, 2] and rb.pending == 0
    with pytest.raises(ValueError):
        list(rb.push(mk(1, 1)))
    merged = shard.merge_in_order([mk(2, 3), mk(0, 2), mk(1, 4, status=0x02, first_bad_record=1)], 2)
    # the sequential reference stops at the bad record: batch 0 (2 records) + 1 record of batch 1, nothing of batch 2
    assert merged.n_records == 3 and merged.stopped_at == 3 and merged.lists()[0].tolist() == [0, 1, 2]
