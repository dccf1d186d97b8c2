"""The N>1 host path with REAL GPU results: two ranks (one process each, `gloo` for the small merges; both on cuda:0 when
the box has a single GPU, on cuda:0 / cuda:1 otherwise) deal the batches of one stream round-robin, every rank runs its
batches through the C ABI, and rank 0 merges them with sabreur_b200/shard.py.  The merged partition must be the
sequential oracle's, including a parse error in the middle of the stream (everything behind it is dropped, as the
sequential reference never gets there)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(damage: bool):
    from sabreur_b200 import synth
    w = synth.CFG3
    tab = synth.table(w)
    per, nb = 20_000, 9
    text = synth.reads_host(w, tab, 123_000, per * nb).tobytes()
    if damage:  # record 5 of batch 4 loses its '+' line marker
        at = (4 * per + 5) * w.rec_bytes + 17 + 151
        text = text[:at] + b"-" + text[at + 1:]
    return w, tab, text, per, nb


def _worker(rank, world, port, damage, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from sabreur_b200 import demux, shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        w, tab, text, per, nb = _case(damage)
        bcs = [bytes(r) for r in tab]
        rec = w.rec_bytes
        dev = rank % torch.cuda.device_count()
        with demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=dev, batch_bytes=per * rec + 4096, n_slots=2, max_records=per + 16) as dm:
            def run_batch(seq, _):
                buf = dm.slot_buffer(0)
                buf[: per * rec] = np.frombuffer(text[seq * per * rec:(seq + 1) * per * rec], dtype=np.uint8)
                dm.submit(0, 0, per * rec, seq, True)
                b = dm.wait(0)
                return shard.BatchResult(int(b.seq), b.n_records, b.bin_count, b.bin_start, b.order, b.status, b.first_bad_record)

            mine = shard.run_sharded(list(range(nb)), world, rank, run_batch)
            n_bins = dm.n_bins
        allr = shard.gather_results(list(reversed(mine)), dst=0)
        if rank == 0:
            merged = shard.merge_in_order(allr, n_bins)
            good = text
            if damage:
                good = text[: (4 * per + 5) * rec]
                ok = merged.stopped_at == 4 * per + 5 and bool(merged.status & 0x02)
            else:
                ok = merged.stopped_at is None and merged.status == 0
            exp = orc.demux(good, bcs, w.mismatch)
            ok = ok and merged.n_records == len(exp["assign"]) and merged.bin_count.tolist() == exp["bin_count"].tolist()
            for b, lst in enumerate(merged.lists()):
                ok = ok and lst.tolist() == exp["order"][exp["bin_start"][b]:exp["bin_start"][b + 1]].tolist()
            q.put("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("damage", [False, True])
def test_two_ranks_real_gpu_results_merge_to_the_sequential_partition(damage):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    mp = torch.multiprocessing
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, damage, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"
