"""Multi-GPU host logic of the demux path: how batches are dealt to GPUs and how per-GPU results are
merged back into the reference's sequential semantics.

The path shards by batch: reads are independent (src/demux.rs:49-66 looks at one record at a time), so
batch `s` goes to GPU `s mod G`, every GPU runs scan -> match -> partition on its batches alone with its
own copy of the barcode table, and there is NO collective on the data path.  What has to be restored on
the host is the reference's observable order:

  * per-barcode counts (`nb_records`, src/demux.rs:58,104,117) are sums over batches        -> merge_counts
  * every output file holds its records in input order (append-mode writes, utils.rs:133-158) -> per-bin
    index lists are appended in batch-sequence order                                        -> ReorderBuffer

Nothing here computes a match: it only moves results.  One process per GPU (torchrun); the process group
(`nccl` on the GPU box, `gloo` in the CPU tests) is used for the small merges only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Iterator

import numpy as np


def owner_of_batch(batch_seq: int, world: int) -> int:
    """round-robin deal: batch s runs on rank s mod G"""
    return batch_seq % world


def batches_of_rank(n_batches: int, world: int, rank: int) -> range:
    return range(rank, n_batches, world)


def shard_range(total_units: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous [lo, hi) of the read index space for the kernel-only scaling runs: every rank generates
    its own shard on its device (bench.py); the first `total % world` ranks take one extra unit"""
    per, extra = divmod(total_units, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


@dataclass
class BatchResult:
    """what one GPU returns for one batch (sbr_result, host copies)"""
    seq: int
    n_records: int
    bin_count: np.ndarray   # u32 [n_bins]
    bin_start: np.ndarray   # u32 [n_bins + 1]
    order: np.ndarray       # u32 [n_records] batch-local record indices grouped by bin
    status: int = 0
    first_bad_record: int = 0


class ReorderBuffer:
    """Releases batch results in batch-sequence order, whatever order the GPUs finish in."""

    def __init__(self, first_seq: int = 0):
        self.next_seq = first_seq
        self._held: dict[int, BatchResult] = {}

    def push(self, r: BatchResult) -> Iterator[BatchResult]:
        if r.seq < self.next_seq or r.seq in self._held:
            raise ValueError(f"batch {r.seq} delivered twice")
        self._held[r.seq] = r
        while self.next_seq in self._held:
            yield self._held.pop(self.next_seq)
            self.next_seq += 1

    @property
    def pending(self) -> int:
        return len(self._held)


@dataclass
class MergedPartition:
    """stream-global view: counts per bin and, per bin, the global record indices in input order"""
    n_bins: int
    bin_count: np.ndarray = field(default=None)
    per_bin: list = field(default=None)
    n_records: int = 0
    status: int = 0
    stopped_at: int | None = None  # global index of the first bad record (reference stops there)

    def __post_init__(self):
        self.bin_count = np.zeros(self.n_bins, dtype=np.uint64)
        self.per_bin = [[] for _ in range(self.n_bins)]

    def append(self, r: BatchResult, error_mask: int = 0x1F):
        """append the next batch IN SEQUENCE; records at or after a parse error are dropped, as the
        sequential reference never reaches them (src/demux.rs:50 / 101,114)"""
        if self.stopped_at is not None:
            return
        nok = r.n_records
        if r.status & error_mask:
            nok = min(nok, r.first_bad_record)
            self.stopped_at = self.n_records + nok
        self.status |= r.status
        for b in range(self.n_bins):
            seg = r.order[r.bin_start[b]: r.bin_start[b + 1]]
            if nok < r.n_records:
                seg = seg[seg < nok]
            if seg.size:
                self.per_bin[b].append(seg.astype(np.uint64) + self.n_records)
                self.bin_count[b] += seg.size
        self.n_records += nok

    def lists(self) -> list[np.ndarray]:
        return [np.concatenate(x) if x else np.zeros(0, np.uint64) for x in self.per_bin]


def merge_counts(local_counts: np.ndarray, group=None) -> np.ndarray:
    """sum of per-rank per-bin counts (the only reduction of the path; it is a few KB and runs on the
    process group the launcher created: nccl on GPUs, gloo on CPU).  world 1: identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local_counts, dtype=np.uint64).copy()
    t = torch.from_numpy(np.asarray(local_counts, dtype=np.int64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(np.uint64)


def _pack(results: list[BatchResult]) -> np.ndarray:
    """one flat int64 vector for a rank's batch results: [count] + per result [seq, n_records, status, first_bad_record,
    n_bins] + bin_count + bin_start + order"""
    parts = [np.array([len(results)], dtype=np.int64)]
    for r in results:
        parts.append(np.array([r.seq, r.n_records, r.status, r.first_bad_record, len(r.bin_count)], dtype=np.int64))
        parts.append(np.asarray(r.bin_count, dtype=np.int64))
        parts.append(np.asarray(r.bin_start, dtype=np.int64))
        parts.append(np.asarray(r.order[: r.n_records], dtype=np.int64))
    return np.concatenate(parts)


def _unpack(flat: np.ndarray) -> list[BatchResult]:
    out, at = [], 1
    for _ in range(int(flat[0])):
        seq, n, status, bad, nb = (int(x) for x in flat[at: at + 5])
        at += 5
        bin_count = flat[at: at + nb].astype(np.uint32)
        at += nb
        bin_start = flat[at: at + nb + 1].astype(np.uint32)
        at += nb + 1
        order = flat[at: at + n].astype(np.uint32)
        at += n
        out.append(BatchResult(seq, n, bin_count, bin_start, order, status, bad))
    return out


def gather_results(local: list[BatchResult], dst: int = 0, group=None) -> list[BatchResult] | None:
    """bring every rank's batch results to `dst` (index lists go to the rank that owns the writers).  Plain tensors on the
    launcher's process group -- sizes first, then the padded payloads -- so that it runs on nccl and gloo alike and nothing
    is pickled"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    flat = _pack(local)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([flat.size], dtype=torch.int64, device=dev), group=group)
    longest = max(int(x.item()) for x in sizes)
    mine = torch.zeros(longest, dtype=torch.int64, device=dev)
    mine[: flat.size] = torch.from_numpy(flat).to(dev)
    everyone = [torch.empty(longest, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(everyone, mine, group=group)
    if rank != dst:
        return None
    out = []
    for size, t in zip(sizes, everyone):
        out += _unpack(t[: int(size.item())].cpu().numpy())
    return out


def run_sharded(batches: list, world: int, rank: int, run_batch: Callable[[int, object], BatchResult]) -> list[BatchResult]:
    """this rank's share of the deal: run_batch(seq, batch) for every batch it owns, in sequence order"""
    return [run_batch(s, batches[s]) for s in batches_of_rank(len(batches), world, rank)]


def merge_in_order(results: list[BatchResult], n_bins: int) -> MergedPartition:
    """results from all ranks in arrival order -> the sequential reference's partition"""
    rb, out = ReorderBuffer(), MergedPartition(n_bins)
    for r in results:
        for ready in rb.push(r):
            out.append(ready)
    if rb.pending:
        raise ValueError(f"{rb.pending} batches are missing their predecessors")
    return out
