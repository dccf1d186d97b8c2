#!/usr/bin/env python
"""bench.py -- reads/s demuxed (record scan + barcode match + partition) on B200, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg4|cfg5] [--also cfg4,cfg5|none]
                  [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one GPU's shard of a synthetic workload, resident in HBM (`value`), or
streamed from pinned host memory through sbr_submit/sbr_wait (`e2e`).  Shards are independent (reads partition by
batch; no collective on the data path): weak scaling.  The line's top level is the primary workload (cfg3, the
configuration BASELINE.json quotes for one GPU); `workloads` holds the same measurements for the other synthetic
configurations (cfg4: the one BASELINE shards over 1/2/4/8 GPUs, cfg5), taken in the same process at the same N.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/s demuxed (match+partition)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5"], help="the primary workload (top level)")
    ap.add_argument("--also", default=None, help="comma list of further workloads for the `workloads` block "
                                                 "(default: cfg4,cfg5 when the primary is cfg3; 'none' = skip)")
    ap.add_argument("--units-per-gpu", type=int, default=0, help="reads (SE) / pairs (PE) resident per GPU (primary; "
                                                                 "the other workloads are scaled by the same factor)")
    ap.add_argument("--batch-units", type=int, default=0, help="reads per kernel batch (multiple of 16)")
    ap.add_argument("--e2e-batch-mb", type=int, default=256)
    ap.add_argument("--e2e-slots", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-lut", action="store_true", help="plain XOR+popc table walk in the match kernel")
    ap.add_argument("--no-overlap", action="store_true", help="match + partition on the caller's stream, behind the scan "
                                                              "(default: on the slot's own stream, beside the next scan)")
    ap.add_argument("--no-fuse", action="store_true", help="match and partition as separate launches")
    ap.add_argument("--slots", type=int, default=3, help="device workspaces the HBM-resident loop rotates through")
    ap.add_argument("--no-verify-merge", action="store_true", help="skip the multi-GPU merge check (sabreur_b200/shard.py)")
    return ap.parse_args()


def default_units(name: str) -> int:
    # cfg3: the whole 100 M-read configuration fits one GPU (32.1 GB).  cfg4: 1 B pairs / 8 GPUs = 125 M
    # pairs (80 GB) per GPU.  cfg5: 250 M reads (66.5 GB) per GPU.
    return {"cfg3": 100_000_000, "cfg4": 125_000_000, "cfg5": 250_000_000}[name]


# ---------------------------------------------------------------------------------------------------
# clocks sampler: NVML polled every 5 ms from a thread (the same counters as the nvidia-smi line of
# B200_PROFILING.md; nvidia-smi's own loop cannot sample a region of a few dozen milliseconds)
# ---------------------------------------------------------------------------------------------------
class Clocks:
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, torch_device_index: int):
        self.samples = []  # (t, sm_mhz, reasons bitmask, power W)
        self.max_mhz = None
        self.src = None
        self._stop = threading.Event()
        self._th = None
        self._nv = None
        self._h = None
        self._smi = None
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            p = torch.cuda.get_device_properties(torch_device_index)
            bdf = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            self._h = nv.nvmlDeviceGetHandleByPciBusId(bdf.encode() if hasattr(bdf, "encode") else bdf)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            self._nv = nv
            self.src = "nvml, 5 ms period"
        except Exception as e:  # no NVML: fall back to the recipe's nvidia-smi loop
            self._why = str(e)[:60]
            self._idx = torch_device_index

    def start(self):
        if self._nv is not None:
            self._th = threading.Thread(target=self._poll, daemon=True)
            self._th.start()
            return
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self._smi = subprocess.Popen(["nvidia-smi", "-i", str(self._idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi -lms 20"
            self._th = threading.Thread(target=self._pump, daemon=True)
            self._th.start()
        except OSError:
            self._smi = None

    def _poll(self):
        nv, h = self._nv, self._h
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = reasons(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.perf_counter(), float(mhz), int(r), pw))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for ln in self._smi.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.max_mhz = mx
            bits = 0
            for (bit, _), v in zip(self.REASONS, f[5:9]):
                if v.lower().startswith("active"):
                    bits |= bit
            try:
                pw = float(f[3])
            except ValueError:
                pw = None
            self.samples.append((time.perf_counter(), mhz, bits, pw))

    def stop(self):
        self._stop.set()
        if self._smi:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=2)
            except Exception:
                self._smi.kill()
        if self._th:
            self._th.join(timeout=2)

    def summary(self, t0: float, t1: float) -> dict:
        """the samples taken inside [t0, t1] (perf_counter): the region the GPU was under the timed load"""
        if self.src is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML / nvidia-smi: " + getattr(self, "_why", "")]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        sm = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        pw = [s[3] for s in inside if s[3] is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(inside), "samples_total": len(self.samples), "power_w_max": max(pw) if pw else None,
                "reasons": [name for bit, name in self.REASONS if bits & bit], "source": self.src}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's reference-faithful loop (one record at a time, byte-wise bc_cmp, first
# match in table order, canonical bytes appended to per-bin sinks).  The reference is single-threaded
# (no thread/rayon crate: Cargo.toml:14-22), so the primary figure uses ONE core; an all-cores figure
# (the same loop on disjoint shards, one per hardware thread) is reported beside it for context.
# ---------------------------------------------------------------------------------------------------
def cpu_reference_parallel(w, tab, sample_text, rec_bytes, threads: int):
    """the reference loop on `threads` disjoint shards at once -> (reads done, seconds, reads/s)"""
    from oracle import oracle as orc
    bcs = [bytes(r) for r in tab]
    n = sample_text.size // rec_bytes
    per = n // threads
    parts = [sample_text[i * per * rec_bytes:(i + 1) * per * rec_bytes] for i in range(threads)]
    res = [0] * threads

    def work(i):
        res[i] = orc.demux_stream(parts[i], bcs, w.mismatch)[0]  # ctypes releases the GIL

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return sum(res), dt, sum(res) / dt


def cpu_reference(w, tab, sample_text, rec_bytes, threads: int):
    from oracle import oracle as orc
    bcs = [bytes(r) for r in tab]
    n = sample_text.size // rec_bytes
    t0 = time.perf_counter()
    got, _, _, _ = orc.demux_stream(sample_text, bcs, w.mismatch)
    dt1 = time.perf_counter() - t0
    assert got == n
    single = n / dt1
    allc = None
    if threads > 1:
        allc = cpu_reference_parallel(w, tab, sample_text, rec_bytes, threads)[2]
    return single, dt1, allc


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU algorithm on the box's host cores (the oracle port: the Rust
    binary cannot be built here).  The reference itself is single-threaded; to give it every host thread it
    could use, the same loop runs on disjoint shards of the sample, one per hardware thread, and `value` is
    that all-threads figure (the single-thread rate is reported beside it).  Rank 0 only.
    The input comes from the oracle's own restatement of the synthetic generator (oracle/synth_ref.c, byte-equal
    to the product's: tests/test_oracle.py), so this process never maps libsabreur_gpu.so."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from sabreur_b200 import synth  # the workload shapes only (plain Python; the CUDA library is not loaded)
    tab = orc.synth_table(w.cfg_id, w.n_barcodes, w.bc_len, w.min_dist)
    rec = w.rec_bytes
    threads = os.cpu_count() or 1
    seed = synth.SEED_READS + w.cfg_id

    def sample(n):
        return orc.synth_reads(w.fmt, w.read_len, 0, seed, tab, 0, n, threads=threads)

    # calibrate a sample that the host handles in roughly a second per step
    s1, _, _ = cpu_reference(w, tab, sample(100_000), rec, 1)
    n = int(max(100_000, min(16_000_000, s1 * 1.0 * threads))) // threads * threads
    text = sample(n)
    for _ in range(args.warmup):
        cpu_reference(w, tab, text[: (n // 8) * rec], rec, 1)
    t_tot = 0.0
    for i in range(args.steps):
        t0 = time.perf_counter()
        cpu_reference_parallel(w, tab, text, rec, threads)
        t_tot += time.perf_counter() - t0
    single, _, _ = cpu_reference(w, tab, text[: (n // threads) * rec], rec, 1)
    div = 2 if w.paired else 1
    value = (n / div) * args.steps / t_tot
    unit = "pairs/s" if w.paired else "reads/s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w.name, "sample_reads_per_step": n},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": f"{n} reads of the same synthetic stream per step, split into {threads} disjoint shards, "
                                   f"one per hardware thread (the reference itself is single-threaded, Cargo.toml:14-22)",
                         "single_thread": {"value": single / div, "cores": 1}},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(local_rank: int):
    """Best effort: run this rank on the CPUs of its GPU's NUMA node, so that the pinned slot buffers (first touch) and
    the copies stay on the local socket.  Returns a short description for the JSON line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
        if node < 0:
            return {"gpu": bdf, "numa_node": None, "host_nodes": len(nodes)}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
        return {"gpu": bdf, "numa_node": node, "host_nodes": len(nodes), "cpus_bound": len(use), "cpus_allowed": len(allowed)}
    except Exception as e:  # no sysfs, no permission ...: run unbound
        return {"error": str(e)[:80]}


class Env:
    """what every workload block needs: ranks, barriers, the peak"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            self.peak, self.peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, x: float, op="max") -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())


def h2d_ceiling(env: Env, seconds: float = 0.4) -> dict:
    """what the box moves from pinned host memory to the GPUs when every rank copies at once: bare 256 MiB
    cudaMemcpyAsync on two streams per rank, no kernels.  The end-to-end number is reported against it."""
    torch = env.torch
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
    s = [torch.cuda.Stream() for _ in range(2)]

    def burst(reps):
        for i in range(reps):
            with torch.cuda.stream(s[i & 1]):
                d[i & 1].copy_(h, non_blocking=True)
        for st in s:
            st.synchronize()

    burst(4)
    env.barrier()
    t0 = time.perf_counter()
    burst(8)
    probe = time.perf_counter() - t0
    reps = int(max(8, min(400, 8 * seconds / max(probe, 1e-3)))) // 2 * 2
    reps = int(env.reduce(float(reps), "max"))
    env.barrier()
    t0 = time.perf_counter()
    burst(reps)
    dt = env.reduce(time.perf_counter() - t0, "max")
    env.barrier()
    return {"h2d_ceiling_gbs": env.world * reps * n / dt / 1e9,
            "h2d_ceiling_how": f"{env.world} rank(s) x {reps} pinned 256 MiB cudaMemcpyAsync on 2 streams each, all at once, max over ranks"}


def verify_merge(env: Env, n_reads: int = 1_048_576, n_batches: int = 16) -> dict:
    """The multi-GPU product path in small: a slice of cfg3 is cut into batches, batch s runs on rank s mod N through the
    streaming API (pinned host text -> sbr_submit -> sbr_wait), the per-batch results are gathered on rank 0 and merged
    back into the sequential reference's view by sabreur_b200/shard.py (counts summed, per-barcode index lists appended in
    batch order).  Checked without the oracle: the merged lists must be ascending, cover every read exactly once, agree
    with the all-reduced counts, and every read's bin must be the first barcode row within the mismatch budget of its
    prefix (src/demux.rs:52-54), recomputed here with numpy."""
    import numpy as np
    from sabreur_b200 import _ffi, demux, shard, synth
    w = synth.CFG3
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    rec, k = w.rec_bytes, w.bc_len
    per = n_reads // n_batches
    first = 77_000_000
    mine = list(shard.batches_of_rank(n_batches, env.world, env.rank))
    results = []
    with demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=env.local_rank, batch_bytes=per * rec + 4096, n_slots=2,
                     max_records=per + 16) as dm:
        for j, s in enumerate(mine):  # two slots: batch j+1 is submitted before batch j is waited for
            buf = dm.slot_buffer(j % 2)
            buf[: per * rec] = synth.reads_host(w, tab, first + s * per, per)
            dm.submit(j % 2, 0, per * rec, s, True)
            if j >= 1:
                b = dm.wait((j - 1) % 2)
                results.append(shard.BatchResult(int(b.seq), b.n_records, b.bin_count, b.bin_start, b.order, b.status, b.first_bad_record))
        if mine:
            b = dm.wait((len(mine) - 1) % 2)
            results.append(shard.BatchResult(int(b.seq), b.n_records, b.bin_count, b.bin_start, b.order, b.status, b.first_bad_record))
        n_bins = dm.n_bins
    local_counts = np.zeros(n_bins, dtype=np.uint64)
    for r in results:
        local_counts += r.bin_count.astype(np.uint64)
    total_counts = shard.merge_counts(local_counts)
    everything = shard.gather_results(results, dst=0)
    if env.rank != 0:
        return {}
    merged = shard.merge_in_order(sorted(everything, key=lambda r: (r.seq % env.world, r.seq)), n_bins)  # arrival order: rank by rank
    lists = merged.lists()
    ok = merged.n_records == n_reads and merged.status == 0 and np.array_equal(merged.bin_count, total_counts)
    assign = np.full(n_reads, 0xFFFF, dtype=np.uint16)
    for b_, idx in enumerate(lists):
        ok = ok and (idx.size < 2 or bool((np.diff(idx.astype(np.int64)) > 0).all()))  # input order inside every file
        assign[idx.astype(np.int64)] = b_
    ok = ok and not (assign == 0xFFFF).any()
    text = synth.reads_host(w, tab, first, n_reads).reshape(n_reads, rec)
    pre = text[:, 17:17 + k]
    want = np.full(n_reads, len(bcs), dtype=np.uint16)
    todo = np.ones(n_reads, dtype=bool)
    for b_ in range(len(bcs)):
        hit = todo & ((pre != tab[b_]).sum(1) <= w.mismatch)
        want[hit] = b_
        todo &= ~hit
    ok = ok and np.array_equal(assign, want)
    if not ok:
        raise AssertionError("verify_merge: the merged multi-GPU partition is not the sequential one")
    return {"reads": n_reads, "batches": n_batches, "ranks": env.world, "ok": True,
            "how": "batch s on rank s mod N through sbr_submit/sbr_wait; shard.gather_results + merge_in_order + merge_counts; "
                   "merged lists ascending, complete, equal to the all-reduced counts and to numpy's first-match over the prefixes"}


def run_block(env: Env, name: str, units: int, primary: bool) -> dict:
    """one workload: generate the shard on the device, self-check every batch, time K steps HBM-resident, then end to end"""
    import numpy as np
    from sabreur_b200 import _ffi, demux, synth
    torch, args = env.torch, env.args
    w = synth.WORKLOADS[name]
    mates = (0, 1) if w.paired else (0,)
    rec = w.rec_bytes
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    # Batch size: the scan cuts a batch into one region of whole tiles per resident CTA, so batches of
    # rows x regions x tile bytes (the largest that stays below the 1 GiB a batch may hold) load every region evenly;
    # 3,000,000-read batches (round 1) left 76 of 592 regions with a 45th tile while the others were done after 44.
    with demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=env.local_rank, batch_bytes=1 << 20, n_slots=1) as probe:
        tile_bytes, regions = probe.scan_geometry(w.fmt)
    row = tile_bytes * regions
    bu = args.batch_units or ((1 << 30) // row * row // rec)
    bu = max(16, min(bu, units) // 16 * 16)
    if bu * rec > (1 << 30):
        bu = ((1 << 30) // rec) // 16 * 16
    first = env.rank * units  # this rank's shard of the read index space
    # ---- resident shard -----------------------------------------------------------------------------
    texts = []
    for mate in mates:
        t = torch.empty(units * rec + 4096, dtype=torch.uint8, device="cuda")
        synth.reads_device(w, tab, first, units, t.data_ptr(), device=env.local_rank, mate=mate)
        texts.append(t)
    torch.cuda.synchronize()
    batches = []  # (tensor index, byte offset, nbytes)
    for ti in range(len(texts)):
        for lo in range(0, units, bu):
            n = min(bu, units - lo)
            batches.append((ti, lo * rec, n * rec))
    overlap = not args.no_overlap
    flags = (_ffi.CFG_TIMING | (_ffi.CFG_NO_LUT if args.no_lut else 0) | (_ffi.CFG_OVERLAP if overlap else 0) |
             (_ffi.CFG_NO_FUSE if args.no_fuse else 0))
    S = max(2, args.slots)
    dm = demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=env.local_rank, batch_bytes=bu * rec + 4096, n_slots=S,
                     max_records=bu + 16, flags=flags)
    tstream = torch.cuda.Stream()  # a real stream: handle 0 would select the slot's own stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    holds = [None] * S   # which batch each device workspace holds
    seq = [0]            # launches so far: workspaces rotate across steps too (a step's last batch and the next step's first
                         # one must not land in the same workspace, or the second would wait for the first's match kernel)

    def launch(i):
        ti, off, nb = batches[i]
        slot = seq[0] % S
        seq[0] += 1
        holds[slot] = i
        dm.demux_device(slot, texts[ti].data_ptr() + off, nb, final=True, stream=stream)
        return slot

    def step():
        for i in range(len(batches)):
            launch(i)

    def check(i, b):
        n = batches[i][2] // rec
        ok = (b.n_records == n and b.status == 0 and b.scan_path == 1 and b.consumed == batches[i][2] and
              int(b.bin_count.sum(dtype=np.uint64)) == n and int(b.bin_start[-1]) == n)
        if not ok and not os.environ.get("SBR_DEBUG_ABLATE"):
            raise AssertionError(f"bench self-check failed on {name} batch {i}: n_records {b.n_records} (want {n}), status "
                                 f"{b.status:#x}, scan_path {b.scan_path}, consumed {b.consumed}, sum(bin_count) {int(b.bin_count.sum())}")
        return int(b.bin_count[-1])

    # ---- warm-up 1: every batch of the step once, each one waited for and checked --------------------------
    unknown = 0
    for i in range(len(batches)):
        slot = launch(i)
        unknown += check(i, dm.wait(slot, copy=False))
        holds[slot] = None  # waited for: nothing in flight there
    for _ in range(max(0, args.warmup - 1)):
        step()
    dm.join(stream)
    torch.cuda.synchronize()
    dm.timing()            # reset the per-kernel event ring
    dm.scan_stats(True)    # ... and the device-side counters of the scan
    clocks = Clocks(env.local_rank)
    clocks.start()
    time.sleep(1.0)        # the sampler is up and has idle samples well before the load starts
    env.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    dm.join(stream)        # the last batches' match + partition run on the slots' own streams
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    env.barrier()
    ms = env.reduce(e0.elapsed_time(e1))
    clk = clocks.summary(t_begin, t_end)
    kms, nb_timed, launches = dm.timing()
    finished, handed_over = dm.scan_stats()
    # the batches still sitting in the slots are the timed loop's last ones: wait for them and check them as well
    last_checked = 0
    for slot in range(S):
        if holds[slot] is not None:
            check(holds[slot], dm.wait(slot, copy=False))
            last_checked += 1
    want_batches = args.steps * len(batches)
    self_check = {"batches": len(batches), "checked_each": ["n_records", "status==0", "scan_path==1", "consumed", "sum(bin_count)"],
                  "timed_batches": want_batches, "scan_finished": finished, "spec_fail": handed_over,
                  "rechecked_after_timed_loop": last_checked, "unknown_frac": unknown / (units * len(mates))}
    if (finished != want_batches or handed_over != 0) and not os.environ.get("SBR_DEBUG_ABLATE"):
        raise AssertionError(f"{name}: the timed loop launched {want_batches} batches, the scan finished {finished} and "
                             f"handed {handed_over} over to the exact scan")
    value = env.world * units * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (record scan) and of the whole path ------------------------------
    avg_batch_reads = sum(nb // rec for _, _, nb in batches) / len(batches)
    reads_timed = avg_batch_reads * nb_timed  # the event ring may cover fewer batches than were launched
    scan_bytes_per_read = rec + 8           # reads the record, writes rec_off + packed prefix slot (L_rec + 8)
    path_bytes_per_read = w.algorithmic_bytes_per_read  # L_rec + k + 22
    scan_launches = max(1, nb_timed)
    scan_avg_ms = kms[0] / scan_launches
    achieved = (reads_timed / scan_launches) * scan_bytes_per_read / (scan_avg_ms * 1e-3) / 1e9 if scan_avg_ms > 0 else 0.0
    step_ms = ms / args.steps
    path_gbs = units * len(mates) * path_bytes_per_read / (step_ms * 1e-3) / 1e9
    ksum_ms = sum(kms)
    roofline = {
        "bound": "hbm", "kernel": "k_scan_fast (record-boundary scan + prefix pack)", "achieved": achieved, "peak": env.peak,
        "unit": "GB/s", "frac": achieved / env.peak, "traffic": None, "peak_source": env.peak_src,
        "bytes_per_launch": (reads_timed / scan_launches) * scan_bytes_per_read, "avg_launch_ms": scan_avg_ms,
        "kernel_ms": {"scan": kms[0], "match": kms[1], "partition": kms[2], "batches": nb_timed,
                      "note": ("match + partition run on the slots' own streams beside the next batch's scan: the three do not "
                               "add up to the step") if overlap else "one stream: the three add up to the step"},
        "path": {"bytes_per_read": path_bytes_per_read, "achieved": path_gbs, "frac": path_gbs / env.peak,
                 "how": "algorithmic bytes of the whole shard / wall time of a step (device events around the timed loop)",
                 "kernel_sum_frac": (reads_timed * path_bytes_per_read / (ksum_ms * 1e-3) / 1e9 / env.peak) if ksum_ms > 0 else 0.0},
    }
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # ncu dram bytes per scan launch, when profiles/ holds a capture of this workload
        try:
            tj = json.load(open(tpath)).get(name, {})
            if tj.get("k_scan_dram_bytes_per_launch"):  # measured on one launch of `reads_per_launch` reads: scale to this batch
                roofline["traffic"] = tj["k_scan_dram_bytes_per_launch"] * (reads_timed / scan_launches) / tj["reads_per_launch"]
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass
    dm.close()
    if overlap:
        # the scan kernel on its own: the same batches through a one-stream context (match + partition BEHIND the scan
        # instead of beside the next one), a few batches only -- the figure that compares with round 1's and with ncu's
        # serialised launch list; `frac` above is the kernel's time in the timed loop, where it shares the SMs
        dm1 = demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=env.local_rank, batch_bytes=bu * rec + 4096, n_slots=2,
                          max_records=bu + 16, flags=flags & ~_ffi.CFG_OVERLAP)
        nb1 = min(len(batches), 12)
        for rnd in range(2):
            if rnd == 1:
                torch.cuda.synchronize()
                dm1.timing()
            for i in range(nb1):
                ti, off, nb = batches[i]
                dm1.demux_device(i % 2, texts[ti].data_ptr() + off, nb, final=True, stream=stream)
        torch.cuda.synchronize()
        k1, n1, _ = dm1.timing()
        dm1.close()
        if n1 and k1[0] > 0:
            reads1 = sum(batches[i][2] // rec for i in range(nb1)) / nb1
            alone = reads1 * scan_bytes_per_read / (k1[0] / n1 * 1e-3) / 1e9
            roofline["alone"] = {"achieved": alone, "frac": alone / env.peak, "avg_launch_ms": k1[0] / n1, "batches": n1,
                                 "match_partition_ms": (k1[1] + k1[2]) / n1,
                                 "how": "the same kernel in a one-stream loop (nothing else on the SMs), untimed extra pass"}

    # ---- end to end: pinned host text -> sbr_submit -> sbr_wait (H2D + kernels + D2H every batch) -------
    e2e = None
    if not args.no_e2e:
        eb = max(16, (args.e2e_batch_mb << 20) // rec // 16 * 16)
        eb = min(eb, units // 16 * 16)
        ES = args.e2e_slots
        dm2 = demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=env.local_rank, batch_bytes=eb * rec + 4096, n_slots=ES,
                          max_records=eb + 16)
        lib = _ffi.load()
        for s in range(ES):  # distinct synthetic batches, generated on the GPU and brought to pinned memory
            buf = dm2.slot_buffer(s)
            t = texts[s % len(texts)]
            start_rec = (s * eb) % max(1, units - eb + 1)
            assert lib.sbr_dev_download(env.local_rank, buf.ctypes.data, t.data_ptr() + start_rec * rec, eb * rec) == 0
        n_e2e = max(ES, (units * len(mates) + eb - 1) // eb)  # batches per step: the whole shard's worth of reads

        def take(i):
            r = dm2.wait(i % ES, copy=False)
            if not (r.n_records == eb and r.status == 0 and r.scan_path == 1 and int(r.bin_start[-1]) == eb):
                raise AssertionError(f"{name} e2e batch {i}: n_records {r.n_records} (want {eb}), status {r.status:#x}, scan_path {r.scan_path}")
            return 4 * (r.n_records + 1) + 6 * r.n_records + 8 * dm2.n_bins + 68

        def e2e_step():
            d2h = 0
            for i in range(n_e2e):
                if i >= ES:
                    d2h += take(i)
                dm2.submit(i % ES, 0, eb * rec, i, True)
            for i in range(n_e2e, n_e2e + ES):
                d2h += take(i)
            return d2h

        e2e_steps = max(1, min(args.steps, 5 if primary else 2))
        e2e_step() if args.warmup else None
        torch.cuda.synchronize()
        env.barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            d2h = e2e_step()
        torch.cuda.synchronize()
        dt = env.reduce(time.perf_counter() - t0)
        env.barrier()
        e2e_units = n_e2e * eb / len(mates)
        h2d = env.world * n_e2e * eb * rec
        e2e = {"value": env.world * e2e_units * e2e_steps / dt, "unit": "pairs/s" if w.paired else "reads/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": env.world * d2h, "steps": e2e_steps,
               "ms_per_step": 1e3 * dt / e2e_steps, "batch_bytes": eb * rec, "slots": ES,
               "h2d_gbs": h2d * e2e_steps / dt / 1e9, "checked_every_batch": True,
               "how": "pinned host text -> sbr_submit (H2D, scan, match, partition) -> sbr_wait (D2H rec_off, assign, order)"}
        dm2.close()
    clocks.stop()
    if e2e is not None:
        clk["e2e_region"] = {k: v for k, v in clocks.summary(t_end, time.perf_counter()).items() if k in ("sm_mhz", "samples", "reasons")}

    # ---- CPU baseline beside it (rank 0, N = 1, primary workload only) -------------------------------------
    cpu = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu:  # (a quarter of the sample for the secondary workloads)
        probe_n = 200_000 if primary else 50_000
        probe = texts[0][: probe_n * rec].cpu().numpy()
        s1, _, _ = cpu_reference(w, tab, probe, rec, 1)
        n = int(max(probe_n, min(units, 16_000_000, s1 * args.cpu_seconds * (1.0 if primary else 0.25))))
        sample = texts[0][: n * rec].cpu().numpy()
        threads = os.cpu_count() or 1
        single, dt1, allc = cpu_reference(w, tab, sample, rec, threads)
        div = 2 if w.paired else 1
        cpu = {"value": single / div, "unit": "pairs/s" if w.paired else "reads/s", "cores": 1, "kind": "port",
               "sample": f"first {n} reads of rank 0's shard ({n * rec / 1e9:.2f} GB), {dt1:.1f} s on one core; oracle "
                         f"port of src/demux.rs:49-66 (the Rust reference cannot be built here and is single-threaded)",
               "allcores": {"value": allc / div if allc else None, "cores": threads,
                            "how": "same loop on disjoint shards, one per hardware thread"}}
    del texts
    torch.cuda.empty_cache()
    total = w.total_units
    return {
        "value": value, "unit": "pairs/s" if w.paired else "reads/s", "ms_per_step": step_ms, "steps": args.steps,
        "warmup": args.warmup, "gpu_launches": int(launches),
        "config": {"workload": w.name, "units_per_gpu": units, "batch_reads": bu, "batches_per_step": len(batches),
                   "record_bytes": rec, "resident_bytes_per_gpu": units * rec * len(mates), "device_slots": S,
                   "batch_geometry": f"{-(-bu * rec // row)} tile rows x {regions} regions x {tile_bytes}-byte tiles "
                                     f"({-(-bu * rec // row) * row - bu * rec} bytes short of full rows)",
                   "scale": f"{env.world * units} of the configuration's {total} {'pairs' if w.paired else 'reads'} "
                            f"({env.world * units / total:.3g}x) over {env.world} GPU(s); a rate metric",
                   "pipeline": ("scan on the caller's stream; match + partition (one cooperative launch) on the slot's own stream, "
                                "beside the next batch's scan") if overlap else "scan, match + partition on one stream",
                   "l2": f"inputs ({units * rec * len(mates) / 1e9:.0f} GB per step) are far larger than the 126 MB L2; no flush needed",
                   "match": "xor+popc table walk" if args.no_lut else "first-match LUT built by the xor+popc kernel"},
        "roofline": roofline, "self_check": self_check, "clocks": clk, "e2e": e2e, "cpu_baseline": cpu,
    }


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from sabreur_b200 import synth
    if args.impl == "reference":
        run_reference(args, synth.WORKLOADS[args.workload], rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demux path has no CPU fallback")
    # stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = Env(args)

    primary_units = args.units_per_gpu or default_units(args.workload)
    scale = primary_units / default_units(args.workload)
    also = args.also if args.also is not None else ("cfg4,cfg5" if args.workload == "cfg3" else "none")
    others = [x for x in also.split(",") if x and x != "none" and x != args.workload]

    merged = None if args.no_verify_merge else verify_merge(env)
    ceiling = None if args.no_e2e else h2d_ceiling(env)
    main_block = run_block(env, args.workload, primary_units, True)
    blocks = {}
    for name in others:
        units = max(16, int(default_units(name) * scale) // 16 * 16)
        blocks[name] = run_block(env, name, units, False)
    for b in [main_block] + list(blocks.values()):
        if b["e2e"] is not None and ceiling:
            b["e2e"].update(ceiling)
            b["e2e"]["h2d_frac_of_ceiling"] = b["e2e"]["h2d_gbs"] / ceiling["h2d_ceiling_gbs"]
            b["e2e"]["host_placement_rank0"] = numa

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_block["value"], "unit": main_block["unit"], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_block["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": main_block["config"], "roofline": main_block["roofline"], "self_check": main_block["self_check"],
            "cpu_baseline": main_block["cpu_baseline"], "e2e": main_block["e2e"], "clocks": main_block["clocks"],
            "gpu_launches": main_block["gpu_launches"] + sum(b["gpu_launches"] for b in blocks.values()),
            "verify_merge": merged,
            "workloads": blocks,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
