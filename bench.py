#!/usr/bin/env python
"""bench.py -- reads/s demuxed (record scan + barcode match + partition) on B200, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg4|cfg5] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one GPU's shard of the synthetic workload, resident in HBM
(`value`), or streamed from pinned host memory through sbr_submit/sbr_wait (`e2e`).  Shards are
independent (reads partition by batch; no collective on the data path): weak scaling.
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
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--units-per-gpu", type=int, default=0, help="reads (SE) / pairs (PE) resident per GPU")
    ap.add_argument("--batch-units", type=int, default=0, help="reads per kernel batch (multiple of 16)")
    ap.add_argument("--e2e-batch-mb", type=int, default=256)
    ap.add_argument("--e2e-slots", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-lut", action="store_true", help="plain XOR+popc table walk in the match kernel")
    return ap.parse_args()


def default_units(w) -> int:
    # cfg3: the whole 100 M-read configuration fits one GPU (32.1 GB).  cfg4: 1 B pairs / 8 GPUs = 125 M
    # pairs (80 GB) per GPU.  cfg5: 250 M reads (66.5 GB) per GPU.
    return {"cfg3": 100_000_000, "cfg4": 125_000_000, "cfg5": 250_000_000}[w]


# ---------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def host_sample(w, tab, n, mate=0):
    from sabreur_b200 import synth
    return synth.reads_host(w, tab, 0, n, mate)


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU algorithm on the box's host cores (the oracle port: the Rust
    binary cannot be built here).  The reference itself is single-threaded; to give it every host thread it
    could use, the same loop runs on disjoint shards of the sample, one per hardware thread, and `value` is
    that all-threads figure (the single-thread rate is reported beside it).  Rank 0 only."""
    if rank != 0:
        return
    from sabreur_b200 import synth
    tab = synth.table(w)
    rec = w.rec_bytes
    threads = os.cpu_count() or 1
    # calibrate a sample that the host handles in roughly a second per step
    probe = host_sample(w, tab, 100_000)
    s1, _, _ = cpu_reference(w, tab, probe, rec, 1)
    n = int(max(100_000, min(16_000_000, s1 * 1.0 * threads))) // threads * threads
    text = host_sample(w, tab, n)
    for _ in range(args.warmup):
        cpu_reference(w, tab, text[: (n // 8) * rec], rec, 1)
    t_tot = 0.0
    for i in range(args.steps):
        t0 = time.perf_counter()
        _, _, allc = cpu_reference_parallel(w, tab, text, rec, threads)
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
        if node < 0:
            return {"gpu": bdf, "numa_node": None}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
        return {"gpu": bdf, "numa_node": node, "cpus_bound": len(use), "cpus_allowed": len(allowed)}
    except Exception as e:  # no sysfs, no permission ...: run unbound
        return {"error": str(e)[:80]}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from sabreur_b200 import synth
    w = synth.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from sabreur_b200 import _ffi, demux

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

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    units = args.units_per_gpu or default_units(args.workload)
    mates = (0, 1) if w.paired else (0,)
    rec = w.rec_bytes
    bu = args.batch_units or (3_000_000 if w.fmt == _ffi.FMT_FASTQ else 3_600_000)
    bu = max(16, min(bu, units) // 16 * 16)
    if bu * rec > (1 << 30):
        bu = ((1 << 30) // rec) // 16 * 16
    tab = synth.table(w)
    bcs = [bytes(r) for r in tab]
    first = rank * units  # this rank's shard of the read index space
    # ---- resident shard -----------------------------------------------------------------------------
    texts = []
    for mate in mates:
        t = torch.empty(units * rec + 4096, dtype=torch.uint8, device="cuda")
        synth.reads_device(w, tab, first, units, t.data_ptr(), device=local_rank, mate=mate)
        texts.append(t)
    torch.cuda.synchronize()
    batches = []  # (tensor index, byte offset, nbytes)
    for ti in range(len(texts)):
        for lo in range(0, units, bu):
            n = min(bu, units - lo)
            batches.append((ti, lo * rec, n * rec))
    flags = _ffi.CFG_TIMING | (_ffi.CFG_NO_LUT if args.no_lut else 0)
    dm = demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=local_rank, batch_bytes=bu * rec + 4096, n_slots=2,
                     max_records=bu + 16, flags=flags)
    tstream = torch.cuda.Stream()  # a real stream: handle 0 would select the slot's own stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step():
        for i, (ti, off, nb) in enumerate(batches):
            dm.demux_device(i & 1, texts[ti].data_ptr() + off, nb, final=True, stream=stream)

    # correctness guard inside the bench: the first batch's counts must add up
    step()
    b = dm.wait((len(batches) - 1) & 1)
    last_n = batches[-1][2] // rec
    if not os.environ.get("SBR_DEBUG_ABLATE"):
        assert b.n_records == last_n and int(b.bin_count.sum()) == last_n and b.status == 0, "bench self-check failed"
    for _ in range(max(0, args.warmup - 1)):
        step()
    torch.cuda.synchronize()
    dm.timing()  # reset the per-kernel event ring
    clocks = Clocks(local_rank)
    barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    kms, nb_timed, launches = dm.timing()
    value = world * units * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (record scan) --------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    avg_batch_reads = sum(nb // rec for _, _, nb in batches) / len(batches)
    reads_timed = avg_batch_reads * nb_timed  # the event ring may cover fewer batches than were launched
    scan_bytes_per_read = rec + 8           # reads the record, writes rec_off + packed prefix slot (L_rec + 8)
    path_bytes_per_read = w.algorithmic_bytes_per_read  # L_rec + k + 22
    scan_launches = max(1, nb_timed)
    scan_avg_ms = kms[0] / scan_launches
    achieved = (reads_timed / scan_launches) * scan_bytes_per_read / (scan_avg_ms * 1e-3) / 1e9 if scan_avg_ms > 0 else 0.0
    path_ms = sum(kms)
    roofline = {
        "bound": "hbm", "kernel": "k_scan (record-boundary scan + prefix pack)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "bytes_per_launch": (reads_timed / scan_launches) * scan_bytes_per_read, "avg_launch_ms": scan_avg_ms,
        "kernel_ms": {"scan": kms[0], "match": kms[1], "partition": kms[2]},
        "path": {"bytes_per_read": path_bytes_per_read,
                 "achieved": reads_timed * path_bytes_per_read / (path_ms * 1e-3) / 1e9 if path_ms > 0 else 0.0,
                 "frac": (reads_timed * path_bytes_per_read / (path_ms * 1e-3) / 1e9 / peak) if path_ms > 0 else 0.0},
    }
    # the profile summary written by profiles/ (ncu dram bytes per launch), when present
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(args.workload, {})
            if tj.get("k_scan_dram_bytes_per_launch"):  # measured on one launch of `reads_per_launch` reads: scale to this batch
                roofline["traffic"] = tj["k_scan_dram_bytes_per_launch"] * (reads_timed / scan_launches) / tj["reads_per_launch"]
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- end to end: pinned host text -> sbr_submit -> sbr_wait (H2D + kernels + D2H every batch) -------
    e2e = None
    if not args.no_e2e:
        eb = max(16, (args.e2e_batch_mb << 20) // rec // 16 * 16)
        eb = min(eb, units // 16 * 16)
        S = args.e2e_slots
        dm2 = demux.Demux(bcs, w.mismatch, fmt=w.fmt, device=local_rank, batch_bytes=eb * rec + 4096, n_slots=S,
                          max_records=eb + 16)
        lib = _ffi.load()
        for s in range(S):  # distinct synthetic batches, generated on the GPU and brought to pinned memory
            buf = dm2.slot_buffer(s)
            t = texts[s % len(texts)]
            start_rec = (s * eb) % max(1, units - eb + 1)
            assert lib.sbr_dev_download(local_rank, buf.ctypes.data, t.data_ptr() + start_rec * rec, eb * rec) == 0
        n_e2e = max(S, (units * len(mates) + eb - 1) // eb)  # batches per step: the whole shard's worth of reads

        def e2e_step():
            d2h = 0
            for i in range(n_e2e):
                s = i % S
                if i >= S:
                    r = dm2.wait(s, copy=False)
                    d2h += 4 * (r.n_records + 1) + 6 * r.n_records + 8 * dm2.n_bins + 68
                dm2.submit(s, 0, eb * rec, i, True)
            for i in range(n_e2e, n_e2e + S):
                r = dm2.wait(i % S, copy=False)
                d2h += 4 * (r.n_records + 1) + 6 * r.n_records + 8 * dm2.n_bins + 68
            return d2h

        e2e_steps = max(1, min(args.steps, 5))
        e2e_step() if args.warmup else None
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            d2h = e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_units = n_e2e * eb / len(mates)
        e2e = {"value": world * e2e_units * e2e_steps / dt, "unit": "pairs/s" if w.paired else "reads/s",
               "h2d_bytes_per_step": world * n_e2e * eb * rec, "d2h_bytes_per_step": world * d2h, "steps": e2e_steps,
               "host_placement_rank0": numa,
               "ms_per_step": 1e3 * dt / e2e_steps, "batch_bytes": eb * rec, "slots": S,
               "how": "pinned host text -> sbr_submit (H2D, scan, match, partition) -> sbr_wait (D2H rec_off, assign, order)"}
        dm2.close()

    clk = clocks.stop()  # sampled over both timed regions (HBM-resident steps and the end-to-end steps)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        probe_n = 200_000
        probe = texts[0][: probe_n * rec].cpu().numpy()
        s1, _, _ = cpu_reference(w, tab, probe, rec, 1)
        n = int(max(probe_n, min(units, 16_000_000, s1 * args.cpu_seconds)))
        sample = texts[0][: n * rec].cpu().numpy()
        threads = os.cpu_count() or 1
        single, dt1, allc = cpu_reference(w, tab, sample, rec, threads)
        div = 2 if w.paired else 1
        cpu = {"value": single / div, "unit": "pairs/s" if w.paired else "reads/s", "cores": 1, "kind": "port",
               "sample": f"first {n} reads of rank 0's shard ({n * rec / 1e9:.2f} GB), {dt1:.1f} s on one core; oracle "
                         f"port of src/demux.rs:49-66 (the Rust reference cannot be built here and is single-threaded)",
               "allcores": {"value": allc / div if allc else None, "cores": threads,
                            "how": "same loop on disjoint shards, one per hardware thread"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s" if w.paired else "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": w.name, "units_per_gpu": units, "batch_reads": bu, "batches_per_step": len(batches),
                       "record_bytes": rec, "resident_bytes_per_gpu": units * rec * len(mates),
                       "l2": "inputs (>= 32 GB per step) are far larger than the 126 MB L2; no flush needed",
                       "match": "xor+popc table walk" if args.no_lut else "first-match LUT built by the xor+popc kernel"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "gpu_launches": int(launches),
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
