#!/usr/bin/env python
"""Copies the summaries of a round-2 GPU session from gpurun_out/ (scratch) into profiles/ (tracked).

  python tools/refresh_profiles_r2.py BENCH_TAG NCU_TAG [LAUNCH_TAG]

BENCH_TAG: tools/gpu_r2.sh session (bench_<tag>.json, bench_ref_<tag>.json, bench_<tag>_serial.json, e2e_cli_<tag>.jsonl);
NCU_TAG: the session that holds prof_{scan,match,cfg4,cfg5}_<tag>.ncu-rep; LAUNCH_TAG: launches_<tag>.csv."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
btag, ntag = sys.argv[1], sys.argv[2]
ltag = sys.argv[3] if len(sys.argv) > 3 else ntag


def last_line(path):
    return open(path).read().strip().splitlines()[-1]


for src, dst in ((f"bench_{btag}.json", "r2_bench.json"), (f"bench_ref_{btag}.json", "r2_bench_ref.json"),
                 (f"bench_{btag}_serial.json", "r2_bench_one_stream.json")):
    if os.path.exists(os.path.join(G, src)):
        open(os.path.join(P, dst), "w").write(last_line(os.path.join(G, src)) + "\n")
if os.path.exists(os.path.join(G, f"e2e_cli_{btag}.jsonl")):
    open(os.path.join(P, "r2_e2e_cli.jsonl"), "w").write(open(os.path.join(G, f"e2e_cli_{btag}.jsonl")).read())
if os.path.exists(os.path.join(G, f"launches_{ltag}.csv")):
    open(os.path.join(P, "r2_launches.csv"), "w").write(open(os.path.join(G, f"launches_{ltag}.csv")).read())

traffic = json.load(open(os.path.join(P, "traffic.json")))
units = {"cfg3": 3_331_296, "cfg4": 3_331_296, "cfg5": 4_020_112}  # one full batch of 49 tile rows each (the first batch of a shard)
for rep, out, wl in ((f"prof_scan_{ntag}", "r2_scan_fast_ncu.txt", "cfg3"), (f"prof_match_{ntag}", "r2_match_partition_ncu.txt", None),
                     (f"prof_cfg4_{ntag}", "r2_cfg4_scan_match_ncu.txt", "cfg4"), (f"prof_cfg5_{ntag}", "r2_cfg5_scan_match_ncu.txt", "cfg5")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path], capture_output=True, text=True).stdout
    open(os.path.join(P, out), "w").write(summ)
    if wl is None:
        continue
    rd = wr = None
    in_scan = False
    for ln in summ.splitlines():
        if ln.startswith("=="):
            in_scan = "k_scan_fast" in ln
        f = ln.split()
        if in_scan and len(f) >= 3 and f[0] == "dram__bytes_read.sum" and rd is None:
            rd = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
        if in_scan and len(f) >= 3 and f[0] == "dram__bytes_write.sum" and wr is None:
            wr = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
    if rd is not None and wr is not None:
        traffic[wl] = {"k_scan_dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "reads_per_launch": units[wl],
                       "source": f"profiles/{out}"}
traffic["_how"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE k_scan_fast launch per workload (ncu --set full --clock-control none, "
                   f"round 2 gpurun session {ntag}, bench.py --units-per-gpu 7000000 (cfg5: 9000000) --no-verify-merge: the shard's first batch, 49 tile rows = "
                   "3,331,296 FASTQ reads or 4,020,112 FASTA reads (1.069 GB of text). Algorithmic bytes of the cfg3 launch: 3,331,296 x (321 + 8) = 1,095,996,384; the "
                   "excess are the tiles every region runs over into (and prefetches) to finish its last record.")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
d = json.loads(open(os.path.join(P, "r2_bench.json")).read())
for k, b in {"cfg3": d, **d.get("workloads", {})}.items():
    r = b["roofline"]
    print(f"{k}: {b['value'] / 1e9:.3f} G {b['unit']}, {b['ms_per_step']:.3f} ms/step, scan {r['achieved']:.0f} GB/s ({r['frac']:.3f}), path {r['path']['frac']:.3f}, "
          f"e2e {b['e2e']['value'] / 1e6:.1f} M/s" if b.get("e2e") else "")
