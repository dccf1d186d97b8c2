#!/usr/bin/env python
"""Copies the summaries of a gpu_round.sh session from gpurun_out/ into profiles/ (tracked).  usage: refresh_profiles.py TAG"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for old in glob.glob(os.path.join(P, "r1_scan_fast_?_ncu.txt")) + glob.glob(os.path.join(P, "r1_launches_?.csv")) + \
        glob.glob(os.path.join(P, "bench_r1_?.json")) + glob.glob(os.path.join(P, "bench_ref_r1_?.json")):
    os.remove(old)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"prof_scan_{tag}.ncu-rep")],
                      capture_output=True, text=True).stdout
open(os.path.join(P, f"r1_scan_fast_{tag}_ncu.txt"), "w").write(summ)
open(os.path.join(P, f"r1_launches_{tag}.csv"), "w").write(open(os.path.join(G, f"launches_{tag}.csv")).read())
for src, dst in ((f"bench_{tag}.json", f"bench_r1_{tag}.json"), (f"bench_ref_{tag}.json", f"bench_ref_r1_{tag}.json")):
    line = open(os.path.join(G, src)).read().strip().splitlines()[-1]
    open(os.path.join(P, dst), "w").write(line + "\n")
rd = wr = None
for ln in summ.splitlines():
    f = ln.split()
    if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
        rd = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
    if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
        wr = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
t = json.load(open(os.path.join(P, "traffic.json")))
t["cfg3"].update({"k_scan_dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "source": f"profiles/r1_scan_fast_{tag}_ncu.txt"})
t["_how"] = t["_how"].replace(t["_how"][t["_how"].index("gpurun session"):t["_how"].index(",", t["_how"].index("gpurun session"))], f"gpurun session {tag}")
json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
d = json.loads(open(os.path.join(P, f"bench_r1_{tag}.json")).read())
r = d["roofline"]
print(f"value {d['value'] / 1e9:.3f} G reads/s, {d['ms_per_step']:.3f} ms/step, scan {r['achieved']:.0f} GB/s ({r['frac']:.3f}), path {r['path']['frac']:.3f}, "
      f"e2e {d['e2e']['value'] / 1e6:.1f} M reads/s, cpu {d['cpu_baseline']['value'] / 1e6:.2f} M (1 core) / {d['cpu_baseline']['allcores']['value'] / 1e6:.1f} M ({d['cpu_baseline']['allcores']['cores']} threads), "
      f"traffic {r['traffic']:.0f}, launches {d['gpu_launches']}, clocks {d['clocks']}")
