#!/bin/bash
# Round-2 GPU session: parity tests, the full default bench (what the driver runs), and the one-stream variant beside it.
# usage (on the GPU box, from the repo root): bash tools/gpu_r2.sh TAG [notests]
TAG=${1:-a}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
timeout 300 python bench.py --no-overlap --no-e2e --no-cpu > $O/bench_${TAG}_serial.json 2> $O/bench_${TAG}_serial.err; echo "bench serial rc=$?"; tail -3 $O/bench_${TAG}_serial.err
python - <<PY
import json
for f in ("$O/bench_$TAG.json", "$O/bench_${TAG}_serial.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    blocks = {"primary": d, **d.get("workloads", {})}
    for k, b in blocks.items():
        r = b["roofline"]; km = r["kernel_ms"]; nb = max(1, km.get("batches", 1))
        print(f.split("/")[-1], k, "value %.3f G/s" % (b["value"] / 1e9), "ms/step %.3f" % b["ms_per_step"],
              "scan frac %.3f" % r["frac"], "path frac %.3f" % r["path"]["frac"],
              "per batch us: scan %.1f match %.1f part %.1f" % (1e3 * km["scan"] / nb, 1e3 * km["match"] / nb, 1e3 * km["partition"] / nb),
              "clk", b["clocks"]["sm_mhz"], b["clocks"]["samples"], b["clocks"]["reasons"],
              "e2e %.3f G/s" % (b["e2e"]["value"] / 1e9) if b.get("e2e") else "", "self", b["self_check"]["spec_fail"], b["self_check"]["scan_finished"])
PY
