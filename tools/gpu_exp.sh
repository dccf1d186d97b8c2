#!/bin/bash
# quick A/B runs of bench.py variants (HBM-resident loop only): bash tools/gpu_exp.sh TAG
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
Q=${Q:-"--units-per-gpu 30000000 --steps 10 --warmup 3 --no-e2e --no-cpu"}
run() {  # label, env assignments (may be empty), bench args
  local label=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py $Q "$@" > $O/exp_${TAG}_$label.json 2> $O/exp_${TAG}_$label.err || { echo "$label FAILED"; tail -3 $O/exp_${TAG}_$label.err; }
  python - "$O/exp_${TAG}_$label.json" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[2], "unreadable", e); sys.exit(0)
for k, b in {"primary": d, **d.get("workloads", {})}.items():
    r = b["roofline"]; km = r["kernel_ms"]; nb = max(1, km.get("batches", 1))
    print("%-22s %-8s %.3f G/s  ms/step %.3f  scan frac %.3f  path frac %.3f  per batch us: scan %.1f match %.1f part %.1f  clk %s/%s" % (
        sys.argv[2], k, b["value"] / 1e9, b["ms_per_step"], r["frac"], r["path"]["frac"], 1e3 * km["scan"] / nb, 1e3 * km["match"] / nb,
        1e3 * km["partition"] / nb, b["clocks"]["sm_mhz"], b["clocks"]["samples"]))
PY
}
