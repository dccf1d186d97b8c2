#!/bin/bash
# Round-2 GPU session: parity tests, the full default bench (what the driver runs) + the reference arm, the one-stream
# variant beside it, a launch list and full captures of the two kernels, and the CLI's end-to-end rows.
# usage (on the GPU box, from the repo root): bash tools/gpu_r2.sh TAG [notests] [noncu] [nocli]
TAG=${1:-a}; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc
if [[ " $* " != *" notests "* ]]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-400 $O/bench_ref_$TAG.json
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
if [[ " $* " != *" noncu "* ]]; then
  # (each capture only after the same command line has just exited 0 without ncu; under ncu kernels are serialised, so the
  # launch list shows every kernel's own time, not the overlap)
  SMALL="--units-per-gpu 7000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
  python bench.py $SMALL > $O/ncu_plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  ONE="$SMALL --no-verify-merge --also none"
  python bench.py $ONE > $O/ncu_plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_scan_fast -s 1 -c 1 -o $O/prof_scan_$TAG -f python bench.py $ONE > $O/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
  python bench.py $ONE > $O/ncu_plain3_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_match -s 1 -c 1 -o $O/prof_match_$TAG -f python bench.py $ONE > $O/ncu_match_$TAG.log 2>&1; echo "ncu match rc=$?"
  python bench.py $ONE --workload cfg5 --units-per-gpu 9000000 > $O/ncu_plain4_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_scan_fast|k_match' -s 0 -c 2 -o $O/prof_cfg5_$TAG -f python bench.py $ONE --workload cfg5 --units-per-gpu 9000000 > $O/ncu_cfg5_$TAG.log 2>&1; echo "ncu cfg5 rc=$?"
  python bench.py $ONE --workload cfg4 > $O/ncu_plain5_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_scan_fast|k_match' -s 0 -c 2 -o $O/prof_cfg4_$TAG -f python bench.py $ONE --workload cfg4 > $O/ncu_cfg4_$TAG.log 2>&1; echo "ncu cfg4 rc=$?"
fi
if [[ " $* " != *" nocli "* ]]; then
  : > $O/e2e_cli_$TAG.jsonl
  for spec in "cfg3 6000000 plain none 0" "cfg3 6000000 gz none 0" "cfg3 4000000 bgzf gz 0" "cfg3 2000000 bz2 none 1" "cfg3 2000000 bz2 none 0" \
              "cfg3 6000000 xz none 1" "cfg3 6000000 xz none 0" "cfg3 6000000 zst none 1" "cfg3 6000000 zst none 0" "cfg5 6000000 zst zst 0" "cfg4 3000000 gz zst 0"; do
    set -- $spec
    timeout 600 python tools/e2e_cli.py --workload $1 --reads $2 --in-format $3 --out-format $4 --inflate-threads $5 >> $O/e2e_cli_$TAG.jsonl 2>> $O/e2e_cli_$TAG.err || echo "e2e $spec FAILED"
  done
  python - <<PY
import json
for ln in open("$O/e2e_cli_$TAG.jsonl"):
    d = json.loads(ln); p = d["pipeline_s"]
    print(d["workload"][:4], d["in_format"], "->", d["out_format"], "thr", d["inflate_threads"], "reads %d wall %.2f s  reader %.2f s (%.2f GB/s text)  gpu-wait %.2f  writer %.2f" % (
        d["reads"], d["wall_s"], p["reader_decompress"], d["reader_text_GBps"] or 0, p["gpu_thread_waiting"], p["writer_gather_compress_write"]))
PY
fi
