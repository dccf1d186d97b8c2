#!/bin/bash
# One GPU session: parity tests, the bench (both arms), a launch list and one full capture of the scan kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh TAG [quick]
TAG=${1:-x}
MODE=${2:-full}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" ; tail -3 $O/pytest_gpu_$TAG.log
BENCH_ARGS=""; [ "$MODE" = "mid" ] && BENCH_ARGS="--no-e2e --no-cpu"
python bench.py $BENCH_ARGS > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_$TAG.json
if [ "$MODE" = "mid" ]; then
SMALL="--units-per-gpu 6000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_scan_fast -s 2 -c 1 -o $O/prof_scan_$TAG -f python bench.py $SMALL > $O/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
fi
if [ "$MODE" = "full" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-800 $O/bench_ref_$TAG.json
SMALL="--units-per-gpu 6000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_scan_fast -s 2 -c 1 -o $O/prof_scan_$TAG -f python bench.py $SMALL > $O/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_match|k_scatter|k_binscan' -s 6 -c 3 -o $O/prof_mp_$TAG -f python bench.py $SMALL > $O/ncu_mp_$TAG.log 2>&1; echo "ncu mp rc=$?"
fi
