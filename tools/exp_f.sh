# round 2, session f: staggered CTA starts (the SBR_SCAN_STAGGER_NS knob was removed after this measurement: no effect) + ncu launch list
source tools/gpu_exp.sh f
TAG=f
timeout 300 python -m pytest tests/test_shard_gpu.py -x -q 2>&1 | tail -3
run overlap "" --also none
run stagger500 "SBR_SCAN_STAGGER_NS=500" --also none
run stagger1000 "SBR_SCAN_STAGGER_NS=1000" --also none
run stagger2000 "SBR_SCAN_STAGGER_NS=2000" --also none
run serial_stagger1000 "SBR_SCAN_STAGGER_NS=1000" --also none --no-overlap
SMALL="--units-per-gpu 6000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
python bench.py $SMALL > $O/ncu_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
python bench.py $SMALL --also none > $O/ncu_plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_fast -s 3 -c 1 -o $O/prof_scan_$TAG -f python bench.py $SMALL --also none > $O/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
python bench.py $SMALL --also none > $O/ncu_plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_match -s 3 -c 1 -o $O/prof_match_$TAG -f python bench.py $SMALL --also none > $O/ncu_match_$TAG.log 2>&1; echo "ncu match rc=$?"
python bench.py $SMALL --workload cfg5 --also none > $O/ncu_plain4_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_scan_fast|k_match' -s 6 -c 2 -o $O/prof_cfg5_$TAG -f python bench.py $SMALL --workload cfg5 --also none > $O/ncu_cfg5_$TAG.log 2>&1; echo "ncu cfg5 rc=$?"
tail -2 $O/ncu_plain_$TAG.log | cut -c1-400
