# round 2, session g: ncu --set full captures of the big launches (profiles/r2_*_ncu.txt)
O=gpurun_out; TAG=g
SMALL="--units-per-gpu 6000000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify-merge --also none"
python bench.py $SMALL > $O/ncu_plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_fast -s 1 -c 1 -o $O/prof_scan_$TAG -f python bench.py $SMALL > $O/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
python bench.py $SMALL > $O/ncu_plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_match -s 1 -c 1 -o $O/prof_match_$TAG -f python bench.py $SMALL > $O/ncu_match_$TAG.log 2>&1; echo "ncu match rc=$?"
python bench.py $SMALL --workload cfg5 > $O/ncu_plain4_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_scan_fast|k_match' -s 2 -c 2 -o $O/prof_cfg5_$TAG -f python bench.py $SMALL --workload cfg5 > $O/ncu_cfg5_$TAG.log 2>&1; echo "ncu cfg5 rc=$?"
python bench.py $SMALL --workload cfg4 > $O/ncu_plain5_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_scan_fast|k_match' -s 2 -c 2 -o $O/prof_cfg4_$TAG -f python bench.py $SMALL --workload cfg4 > $O/ncu_cfg4_$TAG.log 2>&1; echo "ncu cfg4 rc=$?"
