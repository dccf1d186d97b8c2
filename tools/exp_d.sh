# round 2, session d: the lean match kernel / cfg5 beside the scan
source tools/gpu_exp.sh d
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_scale.py -x -q > $O/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_d.log
run overlap "" 
run cfg5_overlap_nofuse "" --workload cfg5 --also none --no-fuse
run cfg5_serial "" --workload cfg5 --also none --no-overlap
