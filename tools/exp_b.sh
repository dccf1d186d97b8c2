source tools/gpu_exp.sh b
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_b.log
run overlap "" 
run serial "" --no-overlap
run ov_nomatchcarve "SBR_MATCH_NO_CARVEOUT=1" --also none
run ov_scancarve_def "SBR_SCAN_CARVEOUT=-1" --also none
run ov_nofuse "" --also none --no-fuse
run serial_nofuse "" --also none --no-fuse --no-overlap
