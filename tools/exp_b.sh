# round 2, session b: the two-stream pipeline with and without the 228 KB shared-memory split (profiles/r2_experiments.md)
source tools/gpu_exp.sh b
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_b.log
run overlap "" 
run serial "" --no-overlap
run ov_nomatchcarve "SBR_MATCH_NO_CARVEOUT=1" --also none
run ov_scancarve_def "SBR_SCAN_CARVEOUT=-1" --also none
run ov_nofuse "" --also none --no-fuse
run serial_nofuse "" --also none --no-fuse --no-overlap
