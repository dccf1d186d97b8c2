# round 2, session h: branch-free compaction against the per-lane loops (variant source removed after the measurement: no gain; see profiles/r2_experiments.md and the git history of k_scan_fast.cu)
source tools/gpu_exp.sh h
L=$PWD/sabreur_b200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
run overlap "" 
run serial "" --no-overlap
run loops_overlap "SABREUR_GPU_LIB=$L/lib_loops/libsabreur_gpu.so" --also none
run loops_serial "SABREUR_GPU_LIB=$L/lib_loops/libsabreur_gpu.so" --also none --no-overlap
