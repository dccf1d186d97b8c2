# round 2, session c: occupancy variants of the scan (build them first: python -m sabreur_b200.build --variant wl160 -DSBR_WL_CAP=160; --variant wl160r40 -DSBR_WL_CAP=160 -DSBR_SCAN_NREG=40; --variant r40 -DSBR_SCAN_NREG=40)
source tools/gpu_exp.sh c
L=$PWD/sabreur_b200
run overlap "" --also none
run serial "" --also none --no-overlap
run wl160_serial "SABREUR_GPU_LIB=$L/lib_wl160/libsabreur_gpu.so" --also none --no-overlap
run wl160r40_serial "SABREUR_GPU_LIB=$L/lib_wl160r40/libsabreur_gpu.so" --also none --no-overlap
run wl160r40_overlap "SABREUR_GPU_LIB=$L/lib_wl160r40/libsabreur_gpu.so" --also none
run r40_serial "SABREUR_GPU_LIB=$L/lib_r40/libsabreur_gpu.so" --also none --no-overlap
run r40_overlap "SABREUR_GPU_LIB=$L/lib_r40/libsabreur_gpu.so" --also none
run r40_overlap_mp2 "SABREUR_GPU_LIB=$L/lib_r40/libsabreur_gpu.so SBR_MP_PER_SM=2" --also none
run overlap_slots2 "" --also none --slots 2
run overlap_slots4 "" --also none --slots 4
run cfg5_serial "" --workload cfg5 --also none --no-overlap
