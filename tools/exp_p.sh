# round 2, session p: L2 prefetch distance of the scan (python -m sabreur_b200.build --variant pf$v -DSBR_PF_AHEAD=$v for v in 1 3 4; the default build has 2)
source tools/gpu_exp.sh p
L=$PWD/sabreur_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic_slice or fixture or overlapped or boundary" 2>&1 | tail -2
for v in 1 3 4; do
run pf${v}_serial "SABREUR_GPU_LIB=$L/lib_pf$v/libsabreur_gpu.so" --also none --no-overlap
run pf${v}_overlap "SABREUR_GPU_LIB=$L/lib_pf$v/libsabreur_gpu.so" --also none
done
run pf2_serial "" --no-overlap
run pf2_overlap ""
