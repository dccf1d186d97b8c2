# round 2, session q: per-thread prefetch.global.L2 of the tile after next (variants pft: on top of the bulk prefetch; pft0: instead of it -- source removed after the measurement; pf0 = -DSBR_PF_AHEAD=0: no prefetch at all)
source tools/gpu_exp.sh q
L=$PWD/sabreur_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic_slice or overlapped" 2>&1 | tail -2
run base_serial "" --also none --no-overlap
for v in pft pft0 pf0; do
run ${v}_serial "SABREUR_GPU_LIB=$L/lib_$v/libsabreur_gpu.so" --also none --no-overlap
done
run pft_overlap "SABREUR_GPU_LIB=$L/lib_pft/libsabreur_gpu.so" --also none
