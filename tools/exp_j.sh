# round 2, session j: ablation builds of the scan (python -m sabreur_b200.build --variant abl$v -DSBR_ABLATE=$v for v in 1 2 3)
source tools/gpu_exp.sh j
L=$PWD/sabreur_b200
export SBR_DEBUG_ABLATE=1
run full_serial "" --also none --no-overlap --no-verify-merge
for v in 1 2 3; do
run abl${v}_serial "SABREUR_GPU_LIB=$L/lib_abl$v/libsabreur_gpu.so" --also none --no-overlap --no-verify-merge
done
