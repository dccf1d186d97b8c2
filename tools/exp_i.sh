# round 2, session i: straight-line commit in the record evaluation (variant source removed after the measurement: no gain)
source tools/gpu_exp.sh i
L=$PWD/sabreur_b200
run old_serial "SABREUR_GPU_LIB=$L/lib_old/libsabreur_gpu.so" --no-overlap
run cl_serial "SABREUR_GPU_LIB=$L/lib_cl/libsabreur_gpu.so" --no-overlap
run old_overlap "SABREUR_GPU_LIB=$L/lib_old/libsabreur_gpu.so"
run cl_overlap "SABREUR_GPU_LIB=$L/lib_cl/libsabreur_gpu.so"
