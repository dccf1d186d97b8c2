# round 2, session l: the new tests + a final check of the tree
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tour.py tests/test_host_gpu.py -x -q 2>&1 | tail -3
