O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "larger_than_the_reserved or fixture or synthetic_slice" 2>&1 | tail -3
timeout 400 python tests/fuzz_gpu.py --seconds ${1:-240} --seed ${2:-21} --out $O/fuzz 2>&1 | tail -8
timeout 400 python tests/fuzz_cli.py --seconds ${1:-240} --seed ${2:-21} --out $O/fuzz_cli 2>&1 | tail -8
