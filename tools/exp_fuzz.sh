# round 2: open-ended randomised hunts (kernels, then the command line) -- bash tools/exp_fuzz.sh SECONDS SEED
O=gpurun_out; mkdir -p $O
timeout $(( ${1:-240} + 200 )) python tests/fuzz_gpu.py --seconds ${1:-240} --seed ${2:-21} --out $O/fuzz 2>&1 | tail -8
timeout $(( ${1:-240} + 200 )) python tests/fuzz_cli.py --seconds ${1:-240} --seed ${2:-21} --out $O/fuzz_cli 2>&1 | tail -8
