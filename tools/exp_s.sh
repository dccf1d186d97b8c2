# round 2, session s: FASTA records past the staged tile fetched in one go -- measured slower, source reverted (profiles/r2_experiments.md)
source tools/gpu_exp.sh s
Q="--steps 10 --warmup 3 --no-e2e --no-cpu --no-verify-merge"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_tour.py -x -q 2>&1 | tail -2
run cfg5_overlap "" --workload cfg5 --also none
run cfg5_serial "" --workload cfg5 --also none --no-overlap
