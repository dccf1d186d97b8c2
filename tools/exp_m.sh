# round 2, session m: batch sizes that load the scan's regions evenly (49 tile rows) against 3,000,000-read batches
source tools/gpu_exp.sh m
Q="--steps 10 --warmup 3 --no-e2e --no-cpu --no-verify-merge"
run balanced "" 
run r1_batches "" --batch-units 3000000 --also none
run balanced_serial "" --no-overlap --also none
