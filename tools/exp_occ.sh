#!/bin/bash
O=gpurun_out
for pad in 0 20000 58000; do
  echo "pad=$pad"; SBR_SCAN_PAD_SMEM=$pad python bench.py --units-per-gpu 30000000 --steps 5 --warmup 2 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['achieved'], d['roofline']['kernel_ms'])"
done
