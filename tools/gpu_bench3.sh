#!/bin/bash
# short benches of the three workloads (no tests)
O=gpurun_out; mkdir -p $O
for w in ${@:-cfg3 cfg4 cfg5}; do
python bench.py --workload $w --units-per-gpu 30000000 --steps 5 --warmup 3 --no-e2e --no-cpu 2>$O/q_$w.err | tail -1 > $O/q_$w.json
python -c "
import json
d=json.loads(open('$O/q_$w.json').read()); r=d['roofline']; print('$w', round(d['value']/1e9,3), d['unit'], 'scan GB/s', round(r['achieved']), round(r['frac'],3), {k:round(v,2) for k,v in r['kernel_ms'].items()}, 'path', round(r['path']['frac'],3))"
done
