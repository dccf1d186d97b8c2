#!/bin/bash
# timing experiment: k_scan_fast with parts switched off (results are wrong on purpose)
O=gpurun_out; mkdir -p $O
SMALL="--units-per-gpu 6000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
for a in 0 1 3; do
  SBR_DEBUG_ABLATE=$a python bench.py $SMALL > $O/abl_plain_$a.log 2>&1 || { echo "plain run $a failed"; tail -5 $O/abl_plain_$a.log; continue; }
  SBR_DEBUG_ABLATE=$a ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_scan_fast -c 6 --csv --log-file $O/abl_$a.csv python bench.py $SMALL > $O/abl_ncu_$a.log 2>&1
  echo "ablate=$a"; grep k_scan_fast $O/abl_$a.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
