#!/bin/bash
# One `ncu --set full` capture each of the lane-group kernels that scan the short models (300.hmm: eight lanes per sequence,
# 100.hmm: four), inside the config-3 workload (100 k sequences).  Text exports into gpurun_out/r02/.
cd "$(dirname "$0")/.."
out=gpurun_out/r02
mkdir -p $out
for m in 300 100; do
  ncu --set full --import-source on --clock-control none -k regex:msv_scan_group_spec_kernel -s 2 -c 1 -o /tmp/ncu_short_$m \
      python tools/sweep_geometry.py --model $m.hmm --sequences 100000 --steps 2 --geometries default > $out/ncu_short_$m.log 2>&1
  ncu -i /tmp/ncu_short_$m.ncu-rep --page details > $out/ncu_short_${m}_details.txt 2>&1
  ncu -i /tmp/ncu_short_$m.ncu-rep --page raw --csv > $out/ncu_short_${m}_raw.csv 2>&1
  ncu -i /tmp/ncu_short_$m.ncu-rep --page source --csv > $out/ncu_short_${m}_source.csv 2>&1
done
