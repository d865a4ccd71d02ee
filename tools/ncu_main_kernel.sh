#!/bin/bash
# One `ncu --set full` capture of the headline scan kernel (msv_scan_warp_kernel on 1400.hmm x 1 M sequences) inside the bench
# command, exported as text (details, per-instruction source page with the warp-stall samples, raw metrics), plus the launch
# list of the whole bench process.  Writes into gpurun_out/r02/; profiles/roofline_traffic.json is refreshed from the raw page
# by tools/update_roofline_traffic.py.  Run AFTER the same bench command has exited 0 without the profiler.
cd "$(dirname "$0")/.."
out=gpurun_out/r02
mkdir -p $out
ncu --set full --import-source on --clock-control none -k regex:msv_scan_warp_kernel -s 2 -c 1 -o /tmp/ncu_main \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_main.log 2>&1
ncu -i /tmp/ncu_main.ncu-rep --page details > $out/ncu_main_details.txt 2>&1
ncu -i /tmp/ncu_main.ncu-rep --page raw --csv > $out/ncu_main_raw.csv 2>&1
ncu -i /tmp/ncu_main.ncu-rep --page source --csv > $out/ncu_main_source.csv 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_launches.log 2>&1
python tools/update_roofline_traffic.py $out/ncu_main_raw.csv
