#!/bin/bash
# ncu evidence for the round: launch list of the default bench command, full capture of the headline kernel and of the
# iso chain
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mip_fast -s 30 -c 2 -o gpurun_out/prof_mip_s4 -f python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iso_fast|conv_xy|occ_|shading" -s 12 -c 6 -o gpurun_out/prof_iso_s4 -f python scripts/exp_iso_e2e.py > gpurun_out/ncu_iso.log 2>&1
ls -la gpurun_out/*.ncu-rep
