#!/bin/bash
# full ncu capture of the iso chain kernels (one frame)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iso_fast|conv_xy|occlusion" -s 8 -c 4 -o gpurun_out/prof_iso_s4 -f python scripts/exp_iso_e2e.py > gpurun_out/ncu_iso.log 2>&1
tail -3 gpurun_out/ncu_iso.log
ls -la gpurun_out/*.ncu-rep
