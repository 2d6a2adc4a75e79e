#!/bin/bash
mkdir -p gpurun_out
EXP_ISO_VARIANTS=0:1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iso_stream" -s 3 -c 1 -o gpurun_out/prof_iso_stream -f python scripts/exp_iso.py > gpurun_out/ncu_stream.log 2>&1
ls -la gpurun_out/prof_iso_stream.ncu-rep
