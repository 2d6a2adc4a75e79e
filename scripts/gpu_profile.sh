#!/bin/bash
# ncu --set full captures (one GPU, never a multi-rank command): the headline max-projection kernel inside the bench
# command and the iso-surface chain.  Summarise with scripts/ncu_summary.py and commit the summaries under profiles/.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mip_fast -s 30 -c 2 -o gpurun_out/prof_mip -f \
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iso_fast|conv_xy|occ_|shading" -s 12 -c 6 \
  -o gpurun_out/prof_iso -f python scripts/exp_iso_e2e.py > gpurun_out/ncu_iso.log 2>&1
ls -la gpurun_out/*.ncu-rep
