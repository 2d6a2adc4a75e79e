#!/bin/bash
# third session: ncu --set full of the packed-FMA filter kernels and of the iso screen-space passes
mkdir -p gpurun_out
EXP_BLUR_ONCE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv_x2|conv_axisw" -c 6 -o gpurun_out/r02s3_prof_blur -f python scripts/exp_blur.py > gpurun_out/r02s3_ncu_blur.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"occ_tile|conv_xy|iso_fast" --launch-skip 12 -c 8 -o gpurun_out/r02s3_prof_iso -f python bench.py --workload iso --vol 1024 --img 1024 --steps 4 --warmup 2 --no-cpu-baseline --no-iso-overlap > gpurun_out/r02s3_ncu_iso.log 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/ncu_summary.py gpurun_out/r02s3_prof_blur.ncu-rep > gpurun_out/r02s3_blur_ncu_summary.json 2> gpurun_out/r02s3_ncu_sum.err
python scripts/ncu_summary.py gpurun_out/r02s3_prof_iso.ncu-rep > gpurun_out/r02s3_iso_ncu_summary.json 2>> gpurun_out/r02s3_ncu_sum.err
tail -3 gpurun_out/r02s3_ncu_sum.err; wc -c gpurun_out/r02s3_*_ncu_summary.json
