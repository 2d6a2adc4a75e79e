#!/bin/bash
# blur workload: tests, bench line, per-variant timings, one ncu --set full capture of every kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_filters.py -m gpu -q > gpurun_out/pytest_filters.log 2>&1; tail -2 gpurun_out/pytest_filters.log
timeout 300 python bench.py --workload blur > gpurun_out/bench_blur.log 2>&1; tail -1 gpurun_out/bench_blur.log | cut -c1-900
timeout 300 python scripts/exp_blur.py > gpurun_out/exp_blur.txt 2>&1
EXP_BLUR_ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_" -c 10 -o gpurun_out/prof_blur -f python scripts/exp_blur.py > gpurun_out/ncu_blur.log 2>&1
ls -la gpurun_out/prof_blur.ncu-rep
