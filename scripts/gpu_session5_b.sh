#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_filters.py -m gpu -q -x > gpurun_out/pytest_filters.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_filters.log
tail -6 gpurun_out/pytest_filters.log
timeout 300 python scripts/exp_blur.py 2>&1 | tee gpurun_out/exp_blur.txt | tail -14
timeout 300 python scripts/exp_e2e.py 2>&1 | tee gpurun_out/exp_e2e.txt | tail -18
EXP_BLUR_ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_" -c 5 -o gpurun_out/prof_blur -f python scripts/exp_blur.py > gpurun_out/ncu_blur.log 2>&1
ls -la gpurun_out/*.ncu-rep
