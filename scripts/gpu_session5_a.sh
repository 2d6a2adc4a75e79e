#!/bin/bash
# session 5, GPU call 1: new filter tests first, then the whole GPU suite, tuning experiments, blur bench + ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_filters.py -m gpu -q > gpurun_out/pytest_filters.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_filters.log
tail -15 gpurun_out/pytest_filters.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_filters.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/exp_mip.py 2>&1 | tee gpurun_out/exp_mip.txt | tail -12
timeout 300 python scripts/exp_e2e.py 2>&1 | tee gpurun_out/exp_e2e.txt | tail -16
timeout 300 python scripts/exp_blur.py 2>&1 | tee gpurun_out/exp_blur.txt | tail -14
timeout 400 python bench.py --workload blur > gpurun_out/bench_blur.log 2>&1; echo "exit $?" >> gpurun_out/bench_blur.log
tail -3 gpurun_out/bench_blur.log | cut -c1-3000
EXP_BLUR_ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_" -c 5 -o gpurun_out/prof_blur -f python scripts/exp_blur.py > gpurun_out/ncu_blur.log 2>&1
ls -la gpurun_out/*.ncu-rep
