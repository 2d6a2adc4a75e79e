#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "iso or sequence or smoke or post or golden" > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_part.log
tail -4 gpurun_out/pytest_gpu_part.log
timeout 300 python scripts/exp_e2e.py 2>&1 | tee gpurun_out/exp_e2e.txt
timeout 300 python scripts/exp_iso_e2e.py 2>&1 | tee gpurun_out/exp_iso_e2e.txt
