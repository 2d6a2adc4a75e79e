#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "display or conversion or sequence" > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_part.log
tail -12 gpurun_out/pytest_gpu_part.log
