#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_composite.py tests/test_frames.py -m gpu -x -q > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_part.log
tail -25 gpurun_out/pytest_gpu_part.log
