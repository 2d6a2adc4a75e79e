#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/exp_e2e.py 2>&1 | tee gpurun_out/exp_e2e.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "sequence or golden or full_size or ingest" > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_part.log
tail -4 gpurun_out/pytest_gpu_part.log
