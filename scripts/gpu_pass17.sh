#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_part.log
tail -40 gpurun_out/pytest_gpu_part.log
timeout 200 python -m pytest tests -m gpu -x -q -k "iso_post or full_pipeline or sequence_equals" 2>&1 | tail -3
EXP_ISO_VARIANTS=4:1 timeout 300 python scripts/exp_iso.py 2>&1 | grep "full chain"
