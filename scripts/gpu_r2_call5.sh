#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_call5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_call5.log
tail -12 gpurun_out/r02_gputests_call5.log
