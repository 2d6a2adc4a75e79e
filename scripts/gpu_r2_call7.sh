#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r02s2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02s2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests.log
tail -6 gpurun_out/r02s2_gputests.log
