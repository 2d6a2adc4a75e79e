#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
python scripts/exp_iso_e2e.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iso|conv_kernel|occlusion|shading" -s 14 -c 28 --csv --log-file gpurun_out/launches_iso.csv python scripts/exp_iso_e2e.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"iso_fast|occlusion" -s 6 -c 2 -o gpurun_out/prof_iso python scripts/exp_iso_e2e.py > /dev/null 2>&1
timeout 300 python bench.py --workload iso --vol 1024 --img 1024 --steps 72 --warmup 6 2>&1 | tail -1 | cut -c1-1200
