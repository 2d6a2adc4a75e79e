#!/bin/bash
# session 4, pass 1: parity suite after the iso traversal / fused blur / occlusion / device conversion changes,
# headline bench line, iso bench line and the iso launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_n1.json
timeout 300 python bench.py --workload iso --vol 1024 --img 1024 --steps 72 --warmup 6 > gpurun_out/bench_iso_n1.json 2> gpurun_out/bench_iso.err; echo "iso bench exit $?"; cut -c1-900 gpurun_out/bench_iso_n1.json
python scripts/exp_iso_e2e.py 2>&1 | tail -4 | tee gpurun_out/iso_timing.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iso|conv|occlusion|shading" -s 14 -c 20 --csv --log-file gpurun_out/launches_iso.csv python scripts/exp_iso_e2e.py > /dev/null 2>&1
tail -12 gpurun_out/launches_iso.csv | cut -c1-200
