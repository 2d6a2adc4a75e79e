#!/bin/bash
# GPU pass: parity tests, smoke, bench (default, --skip, reference arm), ncu launch list, full ncu capture of the
# headline kernel, iso-surface timing on configs[2]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --skip --no-cpu-baseline > gpurun_out/bench_skip.log 2>&1
timeout 400 python bench.py --impl reference --steps 24 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mip_fast -s 30 -c 2 -o gpurun_out/prof_mip_zpair python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
EXP_FRAMES=36 timeout 600 python scripts/exp_iso.py > gpurun_out/exp_iso.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-3000; tail -1 gpurun_out/bench_skip.log | cut -c1-600; tail -1 gpurun_out/bench_ref.log | cut -c1-800; tail -5 gpurun_out/exp_iso.log
