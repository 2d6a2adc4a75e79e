#!/bin/bash
# Round verification on one B200 (run under gpurun): the GPU parity suite, smoke, both bench arms, and the ncu launch
# list of the bench command.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 24 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
grep -h '^{' gpurun_out/bench.log | cut -c1-2500; grep -h '^{' gpurun_out/bench_reference.log | cut -c1-600
