#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_composite.py tests/test_gpu_parity.py -x -q -k "composite or slab or several" > gpurun_out/pytest_comp.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_comp.log
tail -6 gpurun_out/pytest_comp.log
for k in 1 2 4; do
timeout 300 python bench.py --workload slab --vol 1024 --img 1024 --steps 60 --warmup 5 --composite peer --slabs-per-rank $k > gpurun_out/slab_peer_n1_k$k.log 2>&1; tail -1 gpurun_out/slab_peer_n1_k$k.log | cut -c100-330
done
timeout 300 python bench.py --workload slab --vol 1024 --img 1024 --steps 60 --warmup 5 --bricks 8 > gpurun_out/slab_bricks8_n1.log 2>&1; tail -1 gpurun_out/slab_bricks8_n1.log | cut -c1-330
