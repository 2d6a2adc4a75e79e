#!/bin/bash
# The other single-GPU bench workloads (run under gpurun after scripts/gpu_verify.sh): iso surface (configs[2]),
# the blur / spectrum processors, the keyframe record loop.  One JSON line each in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python bench.py --workload iso --steps 72 --warmup 6 2>&1 | grep '^{' > gpurun_out/bench_iso.json
timeout 300 python bench.py --workload blur 2>&1 | grep '^{' > gpurun_out/bench_blur.json
timeout 300 python bench.py --workload keyframes --steps 240 2>&1 | grep '^{' > gpurun_out/bench_keyframes.json
for f in iso blur keyframes; do cut -c1-300 gpurun_out/bench_$f.json; done
