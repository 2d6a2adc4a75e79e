#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_axis_angles.py > gpurun_out/r02s2_exp_axis_angles.txt 2>&1; tail -34 gpurun_out/r02s2_exp_axis_angles.txt
timeout 600 python bench.py --workload keyframes --no-cpu-baseline > gpurun_out/r02s2_bench_keyframes_n1.json 2> gpurun_out/r02s2_bench_keyframes_n1.err; tail -c 300 gpurun_out/r02s2_bench_keyframes_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02s2_bench_keyframes_n1.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["synchronous_value"], d["e2e"]["checksums_agree"], d["gpu_launches"], d["steps"])
PY
