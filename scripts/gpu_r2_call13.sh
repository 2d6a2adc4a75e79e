#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sequence.py tests/test_keyframes.py tests/test_gpu_composite.py -q -x > gpurun_out/r02s2_gputests_iso.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests_iso.log
tail -8 gpurun_out/r02s2_gputests_iso.log
timeout 900 python bench.py --workload iso --steps 72 --warmup 5 > gpurun_out/r02s2_bench_iso_n1.json 2> gpurun_out/r02s2_bench_iso_n1.err; tail -c 300 gpurun_out/r02s2_bench_iso_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02s2_bench_iso_n1.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"], d.get("roofline",{}).get("traffic"))
PY
