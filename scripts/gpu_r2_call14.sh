#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02s2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests.log
tail -4 gpurun_out/r02s2_gputests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_bench_n1_s20_b.json 2> gpurun_out/r02s2_bench_n1_s20_b.err
timeout 900 python bench.py --workload iso --steps 72 --warmup 5 --no-cpu-baseline > gpurun_out/r02s2_bench_iso_n1_b.json 2> gpurun_out/r02s2_bench_iso_n1_b.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02s2_bench_n1_s20_b.json") if l.startswith("{")][-1])
print("sweep", d["value"], d["e2e"]["value"], d["e2e_synchronous"]["value"], d["e2e_synchronous"]["d2h_bytes_per_step"])
d=json.loads([l for l in open("gpurun_out/r02s2_bench_iso_n1_b.json") if l.startswith("{")][-1])
print("iso", d["value"], d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], {k:v for k,v in d.items() if "sync" in k})
PY
