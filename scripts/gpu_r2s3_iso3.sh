#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_composite.py tests/test_gpu_sequence.py tests/test_gpu_configs.py -m gpu -x -q -k "iso or occ" 2>&1 | tail -4
timeout 300 python bench.py --workload iso --vol 1024 --img 1024 --steps 72 --warmup 6 --no-cpu-baseline > gpurun_out/r02s3_bench_iso_v2.json 2> gpurun_out/r02s3_bench_iso_v2.err
python - <<P
import json
d=[json.loads(l) for l in open("gpurun_out/r02s3_bench_iso_v2.json") if l.startswith("{")]
for x in d: print(x.get("value"), x.get("ms_per_step"), "e2e", x.get("e2e",{}).get("value"), "sync", x.get("e2e_synchronous",{}).get("value"), x.get("image_sha1_first8"), x.get("gpu_launches"))
P
tail -2 gpurun_out/r02s3_bench_iso_v2.err
