#!/bin/bash
# sort-last iso surface on N GPUs (N = number of GPUs of this box): 1024^3, 2048^3 and (N = 8) 4096^3
N=$1
for v in 1024 2048 $2; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload iso --vol $v --steps 36 --warmup 5 > gpurun_out/r02_bench_iso_v${v}_n${N}.json 2> gpurun_out/iso_n${N}.err
  tail -c 300 gpurun_out/iso_n${N}.err | grep -v OMP
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_iso_v${v}_n${N}.json") if l.startswith("{")][-1])
    print($v, $N, d["value"], d["ms_per_step"], d["e2e"]["value"], d["image_sha1_first8"])
    print({k:round(v["max_over_ranks_us"],1) for k,v in d["phases_us"].items()})
except Exception as e: print("ERR", e)
PY
done
