#!/bin/bash
# second session of round 2: the driver's command under torchrun on N GPUs (sweep line + c4 record)
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02s2_bench_n${N}_s20.json 2> gpurun_out/r02s2_bench_n${N}_s20.err; echo "exit $?"
tail -c 600 gpurun_out/r02s2_bench_n${N}_s20.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02s2_bench_n${N}_s20.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_tex"]["frac_issued"], d["gpu_launches"])
print(json.dumps(d.get("c4"))[:1800])
PY
