#!/bin/bash
# BASELINE configs[3]: 2048^3 uint16 -> 2048^2 sort-last across N GPUs (run under gpurun --gpus N)
N=${1:-8}
VOL=${2:-2048}
IMG=${3:-2048}
STEPS=${4:-72}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
export -f run; export N
for cfg in "peer 1" "peer 2" "peer 4" "nccl 2"; do
  set -- $cfg
  f=gpurun_out/c4_${1}_k${2}_v${VOL}_n$N.log
  timeout 600 bash -c "run 2951$2 --steps $STEPS --warmup 8 --workload slab --vol $VOL --img $IMG --composite $1 --slabs-per-rank $2" > $f 2>&1; echo "exit $?" >> $f
  echo "== $f"; grep -h '^{' $f | cut -c1-260; grep -h '^{' $f | grep -o '"e2e": {"value": [0-9.]*'; grep -h '^{' $f | grep -o '"image_sha1_first8": "[0-9a-f]*"'; tail -1 $f
done
