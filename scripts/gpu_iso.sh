#!/bin/bash
N=${1:-2}; VOL=${2:-1024}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload iso --vol $VOL --img 1024 --steps 72 --warmup 6 > gpurun_out/iso_v${VOL}_n$N.log 2>&1; echo "exit $?" >> gpurun_out/iso_v${VOL}_n$N.log
timeout 600 python bench.py --gpus 1 --workload iso --vol $VOL --img 1024 --steps 72 --warmup 6 > gpurun_out/iso_v${VOL}_n1.log 2>&1; echo "exit $?" >> gpurun_out/iso_v${VOL}_n1.log
for f in gpurun_out/iso_v${VOL}_n$N.log gpurun_out/iso_v${VOL}_n1.log; do grep -h '^{' $f | cut -c1-1500; tail -2 $f | cut -c1-300; done
