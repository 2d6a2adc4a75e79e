#!/bin/bash
# multi-GPU iso surface (run under gpurun --gpus N): sort-last iso over peer memory vs NCCL
N=${1:-2}
VOL=${2:-1024}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
for comp in peer nccl; do
  timeout 600 bash -c "$(declare -f run); N=$N; run 29513 --steps 72 --warmup 6 --workload iso --vol $VOL --img 1024 --composite $comp" > gpurun_out/bench_iso_${comp}_v${VOL}_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_iso_${comp}_v${VOL}_n$N.log
  echo "== $comp"; grep -h '^{' gpurun_out/bench_iso_${comp}_v${VOL}_n$N.log | cut -c1-900; tail -2 gpurun_out/bench_iso_${comp}_v${VOL}_n$N.log | cut -c1-300
done
