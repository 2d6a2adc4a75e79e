#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): frame-sharded sweep bench and the sort-last slab bench, both composites
N=${1:-2}
VOL=${2:-1024}
IMG=${3:-1024}
STEPS=${4:-120}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29511 --steps 360 --warmup 20" > gpurun_out/bench_sweep_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_sweep_n$N.log
for comp in peer nccl; do
  timeout 900 bash -c "$(declare -f run); N=$N; run 29512 --steps $STEPS --warmup 10 --workload slab --vol $VOL --img $IMG --composite $comp" > gpurun_out/bench_slab_${comp}_v${VOL}_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_slab_${comp}_v${VOL}_n$N.log
done
timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 10 --workload slab --vol $VOL --img $IMG --composite nccl > gpurun_out/bench_slab_v${VOL}_n1.log 2>&1; echo "exit $?" >> gpurun_out/bench_slab_v${VOL}_n1.log
grep -h '^{' gpurun_out/bench_sweep_n$N.log | cut -c1-330
for f in gpurun_out/bench_slab_peer_v${VOL}_n$N.log gpurun_out/bench_slab_nccl_v${VOL}_n$N.log gpurun_out/bench_slab_v${VOL}_n1.log; do echo "== $f"; tail -3 $f | cut -c1-420; done
