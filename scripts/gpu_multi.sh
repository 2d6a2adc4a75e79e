#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): frame-sharded sweep bench and the sort-last slab bench
N=${1:-2}
VOL=${2:-1024}
IMG=${3:-1024}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 360 --warmup 20 > gpurun_out/bench_sweep_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_sweep_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 120 --warmup 10 --workload slab --vol $VOL --img $IMG > gpurun_out/bench_slab_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_slab_n$N.log
timeout 900 python bench.py --gpus 1 --steps 120 --warmup 10 --workload slab --vol $VOL --img $IMG > gpurun_out/bench_slab_n1_v$VOL.log 2>&1; echo "exit $?" >> gpurun_out/bench_slab_n1_v$VOL.log
tail -2 gpurun_out/bench_sweep_n$N.log | cut -c1-700; tail -2 gpurun_out/bench_slab_n$N.log | cut -c1-1800; tail -2 gpurun_out/bench_slab_n1_v$VOL.log | cut -c1-900
