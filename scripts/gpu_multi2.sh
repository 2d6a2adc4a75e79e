#!/bin/bash
# multi-GPU regression (run under gpurun --gpus N): default sweep bench, slab bench (peer), iso bench (peer), timelapse
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29521 --steps 360 --warmup 20" > gpurun_out/bench_sweep_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_sweep_n$N.log
timeout 600 bash -c "$(declare -f run); N=$N; run 29522 --steps 60 --warmup 6 --workload slab --vol 1024 --img 1024 --composite peer" > gpurun_out/bench_slab_peer_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_slab_peer_n$N.log
timeout 600 bash -c "$(declare -f run); N=$N; run 29523 --steps 72 --warmup 6 --workload iso --vol 1024 --img 1024 --composite peer" > gpurun_out/bench_iso_peer_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_iso_peer_n$N.log
timeout 600 bash -c "$(declare -f run); N=$N; run 29524 --steps 8 --warmup 3 --workload timelapse --frames $((4*N)) --tl-shape 256,512,512" > gpurun_out/bench_tl_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_tl_n$N.log
for f in sweep slab_peer iso_peer tl; do echo "== $f"; grep -h '^{' gpurun_out/bench_${f}_n$N.log | cut -c1-420; tail -1 gpurun_out/bench_${f}_n$N.log; done
