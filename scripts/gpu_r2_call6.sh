#!/bin/bash
# round 2, session 2: the driver's commands on the new default path + ncu evidence
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s2_bench_ref_s20.json 2> gpurun_out/r02s2_bench_ref_s20.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s2_bench_n1_s20.json 2> gpurun_out/r02s2_bench_n1_s20.err
tail -c 300 gpurun_out/r02s2_bench_n1_s20.err
timeout 600 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_bench_n1_s720.json 2> gpurun_out/r02s2_bench_n1_s720.err
timeout 600 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline --batch 1 > gpurun_out/r02s2_bench_n1_s720_b1.json 2> gpurun_out/r02s2_bench_n1_s720_b1.err
timeout 600 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline --no-axis > gpurun_out/r02s2_bench_n1_s720_noaxis.json 2> gpurun_out/r02s2_bench_n1_s720_noaxis.err
python - <<'PY'
import json
for f in ("r02s2_bench_ref_s20","r02s2_bench_n1_s20","r02s2_bench_n1_s720","r02s2_bench_n1_s720_b1","r02s2_bench_n1_s720_noaxis"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("e2e_synchronous",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline_tex",{}).get("frac_issued"), d.get("gpu_launches"), d.get("roofline",{}).get("kernel"))
    except Exception as e: print(f, "ERR", e)
PY
# ncu: DRAM bytes per launch of the bench command (sweep), of the iso workload; --set full of the multi-frame launch
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mip_ --csv --log-file gpurun_out/r02s2_traffic_sweep_b10.csv python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_traffic_b10_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:iso_fast --csv --log-file gpurun_out/r02s2_traffic_iso.csv python bench.py --workload iso --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r02s2_traffic_iso_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mip_axis -s 1 -c 2 -o gpurun_out/r02s2_mip_axis -f python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_ncu_axis.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s2_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_launches_run.log 2>&1
ls -la gpurun_out/r02s2*
