#!/bin/bash
# final pass: whole GPU suite, smoke, the driver's commands, traffic captures for the final sources, the alpha / float lines
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02s2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02s2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests.log
tail -4 gpurun_out/r02s2_gputests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s2_bench_ref_s20.json 2> gpurun_out/r02s2_bench_ref_s20.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mip_ --csv --log-file gpurun_out/r02s2_traffic_sweep_b10.csv python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_traffic_b10_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:iso_fast --csv --log-file gpurun_out/r02s2_traffic_iso.csv python bench.py --workload iso --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r02s2_traffic_iso_run.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s2_bench_n1_s20.json 2> gpurun_out/r02s2_bench_n1_s20.err
timeout 600 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline > gpurun_out/r02s2_bench_n1_s720.json 2> gpurun_out/r02s2_bench_n1_s720.err
python - <<'PY'
import json
for f in ("r02s2_bench_ref_s20","r02s2_bench_n1_s20","r02s2_bench_n1_s720"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("e2e_synchronous",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline_tex",{}).get("frac_issued"), d.get("gpu_launches"), d.get("roofline",{}).get("kernel"), d.get("roofline",{}).get("traffic"))
    except Exception as e: print(f, "ERR", e)
PY
