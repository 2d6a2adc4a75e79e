#!/bin/bash
# third session of round 2: the whole GPU suite, smoke, the driver's bench command, the reference arm, the iso and blur
# lines, and the ncu launch lists the committed traffic records are rebuilt from
mkdir -p gpurun_out
T=r02s3
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gputests.log
tail -4 gpurun_out/${T}_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref_s20.json 2> gpurun_out/${T}_bench_ref_s20.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1_s20.json 2> gpurun_out/${T}_bench_n1_s20.err
timeout 900 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline > gpurun_out/${T}_bench_n1_s720.json 2> gpurun_out/${T}_bench_n1_s720.err
timeout 900 python bench.py --workload iso --steps 72 --warmup 6 > gpurun_out/${T}_bench_iso_n1.json 2> gpurun_out/${T}_bench_iso_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mip_ --csv \
  --log-file gpurun_out/${T}_traffic_sweep_launches.csv python bench.py --steps 40 --warmup 10 --no-c4 --no-cpu-baseline > gpurun_out/${T}_traffic_sweep_run.log 2>&1
timeout 600 ncu -k regex:"occ|conv_xy|iso_fast|shading|rect_copy" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/${T}_traffic_iso_launches.csv python bench.py --workload iso --vol 1024 --img 1024 --steps 12 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_traffic_iso_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv \
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/${T}_bench_launches_run.log 2>&1
python - <<'PY'
import json
for f in ("r02s3_bench_ref_s20", "r02s3_bench_n1_s20", "r02s3_bench_n1_s720", "r02s3_bench_iso_n1"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("e2e_synchronous") or {}).get("value"),
              (d.get("roofline") or {}).get("frac"), (d.get("roofline_tex") or {}).get("frac_issued"), d.get("gpu_launches"),
              (d.get("roofline") or {}).get("traffic"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed:", e)
PY
