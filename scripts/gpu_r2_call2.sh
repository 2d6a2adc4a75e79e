#!/bin/bash
# round 2, session 2: the view-aligned copies / multi-frame launches on a B200: new tests, the experiment script, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_axis.py -x -q -s > gpurun_out/r02_gputests_axis.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_axis.log
tail -15 gpurun_out/r02_gputests_axis.log
timeout 600 python scripts/exp_axis.py > gpurun_out/r02_exp_axis_v2.txt 2>&1; tail -22 gpurun_out/r02_exp_axis_v2.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 > gpurun_out/r02_bench_axis_n1_s20.json 2> gpurun_out/r02_bench_axis_n1_s20.err
tail -c 400 gpurun_out/r02_bench_axis_n1_s20.err
python - <<'PY'
import json
for f in ("r02_bench_axis_n1_s20",):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("e2e_synchronous",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline_tex",{}).get("frac_issued"), d.get("gpu_launches"), d["roofline"]["kernel"])
    except Exception as e: print(f, "ERR", e)
PY
