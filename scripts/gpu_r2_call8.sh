#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_axis.py tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -k "axis or atten or alpha" > gpurun_out/r02s2_gputests_alpha.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests_alpha.log
tail -6 gpurun_out/r02s2_gputests_alpha.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --alpha-pow 1 > gpurun_out/r02s2_bench_alpha1_n1.json 2> gpurun_out/r02s2_bench_alpha1_n1.err; tail -c 300 gpurun_out/r02s2_bench_alpha1_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --alpha-pow 1 --batch 1 --no-cpu-baseline > gpurun_out/r02s2_bench_alpha1_n1_b1.json 2> gpurun_out/r02s2_bench_alpha1_n1_b1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --alpha-pow 1 --no-axis --no-cpu-baseline > gpurun_out/r02s2_bench_alpha1_n1_noaxis.json 2> gpurun_out/r02s2_bench_alpha1_n1_noaxis.err
python - <<'PY'
import json
for f in ("r02s2_bench_alpha1_n1","r02s2_bench_alpha1_n1_b1","r02s2_bench_alpha1_n1_noaxis"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("e2e_synchronous",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("gpu_launches"), d.get("roofline",{}).get("kernel"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
