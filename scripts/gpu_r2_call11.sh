#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_axis.py -q -x > gpurun_out/r02s2_gputests_float.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_gputests_float.log
tail -12 gpurun_out/r02s2_gputests_float.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --dtype f32 --vol 128 --img 512 > gpurun_out/r02s2_bench_c1_n1.json 2> gpurun_out/r02s2_bench_c1_n1.err; tail -c 300 gpurun_out/r02s2_bench_c1_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --dtype f32 --vol 128 --img 512 --no-axis --no-cpu-baseline > gpurun_out/r02s2_bench_c1_n1_noaxis.json 2> gpurun_out/r02s2_bench_c1_n1_noaxis.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --dtype f32 --vol 512 --img 1024 --no-cpu-baseline > gpurun_out/r02s2_bench_f32_512_n1.json 2> gpurun_out/r02s2_bench_f32_512_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c4 --dtype f32 --vol 512 --img 1024 --no-axis --no-cpu-baseline > gpurun_out/r02s2_bench_f32_512_n1_noaxis.json 2> gpurun_out/r02s2_bench_f32_512_n1_noaxis.err
python - <<'PY'
import json
for f in ("r02s2_bench_c1_n1","r02s2_bench_c1_n1_noaxis","r02s2_bench_f32_512_n1","r02s2_bench_f32_512_n1_noaxis"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("e2e_synchronous",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline_tex",{}).get("frac_issued"), d.get("gpu_launches"), d.get("roofline",{}).get("kernel"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
