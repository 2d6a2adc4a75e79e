#!/bin/bash
# round 2, GPU call 1: state of the tree on a B200 (tests, the driver's bench invocation, reference arm, traffic capture)
mkdir -p gpurun_out
{ python -c "import pyopencl; print('pyopencl', pyopencl.VERSION_TEXT); print(pyopencl.get_platforms())" 2>&1;
  python -c "import gputools" 2>&1; ls /etc/OpenCL/vendors 2>&1; ldconfig -p | grep -i -E "opencl|pocl" ; nproc; free -g | head -2; } > gpurun_out/r02_probe_pyopencl.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gputests_call1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_call1.log
tail -5 gpurun_out/r02_gputests_call1.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_s20.json 2> gpurun_out/r02_bench_ref_s20.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err
tail -c 600 gpurun_out/r02_bench_n1_s20.err
timeout 600 python bench.py --gpus 1 --steps 720 --warmup 20 --no-c4 --no-cpu-baseline > gpurun_out/r02_bench_n1_s720.json 2> gpurun_out/r02_bench_n1_s720.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mip_ --csv --log-file gpurun_out/r02_traffic_sweep.csv python bench.py --steps 12 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r02_traffic_run.log 2>&1
python - <<'PY'
import json
for f in ("r02_bench_ref_s20","r02_bench_n1_s20","r02_bench_n1_s720"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("roofline_tex",{}).get("frac_issued"), d.get("roofline_tex",{}).get("frac_issued_at_render_footprint"))
        if "c4" in d: print(" c4", json.dumps(d["c4"])[:1500])
    except Exception as e: print(f, "ERR", e)
PY
