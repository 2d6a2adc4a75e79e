#!/bin/bash
# third session of round 2: the table-driven occlusion pass (knob 17) -- parity tests, then configs[2] with and without it
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_composite.py tests/test_gpu_sequence.py -m gpu -x -q -k "iso or occ" 2>&1 | tail -5
for v in "" "--no-occ-table"; do
  tag=$( [ -z "$v" ] && echo table || echo hashed )
  timeout 300 python bench.py --workload iso --vol 1024 --img 1024 --steps 72 --warmup 6 --no-cpu-baseline $v > gpurun_out/r02s3_bench_iso_${tag}.json 2> gpurun_out/r02s3_bench_iso_${tag}.err
  python - <<P
import json
d=[json.loads(l) for l in open("gpurun_out/r02s3_bench_iso_${tag}.json") if l.startswith("{")]
for x in d: print("${tag}", x.get("value"), x.get("ms_per_step"), "e2e", x.get("e2e",{}).get("value"), "sync", x.get("e2e_synchronous",{}).get("value"), x.get("image_sha1_first8"))
P
  tail -2 gpurun_out/r02s3_bench_iso_${tag}.err
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02s3_iso_launches.csv python bench.py --workload iso --vol 1024 --img 1024 --steps 6 --warmup 2 --no-cpu-baseline --no-iso-overlap > gpurun_out/r02s3_iso_launches_run.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02s3_iso_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
t=collections.defaultdict(list)
for r in rows[1:]:
    if r[mi]=="gpu__time_duration.sum": t[r[ki][:60]].append(float(r[vi].replace(",","")))
for k,v in t.items(): print("%-62s n=%3d mean %.1f us" % (k, len(v), sum(v)/len(v)/1e3))
P
