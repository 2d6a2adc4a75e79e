#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_composite.py tests/test_gpu_sequence.py tests/test_gpu_configs.py -m gpu -x -q -k "iso or occ" 2>&1 | tail -4
ncu -k regex:"occ|conv_xy|iso_fast|shading" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 80 --csv --log-file gpurun_out/r02s3_iso_launches.csv python bench.py --workload iso --vol 1024 --img 1024 --steps 6 --warmup 2 --no-cpu-baseline --no-iso-overlap > gpurun_out/r02s3_iso_launches_run.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02s3_iso_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
t=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    t[r[ki][:50]][r[mi]].append(float(r[vi].replace(",","")))
for k,v in t.items():
    d=v["gpu__time_duration.sum"]; a=v["sm__cycles_active.avg"]; e=v["sm__cycles_elapsed.max"]
    print("%-52s n=%3d mean %.1f us  busy %.2f  dram %.1f MB" % (k, len(d), sum(d)/len(d)/1e3, sum(a)/max(sum(e),1), (sum(v["dram__bytes_read.sum"])+sum(v["dram__bytes_write.sum"]))/len(d)/1e6))
P
