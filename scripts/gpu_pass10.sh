#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "iso or sort_last or smoke or post" > gpurun_out/pytest_gpu_iso.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_iso.log
tail -5 gpurun_out/pytest_gpu_iso.log
EXP_ISO_VARIANTS=4:1 timeout 300 python scripts/exp_iso.py 2>&1 | tee gpurun_out/exp_iso_variants.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"iso|conv|occ|shading" -s 18 -c 24 --csv --log-file gpurun_out/launches_iso.csv python scripts/exp_iso_e2e.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_iso.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
d=collections.defaultdict(list)
for r in rows[hdr+2:]:
    if len(r)>5: d[r[4][:50]].append(float(r[-1]))
for k,v in d.items(): print(k,len(v),sum(v)/len(v)/1000)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"occ_queue|iso_fast" -s 4 -c 2 -o gpurun_out/prof_occq -f python scripts/exp_iso_e2e.py > gpurun_out/ncu_occq.log 2>&1
