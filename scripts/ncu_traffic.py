#!/usr/bin/env python
"""Turn an ncu launch list with DRAM byte counters into profiles/r02_mip_traffic.json, the record bench.py's
`roofline.traffic` is read from (keyed by workload; tied to the kernel sources by their SHA-1).

  on the GPU box:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:mip_ --csv --log-file gpurun_out/traffic.csv python bench.py --steps 12 --warmup 3 --no-c4 --no-cpu-baseline
  here:
    python scripts/ncu_traffic.py gpurun_out/traffic.csv sweep_512_1024 [family [frames per launch]]
With frames per launch (multi-frame launches of mip_axis_kernel: the grid's y extent) only those launches are averaged and
the key gets the suffix _b<frames>.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1., "ms": 1e3, "usecond": 1., "nsecond": 1e-3,
        "msecond": 1e3}


def main():
    path, key = sys.argv[1], sys.argv[2]
    family = sys.argv[3] if len(sys.argv) > 3 else "mip"
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        if frames and (("mip_axis" not in r["Kernel Name"]) or int(r["Grid Size"].strip("()").split(",")[1]) != frames):
            continue
        k = (r["ID"], r["Kernel Name"])
        per.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.)
    by_kernel = {}
    for (_, name), m in per.items():
        by_kernel.setdefault(name, []).append(m)
    name = max(by_kernel, key=lambda n: sum(m.get("gpu__time_duration.sum", 0.) for m in by_kernel[n]))
    ms = by_kernel[name]
    rec = {"kernel": name, "launches": len(ms),
           "dram_bytes_read": sum(m["dram__bytes_read.sum"] for m in ms) / len(ms),
           "dram_bytes_write": sum(m["dram__bytes_write.sum"] for m in ms) / len(ms),
           "time_us_under_ncu": sum(m["gpu__time_duration.sum"] for m in ms) / len(ms),
           "source_sha1": bench.source_sha1(family), "from": os.path.basename(path)}
    if frames:
        rec["frames_per_launch"] = frames
        key += "_b%d" % frames
    out = os.path.join(ROOT, "profiles", "r02_mip_traffic.json")
    allrec = {}
    if os.path.exists(out):
        with open(out) as f:
            allrec = json.load(f)
    allrec[key] = rec
    with open(out, "w") as f:
        json.dump(allrec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
