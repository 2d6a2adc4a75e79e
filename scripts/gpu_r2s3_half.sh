#!/bin/bash
# half-size first launch of batched sequences: tests, then the driver's 20 steps with and without it (three times each)
python -m pytest tests/test_gpu_axis.py tests/test_gpu_sequence.py tests/test_keyframes.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for half in True False; do
python - <<P
import sys, runpy, io, json, contextlib
import spimagine_b200.volumerender as v
v.VolumeRenderer.first_batch_half = $half
sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--no-c4", "--no-cpu-baseline"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    runpy.run_path("bench.py", run_name="__main__")
d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
print("first_batch_half=$half", "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "sync %.0f" % d["e2e_synchronous"]["value"])
P
done; done
