#!/usr/bin/env python
"""View-aligned layered copies + several frames per launch (spv_mip_axis.cu) on configs[1]: correctness against the C
oracle on sampled rows and against mip_fast_kernel, then device time per frame by layer axis / lane map / frames per
launch, and end to end through render_sequence."""
import ctypes as C
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import VolumeRenderer, _lib
from oracle import oracle as orc

N = int(os.environ.get("EXP_VOL", 512))
IMG = int(os.environ.get("EXP_IMG", 1024))
vol = scenes.vol_g(N, np.uint16, seed=0)
peak = float(vol.max())
rend = VolumeRenderer((IMG, IMG), pinned_outputs=True)
rend.set_data(vol)
rend.set_max_val(peak)
lib, ctx = rend._lib, rend._ctx
cam = lambda deg: scenes.gui_camera(math.radians(deg), 4.0)
rend.set_projection(cam(0)[1])


def knob(k, v):
    assert lib.spv_set_tuning(ctx, k, v) == 0


# ---- 1. correctness -------------------------------------------------------------------------------------------------
o = orc.OracleRenderer((IMG, IMG), kind="port")
o.set_data(vol)
o.set_projection(cam(0)[1])
o.lib.so_set_row_sampling(0, 64)
rows = slice(0, IMG, 64)
for deg in (0, 40, 90, 200):
    M, P = cam(deg)
    rend.set_modelView(M)
    knob(16, 0)
    rend.render()
    ref, ref_a = rend.output.copy(), rend.output_alpha.copy()
    o.set_modelView(M)
    o.render(maxVal=peak)
    line = "angle %3d: max |axis - oracle| on every 64th row (mip_fast_kernel %.1e):" % (deg, float(np.abs(ref[rows] - o.output[rows]).max()))
    for forced in (None, (2, 0), (1, 1), (0, 2), (1, 0), (2, 1), (2, 2), (0, 0), (0, 1), (1, 2)):
        knob(16, 1 if forced is None else 10 + 3 * forced[0] + forced[1])
        rend.render()
        used = rend.mip_axis_used()
        d = float(np.abs(rend.output[rows] - o.output[rows]).max())
        same_a = bool(np.array_equal(rend.output_alpha, ref_a)) and bool(np.array_equal(rend.output_alpha[rows], o.output_alpha[rows]))
        line += "  %s->%s %.1e%s" % ("auto" if forced is None else "%d%d" % forced, "%d%d" % used, d, "" if same_a else " ALPHA DIFFERS")
        if forced == (2, 0):
            line += " (bits %s)" % ("equal" if np.array_equal(rend.output, ref) else "DIFFER")
    print(line, flush=True)
knob(16, 1)

# batch == single frames
degs = [18.0 * i for i in range(10)]
Ms = [cam(d)[0] for d in degs]
singles = []
for M in Ms:
    rend.set_modelView(M)
    rend.render()
    singles.append((rend.output.copy(), rend.output_alpha.copy(), rend.mip_axis_used()))
which = rend.render_batch(Ms, True)
frames = rend.batch_frames_of(which, copy=True)
print("batch of 10 vs single frames: max |diff| %.2e, alpha equal %s, choices %s" % (
    max(float(np.abs(f[0] - s[0]).max()) for f, s in zip(frames, singles)),
    all(np.array_equal(f[1], s[1]) for f, s in zip(frames, singles)), sorted(set(s[2] for s in singles))), flush=True)
seq = [r.output.copy() for r in rend.render_sequence(Ms)]
print("render_sequence (batched) vs single frames: max |diff| %.2e" % max(float(np.abs(a - s[0]).max()) for a, s in zip(seq, singles)))
seq1 = [r.output.copy() for r in rend.render_sequence(Ms, batch=1)]
print("render_sequence (batch=1) vs single frames: max |diff| %.2e" % max(float(np.abs(a - s[0]).max()) for a, s in zip(seq1, singles)))

# ---- 2. device time -------------------------------------------------------------------------------------------------
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)
params = rend._mip_params()
hits = {}
rend.enable_stats(True)
knob(16, 0)
for d in degs + [0.5 * i for i in range(16)]:
    rend.set_modelView(cam(d)[0])
    rend.render_device_only()
    rend.sync()
    hits[d] = rend.last_stats()[0]
rend.enable_stats(False)
knob(16, 1)


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps  # us


def single_sweep(dlist):
    for d in dlist:
        rend.set_modelView(cam(d)[0])
        rend.render_device_only()


def batched(dlist, F, to_host=False):
    for i in range(0, len(dlist), F):
        rend.render_batch([cam(d)[0] for d in dlist[i:i + F]], to_host)


sweep20 = [18.0 * i for i in range(20)]
for d in sweep20:
    if d not in hits:
        hits[d] = hits[d - 180.0] if (d - 180.0) in hits else None
rend.enable_stats(True)
knob(16, 0)
for d in sweep20:
    if hits.get(d) is None:
        rend.set_modelView(cam(d)[0]); rend.render_device_only(); rend.sync(); hits[d] = rend.last_stats()[0]
rend.enable_stats(False)
samples20 = sum(hits[d] for d in sweep20) * 208
for mode, name in ((0, "mip_fast_kernel"), (10 + 3 * 2 + 0, "axis kernel, z copy, 2x2 quads"), (1, "axis kernel, automatic")):
    knob(16, mode)
    t = timed(lambda: single_sweep(sweep20), 5) / 20
    print("one frame per launch, %-32s: %.1f us per frame, %.0f Gsamples/s" % (name, t, samples20 / 20 / t * 1e-3), flush=True)
knob(16, 1)
for F in (2, 4, 5, 10, 16):
    t = timed(lambda: batched(sweep20, F), 5) / 20
    print("%2d frames per launch: %.1f us per frame, %.0f Gsamples/s (%.3f of 1160)" % (F, t, samples20 / 20 / t * 1e-3,
                                                                                     samples20 / 20 / t * 1e-3 / 1160), flush=True)
print("launch choices over the sweep:", rend.mip_axis_used())
# overlapping two launches on two streams is not done: a launch's tail is 1 / F of what it was

# ---- 3. end to end --------------------------------------------------------------------------------------------------
sweep360 = [cam(float(d))[0] for d in range(360)]
for b in (1, 5, 10, 16):
    t0 = time.perf_counter()
    n = 0
    for r in rend.render_sequence(sweep360, batch=b):
        n += 1
    dt = time.perf_counter() - t0
    print("render_sequence batch=%2d: %.0f frames/s end to end (%d frames)" % (b, n / dt, n), flush=True)
