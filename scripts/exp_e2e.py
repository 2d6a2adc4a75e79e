#!/usr/bin/env python
"""End-to-end frame rate of the synchronous render() call against the read-back strategy (C2 workload)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from spimagine_b200 import VolumeRenderer

vol = scenes.vol_g(512, np.uint16, seed=0)
cams = [scenes.gui_camera(2 * math.pi * f / 360, 4.0) for f in range(360)]
rend = VolumeRenderer((1024, 1024), pinned_outputs=True)
rend.set_data(vol)
rend.set_max_val(60000.)
rend.set_projection(cams[0][1])
lib, ctx = rend._lib, rend._ctx
ref = None
def run(label, n=360):
    global ref
    for i in range(5):
        rend.set_modelView(cams[i][0]); rend.render()
    t0 = time.perf_counter()
    for i in range(n):
        rend.set_modelView(cams[i][0]); rend.render()
    dt = time.perf_counter() - t0
    dev = []
    for i in range(0, 360, 30):
        rend.set_modelView(cams[i][0]); rend.render(); dev.append(rend.last_render_ms() * 1e3)
    t1 = time.perf_counter()
    for i in range(n):
        rend.set_modelView(cams[33][0]); rend.render()
    dt33 = (time.perf_counter() - t1) / n * 1e6
    rend.set_modelView(cams[33][0]); rend.render()
    k33 = rend.last_render_ms() * 1e3
    label += " [kernel %.0f us avg; frame 33: kernel %.0f us, e2e %.0f us]" % (sum(dev) / len(dev), k33, dt33)
    img = rend.output.copy()
    if ref is None: ref = img
    print("%-90s %.0f frames/s (%.1f us/frame) identical=%s" % (label, n / dt, 1e6 * dt / n, np.array_equal(img, ref)), flush=True)
lib.spv_set_tuning(ctx, 7, 2)
for mode in (0, 1):
    lib.spv_set_tuning(ctx, 8, mode)
    for b in (4, 6, 8, 10, 12, 14):
        lib.spv_set_tuning(ctx, 2, b); run("row order=%d copy streams=2 bands=%d" % (mode, b))
lib.spv_set_tuning(ctx, 7, 2); lib.spv_set_tuning(ctx, 8, 0); lib.spv_set_tuning(ctx, 2, 12)
# host-side cost alone
t0 = time.perf_counter()
for i in range(360): rend.set_modelView(cams[i][0])
print("set_modelView alone: %.1f us" % ((time.perf_counter() - t0) / 360 * 1e6))
t0 = time.perf_counter()
for i in range(360): rend.render_device_only()
rend.sync()
print("device-only launches: %.1f us/frame" % ((time.perf_counter() - t0) / 360 * 1e6))
for mode in (0, 1):
    lib.spv_set_tuning(ctx, 8, mode)
    t0 = time.perf_counter()
    n = 0
    for r in rend.render_sequence(cams[i][0] for i in range(360)): n += 1
    print("render_sequence (row order %d): %.1f us/frame" % (mode, (time.perf_counter() - t0) / 360 * 1e6))
