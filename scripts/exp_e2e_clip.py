#!/usr/bin/env python
"""Synchronous render() and render_sequence() on the C2 workload against the band count, with the rows the box cannot
touch left out of the read-back (knob 9) and without."""
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


def sync_rate(n=360):
    for i in range(5):
        rend.set_modelView(cams[i][0]); rend.render()
    t0 = time.perf_counter()
    for i in range(n):
        rend.set_modelView(cams[i][0]); rend.render()
    return (time.perf_counter() - t0) / n * 1e6


def seq_rate(n=360):
    for _ in rend.render_sequence(cams[i][0] for i in range(10)):
        pass
    t0 = time.perf_counter()
    for _ in rend.render_sequence(cams[i][0] for i in range(n)):
        pass
    return (time.perf_counter() - t0) / n * 1e6


for clip in (0, 1):
    lib.spv_set_tuning(ctx, 9, clip)
    for b in (8, 12, 16, 20, 24, 32):
        lib.spv_set_tuning(ctx, 2, b)
        print("clip %d bands %2d: synchronous %.1f us/frame" % (clip, b, sync_rate()), flush=True)
    lib.spv_set_tuning(ctx, 2, 12)
    print("clip %d: render_sequence %.1f us/frame" % (clip, seq_rate()), flush=True)
