#!/usr/bin/env python
"""configs[1] through the library, by view angle: device time per frame of mip_fast_kernel, of mip_axis_kernel one frame
per launch, and of launches of 10 frames half a degree apart around the angle; then the read-back band count of the
multi-frame launches (tuning knob 2) against the end-to-end rate of a 20-frame and a 360-frame sequence."""
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
from spimagine_b200 import VolumeRenderer

vol = scenes.vol_g(512, np.uint16, seed=0)
rend = VolumeRenderer((1024, 1024), pinned_outputs=True)
rend.set_data(vol)
rend.set_max_val(60000.)
lib, ctx = rend._lib, rend._ctx
cam = lambda deg: scenes.gui_camera(math.radians(deg), 4.0)
rend.set_projection(cam(0)[1])
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)


def knob(k, v):
    assert lib.spv_set_tuning(ctx, k, v) == 0


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def one(deg):
    rend.set_modelView(cam(deg)[0])
    rend.render_device_only()


print("angle  hit rays   mip_fast   axis, 1 frame/launch   axis, 10 frames/launch (us per frame | Gsamples/s | of 1160)")
for deg in range(0, 360, 15):
    rend.enable_stats(True)
    knob(16, 0)
    one(deg)
    rend.sync()
    hits = rend.last_stats()[0]
    rend.enable_stats(False)
    t_fast = timed(lambda: one(deg), 10)
    knob(16, 1)
    t_one = timed(lambda: one(deg), 10)
    views = [cam(deg + 0.5 * (f - 5))[0] for f in range(10)]
    t_ten = timed(lambda: rend.render_batch(views, False), 5) / 10
    g = hits * 208 / t_ten * 1e-3
    print("%5d  %8d   %7.1f   %7.1f                 %7.1f | %4.0f | %.3f   %s" % (deg, hits, t_fast, t_one, t_ten, g, g / 1160., rend.mip_axis_used()),
          flush=True)

sweep20 = [cam(18. * i)[0] for i in range(20)]
sweep360 = [cam(float(i))[0] for i in range(360)]
for bands in (1, 2, 4, 8, 12, 16):
    knob(2, bands)
    res = []
    for views in (sweep20, sweep360):
        best = 0.
        for _ in range(3):
            t0 = time.perf_counter()
            n = sum(1 for _ in rend.render_sequence(views, batch=10))
            best = max(best, n / (time.perf_counter() - t0))
        res.append(best)
    print("read-back bands per launch %2d: render_sequence 20 frames %.0f frames/s, 360 frames %.0f frames/s" % (bands, res[0], res[1]), flush=True)
