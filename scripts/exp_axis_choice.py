#!/usr/bin/env python
"""Does choose_axis (spv_api.cu) pick the fastest copy / lane map?  configs[1]'s volume and image, camera families other
than the spin about y: every forced (axis, lane map) against the automatic choice, 10 frames per launch."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import VolumeRenderer
from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate

vol = scenes.vol_g(512, np.uint16, seed=0)
rend = VolumeRenderer((1024, 1024), pinned_outputs=True)
rend.set_data(vol)
rend.set_max_val(60000.)
lib, ctx = rend._lib, rend._ctx
rend.set_projection(scenes.gui_camera(0, 4.0)[1])
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)
T = mat4_translate(0, 0, -4.)


def knob(k, v):
    assert lib.spv_set_tuning(ctx, k, v) == 0


def timed(fn, reps=4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


families = {
    "spin about x": lambda t: np.dot(T, mat4_rotation(t, 1., 0., 0.)),
    "spin about z (roll)": lambda t: np.dot(T, mat4_rotation(t, 0., 0., 1.)),
    "spin about y, camera rolled 30 deg": lambda t: np.dot(T, np.dot(mat4_rotation(math.radians(30), 0., 0., 1.), mat4_rotation(t, 0., 1., 0.))),
    "spin about y, camera rolled 45 deg": lambda t: np.dot(T, np.dot(mat4_rotation(math.radians(45), 0., 0., 1.), mat4_rotation(t, 0., 1., 0.))),
    "spin about y, tilted 35 deg about x": lambda t: np.dot(T, np.dot(mat4_rotation(math.radians(35), 1., 0., 0.), mat4_rotation(t, 0., 1., 0.))),
    "spin about (1,1,1)": lambda t: np.dot(T, mat4_rotation(t, 1., 1., 1.)),
}
for name, fam in families.items():
    print(name)
    for deg in (10, 40, 70, 100):
        views = [fam(math.radians(deg + 0.5 * f)) for f in range(10)]
        res = {}
        for axis in range(3):
            for quad in range(3):
                knob(16, 10 + 3 * axis + quad)
                res[(axis, quad)] = timed(lambda: rend.render_batch(views, False)) / 10
        knob(16, 1)
        t_auto = timed(lambda: rend.render_batch(views, False)) / 10
        pick = rend.mip_axis_used()
        best = min(res, key=res.get)
        knob(16, 0)

        def single():
            for M in views:
                rend.set_modelView(M)
                rend.render_device_only()
        t_fast = timed(single, 2) / 10
        print("  %3d deg: automatic %s %.1f us | best %s %.1f us | worst %.1f us | mip_fast_kernel %.1f us   %s" % (
            deg, pick, t_auto, best, res[best], max(res.values()), t_fast,
            " ".join("%d%d:%.0f" % (k[0], k[1], v) for k, v in sorted(res.items()))), flush=True)
