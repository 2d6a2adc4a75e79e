#!/usr/bin/env python
"""Tuning experiments for the max-projection kernel on the C2 workload (512^3 uint16 -> 1024^2 sweep).
Prints frames/s for each CTA shape / option; results never change (checked against variant 0)."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import VolumeRenderer, _lib

N = int(os.environ.get("EXP_VOL", 512))
IMG = int(os.environ.get("EXP_IMG", 1024))
DT = {"u16": np.uint16, "f32": np.float32, "u8": np.uint8}[os.environ.get("EXP_DTYPE", "u16")]
vol = scenes.vol_g(N, DT, seed=0)
peak = float(vol.max())
rend = VolumeRenderer((IMG, IMG), pinned_outputs=True)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)
rend.set_data(vol)
rend.set_max_val(peak)
cams = [scenes.gui_camera(2 * math.pi * f / 360, 4.0) for f in range(360)]
rend.set_projection(cams[0][1])
mats = []
for M, P in cams:
    rend.set_modelView(M)
    mats.append((rend._invP.copy(), rend._invM.copy()))
lib, ctx = rend._lib, rend._ctx
params = _lib.MipParams(rend._box(), 0., peak, 1., 0., 1, 0, 200, 0)


def sweep(n=360):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(10):
        lib.spv_set_matrices(ctx, _lib.fp(mats[i][0]), _lib.fp(mats[i][1]))
        lib.spv_render_mip(ctx, C.byref(params))
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        lib.spv_set_matrices(ctx, _lib.fp(mats[i % 360][0]), _lib.fp(mats[i % 360][1]))
        rc = lib.spv_render_mip(ctx, C.byref(params))
        assert rc == 0
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3)


ref = None
for persistent in (0,):
    for variant in range(8):
        lib.spv_set_tuning(ctx, 0, variant)
        lib.spv_set_tuning(ctx, 1, persistent)
        rend.set_modelView(cams[40][0])
        rend.render()
        img = rend.output.copy()
        if ref is None:
            ref = img
        same = bool(np.array_equal(img, ref))
        fps = [sweep() for _ in range(3)]
        print("persistent %d variant %d: %s frames/s  identical=%s" % (persistent, variant, ["%.0f" % f for f in fps], same),
              flush=True)
print("tex probe: %.1f Gsamples/s" % (rend.texrate_probe(4000) / 1e9))
