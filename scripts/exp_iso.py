#!/usr/bin/env python
"""Iso-surface timing on BASELINE configs[2]: 1024^3 uint16 -> 1024x1024, iso at maxVal/2, AO defaults."""
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
import bench
from spimagine_b200 import VolumeRenderer, _lib

N = int(os.environ.get("EXP_VOL", 1024))
IMG = int(os.environ.get("EXP_IMG", 1024))
NF = int(os.environ.get("EXP_FRAMES", 36))
dev = torch.device("cuda", 0)
vol = bench.vol_g_slab_device(N, 0, N, 1, dev)
rend = VolumeRenderer((IMG, IMG), pinned_outputs=True)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)
rend.set_data_device(vol.data_ptr(), (N, N, N), np.uint16)
rend.sync()
del vol
torch.cuda.empty_cache()
print("min/max", rend.data_min_max)
maxVal = 30000.
rend.set_max_val(maxVal)
cams = [scenes.gui_camera(2 * math.pi * f / NF, 4.0) for f in range(NF)]
rend.set_projection(cams[0][1])
mats = []
for M, P in cams:
    rend.set_modelView(M)
    mats.append((rend._invP.copy(), rend._invM.copy()))
lib, ctx = rend._lib, rend._ctx
import hashlib
for variant in os.environ.get("EXP_ISO_VARIANTS", "4:1").split(","):
    segments, centre = (int(v) for v in variant.split(":")[:2])
    lib.spv_set_tuning(ctx, 4, segments)
    lib.spv_set_tuning(ctx, 5, centre)
    lib.spv_set_tuning(ctx, 6, int(os.environ.get("EXP_OCC_CTAS", "5")))
    print("iso search: %d segment(s) per ray, centre-out order %d" % (segments, centre))
    for flags, name in ((_lib.ISO_RAW_ONLY, "iso_surface kernel alone"), (0, "full chain (5 launches)")):
        p = _lib.IsoParams(rend._box(), maxVal / 2, 1., 200, .1, 21, 30, flags)
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(NF):
                lib.spv_set_matrices(ctx, _lib.fp(mats[i][0]), _lib.fp(mats[i][1]))
                rc = lib.spv_render_iso(ctx, C.byref(p))
                assert rc == 0
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / NF
        print("%s: %.3f ms/frame  %.0f frames/s" % (name, ms, 1e3 / ms), flush=True)
    rend.set_modelView(cams[3][0])
    rend.render(method="iso_surface")
    print("    image sha1", hashlib.sha1(rend.output.tobytes() + rend.output_depth.tobytes() +
                                         rend.output_normals.tobytes()).hexdigest()[:12])
rend.enable_stats(True)
for split in (1, 2, 4):
    lib.spv_set_tuning(ctx, 4, split)
    for f in (3, 12):
        rend.set_modelView(cams[f][0])
        rend.render(method="iso_surface")
        mx, sm = rend.last_warp_cycles()
        nw = (IMG // 8) * (IMG // 4) * split
        print("segments %d frame %2d: longest warp %d cycles (%.1f us at 1.9 GHz): %d search iterations, %d lockstep skip "
              "steps; mean warp %.0f cycles" % (split, f, mx >> 32, (mx >> 32) / 1.9e3, (mx >> 16) & 0xffff, mx & 0xffff,
                                                sm / nw))
        print("    warps by log2(cycles):", {b: c for b, c in enumerate(rend.last_warp_histogram) if c})
rend.set_modelView(cams[3][0])
rend.render(method="iso_surface")
print("hit rays, samples:", rend.last_stats(), " hit pixels:", int(np.isfinite(rend.output_depth).sum()))
