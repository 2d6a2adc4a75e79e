#!/usr/bin/env python
"""Experiment: box / ring geometries of the software-sampled max projection (tuning knob 11) and hybrid shares (knob 10)."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import scenes
from spimagine_b200 import VolumeRenderer
vol = scenes.vol_g(512, np.uint16, seed=0)
r = VolumeRenderer((1024, 1024), max_steps=200)
r.set_data(vol); r.set_max_val(60000.)
def timeit():
    ts = []
    for i in range(10):
        r.render_device_only(); r.sync(); ts.append(r.last_render_ms() * 1e3)
    return float(np.median(ts[2:]))
CFGS = [int(x) for x in os.environ.get("CFGS", "0,1,2,3,4").split(",")]
HYB = [int(x) for x in os.environ.get("HYB", "0,4").split(",")]
print("angle  tmu_us | cfg: [sw_share] us per tex_of8 in %s" % HYB)
for deg in (0, 30, 45, 90):
    M, P = scenes.gui_camera(math.radians(deg), 4.0)
    r.set_projection(P); r.set_modelView(M)
    r.set_mip_path("tmu"); r.render(); ref = r.output.copy()
    line = "%5d  %6.1f |" % (deg, timeit())
    r.set_mip_path("smem")
    for cfg in CFGS:
        r._check(r._lib.spv_set_tuning(r._ctx, 11, cfg))
        r._check(r._lib.spv_set_tuning(r._ctx, 10, 0))
        r.enable_stats(True); r.render()
        v = (C.c_ulonglong * 4)(); r._check(r._lib.spv_last_stats(r._ctx, v, 4)); r.enable_stats(False)
        assert np.abs(r.output - ref).max() < 2e-4, (cfg, np.abs(r.output - ref).max())
        line += "  %d: [%.3f]" % (cfg, v[2] / max(1, v[1]))
        for n in HYB:
            r._check(r._lib.spv_set_tuning(r._ctx, 10, n))
            line += " %.1f" % timeit()
    print(line, flush=True)
r.close()
