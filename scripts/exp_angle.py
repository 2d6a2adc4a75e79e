#!/usr/bin/env python
"""Frame time of the max-projection kernel against the view angle (C2 workload), both layouts."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from spimagine_b200 import VolumeRenderer, _lib

N = int(os.environ.get("EXP_VOL", 512))
IMG = int(os.environ.get("EXP_IMG", 1024))
vol = scenes.vol_g(N, np.uint16, seed=0)
for layout in ("zpair", "3d"):
    rend = VolumeRenderer((IMG, IMG), pinned_outputs=True)
    rend.set_layout(layout)
    rend.set_data(vol)
    rend.set_max_val(60000.)
    rend.enable_stats(True)
    for axis in ((0, 1, 0), (1, 0, 0), (0, 0, 1)):
        line = []
        for deg in range(0, 181, 15):
            from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_rotation, mat4_translate
            M = np.dot(mat4_translate(0, 0, -4.), mat4_rotation(math.radians(deg) + 1e-3, *axis))
            rend.set_projection(mat4_perspective(60, 1., .1, 10))
            rend.set_modelView(M)
            ts = []
            for rep in range(12):
                rend.render_device_only()
                rend.sync()
                ts.append(rend.last_render_ms())
            hits, issued = rend.last_stats()
            t = float(np.median(ts[2:]))
            line.append("%3d:%5.0fus/%4.0fG" % (deg, t * 1e3, issued / t / 1e6))
        print(layout, "axis", axis, " ".join(line), flush=True)
    rend.close()
