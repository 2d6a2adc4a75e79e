#!/usr/bin/env python
"""Render a few frames on one kernel path for an ncu capture: python scripts/exp_smem_ncu.py smem|tmu deg [deg ...]"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from spimagine_b200 import VolumeRenderer  # noqa: E402

path = sys.argv[1]
degs = [float(a) for a in sys.argv[2:]] or [0.]
vol = scenes.vol_g(512, np.uint16, seed=0)
r = VolumeRenderer((1024, 1024), max_steps=200)
r.set_data(vol)
r.set_max_val(60000.)
r.set_mip_path(path)
for deg in degs:
    M, P = scenes.gui_camera(math.radians(deg), 4.0)
    r.set_projection(P)
    r.set_modelView(M)
    for i in range(3):
        r.render_device_only()
    r.sync()
r.close()
