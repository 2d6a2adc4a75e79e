#!/usr/bin/env python
"""Where the time of one synchronous iso_surface frame goes (configs[2]: 1024^3 uint16 -> 1024^2)."""
import ctypes as C, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes, bench
from spimagine_b200 import VolumeRenderer, _lib
N = int(os.environ.get("EXP_VOL", 1024)); W = 1024
vol = bench.vol_g_slab_device(N, 0, N, 1, torch.device("cuda", 0))
rend = VolumeRenderer((W, W), pinned_outputs=True)
rend.set_data_device(vol.data_ptr(), (N, N, N), np.uint16); rend.sync(); del vol
rend.set_max_val(30000.)
cams = [scenes.gui_camera(2 * math.pi * f / 36, 4.0) for f in range(36)]
rend.set_projection(cams[0][1])
p = _lib.IsoParams(rend._box(), 15000., 1., 200, .1, 21, 30, 0)
def t(fn, n=36):
    for i in range(3): fn(i)
    rend.sync(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    rend.sync(); return (time.perf_counter() - t0) / n * 1e6
def launch(i):
    rend.set_modelView(cams[i % 36][0]); rend._lib.spv_render_iso(rend._ctx, C.byref(p))
def launch_sync(i):
    launch(i); rend.sync()
def fetch(i):
    rend._fetch(7)
def full(i):
    rend.set_modelView(cams[i % 36][0]); rend.render(method="iso_surface")
print("enqueue only %.0f us | render+sync %.0f us | fetch 7 planes %.0f us | render() %.0f us" % (t(launch), t(launch_sync), t(fetch), t(full)))
for planes in (1, 2, 3, 7):
    print("fetch %d planes: %.0f us" % (planes, t(lambda i: rend._fetch(planes))))
