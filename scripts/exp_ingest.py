#!/usr/bin/env python
"""Volume ingest timing for one timelapse frame of BASELINE configs[4] (512 x 1024 x 1024 uint16 = 1 GiB)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import VolumeRenderer, pinned_empty

shape = tuple(int(x) for x in os.environ.get("EXP_SHAPE", "512,1024,1024").split(","))
rng = np.random.default_rng(0)
a = scenes.vol_g(0, np.uint16, seed=0, shape=shape)
b = np.ascontiguousarray(a[::-1])
gib = a.nbytes / 2 ** 30
pa = pinned_empty(shape, np.uint16); pa[...] = a
pb = pinned_empty(shape, np.uint16); pb[...] = b
da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
M, P = scenes.gui_camera(0.5, 4.0)
for layout in ("zpair", "3d"):
    rend = VolumeRenderer((1024, 1024), pinned_outputs=True)
    rend.set_layout(layout)
    t0 = time.perf_counter(); rend.set_data(a); rend.sync(); t_first = time.perf_counter() - t0
    rend.set_modelView(M); rend.set_projection(P); rend.set_max_val(60000.)
    def timed(fn, n=4):
        ts = []
        for i in range(n):
            t0 = time.perf_counter(); fn(i); rend.sync(); ts.append(time.perf_counter() - t0)
        return min(ts)
    t_page = timed(lambda i: rend.update_data(b if i % 2 else a))
    t_pin = timed(lambda i: rend.update_data(pb if i % 2 else pa, pinned=True))
    t_dev = timed(lambda i: rend.set_data_device((db if i % 2 else da).data_ptr(), shape, np.uint16))
    rend.render(); ref = rend.output.copy()
    t_mm = timed(lambda i: rend.data_min_max, 1)
    def play(i):
        rend.update_data(pb if i % 2 else pa, pinned=True); rend.render()
    t_play = timed(play)
    def play_dev(i):
        rend.set_data_device((db if i % 2 else da).data_ptr(), shape, np.uint16); rend.render()
    t_play_dev = timed(play_dev)
    print("%s %.2f GiB: first set_data %.1f ms | update pageable %.1f ms (%.1f GB/s) | pinned async %.1f ms (%.1f GB/s) | "
          "from device %.2f ms | min/max (brick build) %.2f ms | upload+render+readback pinned %.1f ms, device %.2f ms" % (
              layout, gib, 1e3 * t_first, 1e3 * t_page, a.nbytes / t_page / 1e9, 1e3 * t_pin, a.nbytes / t_pin / 1e9,
              1e3 * t_dev, 1e3 * t_mm, 1e3 * t_play, 1e3 * t_play_dev), flush=True)
    rend.close()
