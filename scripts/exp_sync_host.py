#!/usr/bin/env python
"""Where does a blocking render() of configs[1] go (set_modelView + render, the reference's frame loop)?  Wall time per
frame, time inside each library call, and the Python around them."""
import os, sys, time, math, collections, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
import bench
from spimagine_b200 import VolumeRenderer, _lib

K = 360
rend = VolumeRenderer((1024, 1024), device=0, max_steps=200, pinned_outputs=True)
vol = scenes.vol_g(512, np.uint16, seed=0)
rend.set_data(vol)
rend.set_max_val(60000.)
cams = bench.sweep_cameras(K)
rend.set_projection(cams[0][1])
for i in range(20):
    rend.set_modelView(cams[i][0]); rend.render()

def loop():
    t0 = time.perf_counter()
    for i in range(K):
        rend.set_modelView(cams[i][0])
        rend.render()
    return (time.perf_counter() - t0) / K * 1e6
print("blocking render(): %.1f us per frame" % loop())
calls = collections.defaultdict(lambda: [0, 0.0])
lib = rend._lib
class Wrapped(object):
    def __init__(self, lib):
        object.__setattr__(self, "_l", lib)
    def __getattr__(self, name):
        f = getattr(self._l, name)
        def g(*a):
            t0 = time.perf_counter()
            r = f(*a)
            c = calls[name]; c[0] += 1; c[1] += time.perf_counter() - t0
            return r
        return g
rend._lib = Wrapped(lib)
us = loop()
rend._lib = lib
print("with wrapped calls: %.1f us per frame; inside library calls:" % us)
for name, (n, t) in sorted(calls.items(), key=lambda kv: -kv[1][1]):
    print("   %-28s %5d calls  %7.1f us per frame  (%.1f us per call)" % (name, n, t / K * 1e6, t / n * 1e6))
print("   sum %.1f us per frame" % (sum(t for _, t in calls.values()) / K * 1e6))
ms = []
for i in range(60):
    rend.set_modelView(cams[i * 6][0]); rend.render(); ms.append(rend.last_render_ms())
print("device time of the frame's launch (events): %.1f us" % (np.mean(ms) * 1e3))
pr = cProfile.Profile(); pr.enable(); loop(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(14); print(st.getvalue()[:3500])
