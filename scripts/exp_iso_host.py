#!/usr/bin/env python
"""Where does an iso-surface frame of render_sequence go?  configs[2] (1024^3 uint16 -> 1024^2): wall time per frame
(a) as bench.py's e2e line, (b) without read-backs (renders only, device-resident), (c) host time of the issue calls
alone (time spent inside the generator between yields, measured per call through wrapped library functions)."""
import os, sys, time, math, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
import bench
from spimagine_b200 import VolumeRenderer, _lib

N, W, K = int(os.environ.get("EXP_VOL", 1024)), 1024, int(os.environ.get("EXP_FRAMES", 144))
dev = torch.device("cuda", 0)
rend = VolumeRenderer((W, W), device=0, max_steps=200, pinned_outputs=True)
stream = torch.cuda.Stream(device=0)
torch.cuda.set_stream(stream)
rend.use_stream(stream.cuda_stream)
vol = bench.vol_g_slab_device(N, 0, N, 1, dev)
rend.set_data_device(vol.data_ptr(), (N, N, N), np.uint16)
rend.sync()
rend.set_max_val(30000.)
NF = 36
cams = [scenes.gui_camera(2 * math.pi * f / NF, 4.0) for f in range(NF)]
rend.set_projection(cams[0][1])


def seq(n, planes=2):
    t0 = time.perf_counter()
    for r_ in rend.render_sequence((cams[i % NF][0] for i in range(n)), method="iso_surface", iso_planes=planes):
        pass
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


seq(12)
for stage in (0, 1, 0, 1):
    _lib.check(rend._lib.spv_set_tuning(rend._ctx, 18, stage), rend._ctx)
    seq(12)
    print("render_sequence, output + alpha, device staging %d:  %.1f us per frame" % (stage, seq(K)), flush=True)
print("render_sequence, all 7 planes:    %.1f us per frame" % seq(K, 7), flush=True)

# per-call host time: wrap the library functions the generator calls
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
t_py0 = time.perf_counter()
us = seq(K)
total = time.perf_counter() - t_py0
rend._lib = lib
print("with wrapped calls: %.1f us per frame; time inside library calls per frame:" % us)
for name, (n, t) in sorted(calls.items(), key=lambda kv: -kv[1][1]):
    print("   %-28s %5d calls  %7.1f us per frame  (%.1f us per call)" % (name, n, t / K * 1e6, t / n * 1e6))
print("   sum %.1f us per frame" % (sum(t for _, t in calls.values()) / K * 1e6))

# host cost of set_modelView alone
t0 = time.perf_counter()
for i in range(2000):
    rend.set_modelView(cams[i % NF][0])
print("set_modelView: %.1f us per call" % ((time.perf_counter() - t0) / 2000 * 1e6))

# device-only frames (no read-back), as bench.py's `value`
_lib.check(lib.spv_set_tuning(rend._ctx, 14, 1), rend._ctx)
import ctypes as C
def dev_frames(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p = _lib.IsoParams(rend._box(), 15000., 1., 200, .1, 21, 30, 0)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        _lib.check(lib.spv_select_slot(rend._ctx, i & 1), rend._ctx)
        rend.set_modelView(cams[i % NF][0])
        _lib.check(lib.spv_render_iso(rend._ctx, C.byref(p)), rend._ctx)
    t_issue = (time.perf_counter() - t0) / n * 1e6
    _lib.check(lib.spv_stream_join(rend._ctx), rend._ctx)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, t_issue
dev_frames(12)
d, h = dev_frames(K)
print("device-only frames: %.1f us per frame on the device, %.1f us per frame of host time to issue them" % (d, h))

# the read-back alone: one rendered frame, its rectangle copied again and again (no rendering in between)
import ctypes as C
for clip in (1, 0):
    _lib.check(lib.spv_set_tuning(rend._ctx, 9, clip), rend._ctx)
    _lib.check(lib.spv_set_tuning(rend._ctx, 14, 0), rend._ctx)
    _lib.check(lib.spv_select_slot(rend._ctx, 0), rend._ctx)
    rend.set_modelView(cams[3][0])
    p = _lib.IsoParams(rend._box(), 15000., 1., 200, .1, 21, 30, 0)
    _lib.check(lib.spv_render_iso(rend._ctx, C.byref(p)), rend._ctx)
    rend.sync()
    host = _lib._FP()
    for rep in range(2):
        b0 = rend.d2h_bytes()
        t0 = time.perf_counter()
        for i in range(50):
            _lib.check(lib.spv_read_pinned_async(rend._ctx, 2), rend._ctx)
            _lib.check(lib.spv_wait_slot(rend._ctx, 0, C.byref(host)), rend._ctx)
        dt = (time.perf_counter() - t0) / 50
        nb = (rend.d2h_bytes() - b0) / 50
    print("read-back alone, clipped=%d: %.1f us per frame for %.2f MB = %.1f GB/s" % (clip, dt * 1e6, nb / 1e6, nb / dt / 1e9))
_lib.check(lib.spv_set_tuning(rend._ctx, 9, 1), rend._ctx)
