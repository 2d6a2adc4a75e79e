#!/usr/bin/env python
"""BlurProcessor(sigma) on an N^3 volume: device time of the three-pass and the fused (x + y, z) variants per element
type and tap count, checked against each other.  Under ncu (EXP_BLUR_ONCE=1) it launches every kernel once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import imageprocessor as ip

N = int(os.environ.get("EXP_VOL", 512))
once = os.environ.get("EXP_BLUR_ONCE") == "1"
vf = ip.VolumeFilter(0)
base = scenes.vol_g(N, np.uint16, seed=0)
for dt in ((np.uint16, np.float32) if once else (np.uint16, np.float32, np.uint8)):
    vol = base if dt == np.uint16 else (base.astype(np.float32) if dt == np.float32 else (base >> 8).astype(np.uint8))
    t = torch.from_numpy(vol.view(np.int16) if dt == np.uint16 else vol).cuda()
    for sigma in ((4.,) if once else (1., 2., 4., 7.)):
        taps = ip.BlurProcessor(sigma)._taps()
        res = {}
        for wide in ((1,) if once else (16, 32, 1602, 1604)):
            vf.set_tuning(1, wide); vf.set_tuning(0, 0)
            ms = []
            for i in range(1 if once else 6):
                vf.load_device(t.data_ptr(), vol.shape, dt)
                vf.convolve_sep3(*taps)
                vf.sync()
                ms.append(vf.last_ms())
            print("   axis kernel variant %d: three passes %.3f ms" % (wide, min(ms)), flush=True)
        vf.set_tuning(1, int(os.environ.get("EXP_WIDE", 1)))
        for fuse in (0, 1):
            vf.set_tuning(0, fuse)
            ms = []
            for i in range(1 if once else 8):
                vf.load_device(t.data_ptr(), vol.shape, dt)
                vf.convolve_sep3(*taps)
                vf.sync()
                ms.append(vf.last_ms())
            res[fuse] = (min(ms), vf.result() if not once else None)
        same = once or bool(np.array_equal(res[0][1], res[1][1]))
        nv = float(N) ** 3
        es = np.dtype(dt).itemsize
        print("%-8s sigma %g (%2d taps): three passes %.3f ms, fused x+y then z %.3f ms (%.1f Gvoxel/s, %.0f GB/s algorithmic, "
              "%.0f GB/s moved) identical=%s" % (np.dtype(dt).name, sigma, len(taps[0]), res[0][0], res[1][0],
                                                 nv / res[1][0] / 1e6, nv * (es + 4) / res[1][0] / 1e6,
                                                 nv * (es + 12) / res[1][0] / 1e6, same), flush=True)
    del t
