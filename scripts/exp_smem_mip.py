#!/usr/bin/env python
"""Experiment: the TMA-staged shared-memory max projection (spv_set_mip_path SMEM) against the texture-unit kernel on
BASELINE configs[1]: parity vs the oracle on sampled rows, share of samples taken in software, time per frame by angle."""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from spimagine_b200 import VolumeRenderer, _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
check = "--no-oracle" not in sys.argv
vol = scenes.vol_g(N, np.uint16, seed=0)
r = VolumeRenderer((W, W), max_steps=200)
r.set_data(vol)
r.set_max_val(60000.)
if check:
    from oracle import oracle
    o = oracle.OracleRenderer((W, W), kind="port")
    o.set_data(vol)
HYB = [int(x) for x in os.environ.get("HYB", "0,2,3,4,5").split(",")]
print("angle   tmu_us  smem_us  sw_share  max|smem-tmu|  max|smem-oracle|  max|tmu-oracle| (rows/32) | hybrid tex_of8 -> us")
for deg in (0, 15, 30, 45, 60, 90, 135, 200, 270, 315):
    M, P = scenes.gui_camera(math.radians(deg), 4.0)
    r.set_projection(P)
    r.set_modelView(M)
    res = {}
    for path in ("tmu", "smem"):
        r.set_mip_path(path)
        r.enable_stats(True)
        r.render()
        v = (C.c_ulonglong * 4)()
        r._check(r._lib.spv_last_stats(r._ctx, v, 4))
        r.enable_stats(False)
        img, alpha = r.output.copy(), r.output_alpha.copy()
        assert r.mip_path_used() == path, (r.mip_path_used(), path)
        ts = []
        for i in range(12):
            r.render_device_only()
            r.sync()
            ts.append(r.last_render_ms() * 1e3)
        res[path] = (img, alpha, float(np.median(ts[2:])), int(v[0]), int(v[1]), int(v[2]))
    it, at, tt, _, tot_t, _ = res["tmu"]
    is_, as_, ts_, hits, tot, nsw = res["smem"]
    assert np.array_equal(at, as_), "alpha planes differ"
    line = "%5d  %7.1f  %7.1f  %7.3f  %12.3g" % (deg, tt, ts_, nsw / max(1, tot), np.abs(is_ - it).max())
    if check:
        o.set_modelView(M)
        o.set_projection(P)
        o.lib.so_set_row_sampling(0, 32)
        o.render(maxVal=60000.)
        o.lib.so_set_row_sampling(0, 1)
        rows = slice(0, W, 32)
        line += "  %12.3g  %12.3g" % (np.abs(is_[rows] - o.output[rows]).max(), np.abs(it[rows] - o.output[rows]).max())
    hyb = []
    for n in HYB:
        r._check(r._lib.spv_set_tuning(r._ctx, 10, n))
        ts = []
        for i in range(12):
            r.render_device_only()
            r.sync()
            ts.append(r.last_render_ms() * 1e3)
        hyb.append("%d:%.1f" % (n, float(np.median(ts[2:]))))
        if n:
            r.render()
            assert np.array_equal(r.output_alpha, at)
            assert np.abs(r.output - it).max() < 2e-4
    r._check(r._lib.spv_set_tuning(r._ctx, 10, 0))
    print(line + "   (hit rays %d, samples ok: %s) | %s" % (hits, tot == hits * 208, " ".join(hyb)), flush=True)
r.close()
