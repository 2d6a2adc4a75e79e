#!/usr/bin/env python
"""Raw slab render time of ONE rank against the slab count (what bounds sort-last scaling), on one GPU."""
import ctypes as C
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes, bench
from spimagine_b200 import _lib
from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs, slab_with_halo

N = int(os.environ.get("EXP_VOL", 1024)); IMG = int(os.environ.get("EXP_IMG", 1024))
dev = torch.device("cuda", 0)
for world in [int(w) for w in os.environ.get('EXP_WORLDS', '1,2,4,8').split(',')]:
    for rank in sorted(set([0, world // 2])):
        z0, z1 = partition_slabs(N, world)[rank]
        lo, hi = slab_with_halo(z0, z1, N)
        slab = bench.vol_g_slab_device(N, lo, hi, 2, dev)
        rend = SlabMaxProjector((IMG, IMG), rank=rank, world=world)
        rend.set_layout(os.environ.get("EXP_LAYOUT", "zpair"))
        rend.set_slab((np.uint16, N, N), N, z0, z1, device_ptr=slab.data_ptr())
        rend.sync(); del slab; torch.cuda.empty_cache()
        rend.enable_stats(True)
        P = scenes.gui_camera(0, 4.0)[1]
        rend.set_projection(P)
        p = _lib.MipParams(rend._box(), 0., 60000., 1., 0., 1, 0, 200, _lib.MIP_RAW_ONLY)
        line = []
        for deg in (0, 30, 60, 90):
            rend.set_modelView(scenes.gui_camera(math.radians(deg), 4.0)[0])
            ts = []
            for rep in range(8):
                rend._check(rend._lib.spv_render_mip(rend._ctx, C.byref(p)))
                rend.sync(); ts.append(rend.last_render_ms())
            hits, issued = rend.last_stats()
            line.append("%2d deg %6.1f us (%5.1fM samples)" % (deg, 1e3 * float(np.median(ts[2:])), issued / 1e6))
        print("world %d rank %d: %s" % (world, rank, " | ".join(line)), flush=True)
        rend.close()
