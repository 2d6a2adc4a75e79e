"""Are the iso / max-projection frames of the bench scenes the same with and without row clipping of the read-back
(spv_set_tuning knob 9)?  Prints one digest per setting; they must agree."""
import hashlib
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import scenes  # noqa: E402
from spimagine_b200 import VolumeRenderer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
vol = bench.vol_g_device((N, N, N), 1, 0, dev)
for clip in (1, 0, 1):
    rend = VolumeRenderer((1024, 1024), device=0, max_steps=bench.MAX_STEPS, pinned_outputs=bool(clip))
    rend._check(rend._lib.spv_set_tuning(rend._ctx, 9, clip))
    rend.set_data_device(vol.data_ptr(), (N, N, N), np.uint16)
    rend.sync()
    lo, hi = rend.data_min_max
    rend.set_max_val(hi)
    cams = [scenes.gui_camera(2 * math.pi * f / 36, 4.0) for f in range(36)]
    rend.set_projection(cams[0][1])
    d_iso, d_mip, d_all = hashlib.sha1(), hashlib.sha1(), hashlib.sha1()
    hit = 0
    for i in range(8):
        rend.set_modelView(cams[i][0])
        rend.render(method="iso_surface")
        d_iso.update(np.ascontiguousarray(rend.output).tobytes() + np.ascontiguousarray(rend.output_alpha).tobytes())
        for a in (rend.output, rend.output_depth, rend.output_normals, rend.output_occlusion):
            d_all.update(np.ascontiguousarray(a).tobytes())
        hit = int(np.isfinite(rend.output_depth).sum())
        rend.render()
        d_mip.update(np.ascontiguousarray(rend.output).tobytes() + np.ascontiguousarray(rend.output_alpha).tobytes())
    print("clip %d: iso out+alpha %s  iso all planes %s  mip %s  surface pixels %d  max %g" % (
        clip, d_iso.hexdigest()[:12], d_all.hexdigest()[:12], d_mip.hexdigest()[:12], hit, hi))
    rend.close()
