import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from spimagine_b200 import VolumeRenderer
N = int(sys.argv[1]); W = int(sys.argv[2])
vol = scenes.vol_g(N, np.uint16, seed=0)
r = VolumeRenderer((W, W), max_steps=200)
r.set_data(vol); r.set_max_val(60000.)
M, P = scenes.gui_camera(0.3, 4.0)
r.set_projection(P); r.set_modelView(M)
r.set_mip_path("smem")
for stats in (False, True):
    r.enable_stats(stats)
    r.render()
    print("ok stats", stats, r.output.max(), r.mip_path_used(), flush=True)
