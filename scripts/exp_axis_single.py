#!/usr/bin/env python
"""One frame per launch: the full-grid launch of spv_render_mip against the rectangle-restricted launch of
spv_render_mip_batch with n = 1, and blocking render() against render_batch + wait (configs[1], the 20 views of the bench)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, scenes
from spimagine_b200 import VolumeRenderer
vol = scenes.vol_g(512, np.uint16, seed=0)
rend = VolumeRenderer((1024, 1024), pinned_outputs=True)
rend.set_data(vol); rend.set_max_val(60000.)
cams = [scenes.gui_camera(math.radians(18. * i), 4.0) for i in range(20)]
rend.set_projection(cams[0][1])
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); rend.use_stream(stream.cuda_stream)
def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
def full():
    for M, _ in cams:
        rend.set_modelView(M); rend.render_device_only()
def rect():
    for M, _ in cams:
        rend.render_batch([M], False)
print("device, one frame per launch: full grid %.1f us per frame, rectangle only %.1f us per frame" % (timed(full) / 20, timed(rect) / 20))
def blocking():
    for M, _ in cams:
        rend.set_modelView(M); rend.render()
def blocking_batch():
    for M, _ in cams:
        rend.batch_frames_of(rend.render_batch([M], True))
for name, fn in (("render()", blocking), ("render_batch([M]) + wait", blocking_batch)):
    fn(); t0 = time.perf_counter()
    for _ in range(5): fn()
    print("blocking per frame, %s: %.1f us" % (name, (time.perf_counter() - t0) / 100 * 1e6))
