import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from spimagine_b200 import imageprocessor as ip
vf = ip.VolumeFilter(0)
taps = ip.BlurProcessor(4.)._taps()
one = (np.array([1.]),) 
for shape in ((512,512,512),(512,512,520),(512,512,544),(500,520,520),(256,1024,1024)):
    for dt in (np.uint16, np.float32, np.uint8):
        vol = (np.random.default_rng(0).random(shape)*200).astype(dt)
        t = torch.from_numpy(vol.view(np.int16) if dt==np.uint16 else vol).cuda()
        vf.set_tuning(0,0)
        res=[]
        for tp in ((taps[0],[1.],[1.]), ([1.],taps[0],[1.])):
            ms=[]
            for i in range(5):
                vf.load_device(t.data_ptr(), vol.shape, dt); vf.convolve_sep3(*tp); vf.sync(); ms.append(vf.last_ms())
            res.append(min(ms))
        nv=np.prod(shape)
        print(shape, np.dtype(dt).name, "x19+y1+z1: %.3f ms  x1+y19+z1: %.3f ms  (per Gvoxel: %.2f / %.2f ms)"%(res[0],res[1],res[0]/nv*1e9,res[1]/nv*1e9), flush=True)
