"""Generates tests/golden/renderer_host_ref.json: the HOST side of the REFERENCE's VolumeRenderer
(/root/reference/spimagine/volumerender/volumerender.py) executed without OpenCL.  gputools is stubbed, the class is
instantiated with __new__ (its __init__ builds the OpenCL program), the two matrix buffers are recorders: what
set_units / set_modelView / set_projection -> update_matrices would upload (invM, invP as float32, row-major) is
captured for a set of volume shapes, units, cameras and projections; plus _stack_scale_mat and
_get_downsampled_data_slices.

    python tests/golden/make_renderer_host_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    g = types.ModuleType("gputools")
    for n in ("init_device", "get_device", "OCLProgram", "OCLArray", "OCLImage"):
        setattr(g, n, type(n, (object,), {}))
    sys.modules["gputools"] = g
    for name in ("spimagine", "spimagine.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_volumerender", os.path.join(REF, "spimagine/volumerender/volumerender.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Buf(object):
    def write_array(self, a):
        self.last = np.array(a)


def cases():
    import spimagine.utils.transform_matrices as tm
    rng = np.random.default_rng(3)
    out = []
    for shape, units in (((512, 512, 512), (1., 1., 1.)), ((1024, 512, 100), (.162, .162, .81)), ((33, 65, 17), (1, 2, 5)),
                         ((7, 1, 300), (.5, .5, .5))):
        for k in range(4):
            M = np.dot(tm.mat4_translate(*(rng.normal(size=3) * [.3, .3, 1] + [0, 0, -4])),
                       np.dot(tm.mat4_rotation(rng.uniform(0, 6.3), *rng.normal(size=3)), tm.mat4_scale(*(rng.random(3) + .5))))
            P = tm.mat4_perspective(60, 1., .1, 10) if k % 2 == 0 else tm.mat4_ortho(-2., 2., -2., 2., -1.5, 1.5)
            out.append((shape, units, M, P))
    out.append(((64, 64, 64), (1., 1., 1.), tm.mat4_identity(), tm.mat4_perspective()))       # the constructor's defaults
    out.append(((64, 64, 64), (1., 1., 1.), tm.mat4_translate(0, 0, -5.), tm.mat4_perspective(60, 1., 1, 10)))
    return out


def main():
    ref = import_reference()
    VR = ref.VolumeRenderer
    rows = []
    for shape, units, M, P in cases():
        r = VR.__new__(VR)
        r.invMBuf, r.invPBuf = _Buf(), _Buf()
        r.dataImg = type("Img", (), {"shape": shape})()
        r.set_units(units)
        r.modelView = np.identity(4)            # __init__ sets both before the first update_matrices
        r.set_projection(P)
        r.set_modelView(M)
        rows.append({"shape_xyz": list(shape), "units": [float(u) for u in units], "modelView": np.asarray(M, np.float64).tolist(),
                     "modelView_dtype": str(np.asarray(M).dtype),
                     "projection": np.asarray(P, np.float64).tolist(), "projection_dtype": str(np.asarray(P).dtype),
                     "mScale": np.asarray(r._stack_scale_mat(), np.float64).tolist(),
                     "invM_f32": r.invMBuf.last.astype(np.float64).tolist(), "invM_dtype": str(r.invMBuf.last.dtype),
                     "invP_f32": r.invPBuf.last.astype(np.float64).tolist()})
    slices = []
    r = VR.__new__(VR)
    for shape, dtype, mem in (((64, 64, 64), "uint16", 1e9), ((64, 64, 64), "uint16", 2e5), ((100, 50, 30), "float32", 1e4),
                              ((10, 200, 17), "uint8", 3e3), ((16, 16, 16), "float32", 16384), ((16, 16, 16), "float32", 16383)):
        r.memMax = mem
        s = r._get_downsampled_data_slices(np.zeros(shape, dtype))
        slices.append({"shape": list(shape), "dtype": dtype, "memMax": mem,
                       "slices": None if s is None else [[x.start, x.stop, x.step] for x in s]})
    with open(os.path.join(HERE, "renderer_host_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_renderer_host_golden.py", "matrices": rows, "downsample": slices}, f)
    print("wrote %d matrix cases, %d downsample cases; invM dtype %s" % (len(rows), len(slices), rows[0]["invM_dtype"]))


if __name__ == "__main__":
    main()
