"""Generates tests/golden/setdata_ref.json: the REFERENCE's VolumeRenderer.set_data / set_dtype / set_shape /
update_data (/root/reference/spimagine/volumerender/volumerender.py:166-294) executed without OpenCL -- gputools'
OCLImage / OCLArray are recorders -- on seeded arrays of many element types: which element type the renderer settles
on, the image shape and type it allocates, the strided downsampling it applies under a memory budget, the texels it
uploads (SHA-1), whether it keeps a reference to the caller's array, and the exceptions it raises.
OCLImage.empty(shape, dtype).shape is taken to be shape[::-1] = (Nx, Ny, Nz), which is how the reference reads it
(volumerender.py:320); gputools itself is not available.

    python tests/golden/make_setdata_golden.py
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class _Image(object):
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape)[::-1], np.dtype(dtype)
        self.written = None

    @classmethod
    def empty(cls, shape, dtype=np.float32):
        return cls(shape, dtype)

    def write_array(self, a):
        self.written = np.array(a)


def import_reference():
    g = types.ModuleType("gputools")
    for n in ("init_device", "get_device", "OCLProgram"):
        setattr(g, n, type(n, (object,), {}))
    g.OCLImage = _Image
    g.OCLArray = _Image
    sys.modules["gputools"] = g
    for name in ("spimagine", "spimagine.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_volumerender", os.path.join(REF, "spimagine/volumerender/volumerender.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make(seed, shape, dtype):
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return (rng.normal(size=shape) * 300).astype(dt)
    if dt.kind == "b":
        return rng.random(shape) > .5
    info = np.iinfo(dt)
    return rng.integers(max(info.min, -70000), min(int(info.max), 70000) + 1, size=shape).astype(dt)


CASES = [  # seed, shape (z, y, x), dtype, memMax, autoConvert, copyData
    (0, (5, 6, 7), "float32", 1e9, True, False), (1, (5, 6, 7), "uint16", 1e9, True, False),
    (2, (5, 6, 7), "uint8", 1e9, True, True), (3, (4, 3, 9), "float64", 1e9, True, False),
    (4, (4, 3, 9), "int16", 1e9, True, False), (5, (4, 3, 9), "int32", 1e9, True, False),
    (6, (4, 3, 9), "uint32", 1e9, True, False), (7, (4, 3, 9), "int64", 1e9, True, False),
    (8, (4, 3, 9), "float16", 1e9, True, False), (9, (4, 3, 9), "bool", 1e9, True, False),
    (10, (4, 3, 9), "int8", 1e9, True, False),
    (11, (16, 12, 20), "uint16", 2000., True, False), (12, (16, 12, 20), "float32", 700., True, False),
    (13, (16, 12, 20), "float64", 700., True, False),
    (14, (3, 3, 3), "float64", 1e9, False, False), (15, (3, 3, 3), "int16", 1e9, False, False),
]


def main():
    ref = import_reference()
    VR = ref.VolumeRenderer
    # the reference indexes with a LIST of slices (data[self.dataSlices], volumerender.py:251, 283), which numpy
    # stopped accepting in 1.23: handed over as a tuple here, which is what the list meant
    orig = VR._get_downsampled_data_slices
    VR._get_downsampled_data_slices = lambda self, d: (lambda sl: None if sl is None else tuple(sl))(orig(self, d))
    rows = []
    for first in ("float32", "uint16"):        # the element type the renderer holds before the call
        for seed, shape, dtype, mem, auto, copy in CASES:
            r = VR.__new__(VR)
            r.isGPU = True
            r.width = r.height = 8
            r.memMax = mem
            r.invMBuf, r.invPBuf = _Image((16,), np.float32), _Image((16,), np.float32)
            r.modelView, r.projection = np.identity(4), np.identity(4)
            r.stackUnits = np.ones(3)
            r.set_dtype(np.dtype(first).type)
            data = make(seed, shape, dtype)
            row = {"first": first, "seed": seed, "shape": list(shape), "dtype": dtype, "memMax": mem, "autoConvert": auto,
                   "copyData": copy}
            try:
                r.set_data(data, autoConvert=auto, copyData=copy)
            except NotImplementedError:
                row["raises"] = "NotImplementedError"
                rows.append(row)
                continue
            w = r.dataImg.written
            row.update({"renderer_dtype": np.dtype(r.dtype).name, "image_shape_xyz": list(r.dataImg.shape),
                        "image_dtype": r.dataImg.dtype.name,
                        "slices": None if r.dataSlices is None else [[s.start, s.stop, s.step] for s in r.dataSlices],
                        "uploaded_shape": list(w.shape), "uploaded_dtype": w.dtype.name,
                        "uploaded_sha1": hashlib.sha1(np.ascontiguousarray(w).tobytes()).hexdigest(),
                        "keeps_callers_array": bool(r._data is data)})
            rows.append(row)
    with open(os.path.join(HERE, "setdata_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_setdata_golden.py", "cases": rows}, f)
    print("wrote %d cases" % len(rows))
    for row in rows[:18]:
        print({k: row[k] for k in ("first", "dtype", "renderer_dtype", "uploaded_dtype", "slices", "keeps_callers_array", "raises") if k in row})


if __name__ == "__main__":
    main()
