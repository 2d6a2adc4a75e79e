"""Generates tests/golden/dispatch_ref.json: the REFERENCE's VolumeRenderer constructed and driven through render()
without OpenCL.  gputools is a recorder: OCLProgram notes its build options and every run_kernel call (kernel name,
global size, scalar arguments with their types), OCLArray / OCLImage are host stand-ins.  What is pinned: the
constructor's defaults, the interpolation defines, which kernel a volume's element type selects, the order of the
launches of an iso-surface frame and every scalar the kernels receive (box, window, gamma, alpha_pow, numParts /
currentPart, isoVal = maxVal / 2, the blur radii 7 and 5, the occlusion parameters), for a list of render() calls.

    python tests/golden/make_dispatch_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
LOG = []


class _Arr(object):
    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.data = self

    @classmethod
    def empty(cls, shape, dtype=np.float32):
        return cls(shape, dtype)

    def write_array(self, a):
        self.written = np.array(a)

    def get(self):
        return np.zeros(self.shape, self.dtype)


class _Img(_Arr):
    def __init__(self, shape, dtype=np.float32):
        _Arr.__init__(self, shape, dtype)
        self.shape = self.shape[::-1]      # (Nx, Ny, Nz), as the reference reads it (volumerender.py:320)


class _Prog(object):
    def __init__(self, fname, build_options=()):
        self.fname = os.path.basename(fname)
        self.build_options = [o for o in build_options]
        LOG.append({"program": self.fname, "build_options": [o if not o.startswith("/") else "<kernels dir>" for o in self.build_options]})

    def run_kernel(self, name, global_size, local_size, *args):
        scal = []
        for a in args:
            if isinstance(a, (np.floating, np.integer)):
                scal.append([type(a).__name__, float(a)])
        LOG.append({"kernel": name, "global": [int(g) for g in global_size], "local": local_size, "scalars": scal})


def import_reference():
    g = types.ModuleType("gputools")
    g.init_device = lambda **k: None
    g.get_device = lambda: type("Dev", (), {"get_info": staticmethod(lambda what: 4e9)})()
    g.OCLProgram, g.OCLArray, g.OCLImage = _Prog, _Arr, _Img
    sys.modules["gputools"] = g
    for name in ("spimagine", "spimagine.utils", "spimagine.config"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    cfg = sys.modules["spimagine.config"]
    cfg.__DEFAULTMAXSTEPS__ = 200          # config/config.py default
    cfg.__QUALIFIER_CONSTANT_TO_GLOBAL__ = False
    sys.modules["spimagine"].config = cfg
    spec = importlib.util.spec_from_file_location("ref_volumerender", os.path.join(REF, "spimagine/volumerender/volumerender.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# (volume dtype, render kwargs, setter calls before it)
CALLS = [
    ("uint16", {}, []),
    ("uint16", {"maxVal": 60000., "minVal": 100., "gamma": .7}, []),
    ("uint16", {"numParts": 4, "currentPart": 3}, [("set_alpha_pow", .6)]),
    ("uint8", {"maxVal": 200.}, [("set_box_boundaries", [-.5, .8, -1, 1, -.2, .3])]),
    ("float32", {"maxVal": 1.5}, []),
    ("float32", {"method": "iso_surface", "maxVal": 3.}, []),
    ("uint16", {"method": "iso_surface", "maxVal": 30000., "gamma": 2.}, [("set_occ_strength", .4), ("set_occ_radius", 11), ("set_occ_n_points", 50)]),
    ("uint8", {"method": "iso_surface"}, [("set_max_val", 77.)]),
    ("uint16", {"method": "nonsense"}, []),
]


def main():
    ref = import_reference()
    rows = []
    for interp in ("linear", "nearest"):
        del LOG[:]
        r = ref.VolumeRenderer((48, 32), interpolation=interp)
        rows.append({"what": "constructor", "interpolation": interp, "log": list(LOG),
                     "defaults": {"gamma": float(r.gamma), "maxVal": float(r.maxVal), "minVal": float(r.minVal),
                                  "alphaPow": float(r.alphaPow), "occ_strength": float(r.occ_strength),
                                  "occ_radius": int(r.occ_radius), "occ_n_points": int(r.occ_n_points),
                                  "boxBounds": [float(b) for b in r.boxBounds], "stackUnits": [float(u) for u in r.stackUnits],
                                  "dtype": np.dtype(r.dtype).name, "width": r.width, "height": r.height,
                                  "projection": np.asarray(r.projection, np.float64).tolist(),
                                  "modelView": np.asarray(r.modelView, np.float64).tolist()}})
    try:
        ref.VolumeRenderer((8, 8), interpolation="cubic")
        rows.append({"what": "bad interpolation", "raises": None})
    except KeyError:
        rows.append({"what": "bad interpolation", "raises": "KeyError"})
    r = ref.VolumeRenderer((48, 32))
    del LOG[:]
    ret = r.render()
    rows.append({"what": "render without data", "returns_none": ret is None, "log": list(LOG)})
    for dtype, kw, setters in CALLS:
        r = ref.VolumeRenderer((48, 32))
        r.set_data(np.zeros((5, 6, 7), dtype))
        for name, v in setters:
            getattr(r, name)(v)
        del LOG[:]
        ret = r.render(**kw)
        rows.append({"what": "render", "dtype": dtype, "kwargs": kw, "setters": setters, "returns_none": ret is None,
                     "log": list(LOG), "output_shape": list(np.shape(r.output))})
    with open(os.path.join(HERE, "dispatch_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_dispatch_golden.py", "rows": rows}, f)
    for row in rows:
        print(row["what"], row.get("dtype", ""), row.get("kwargs", ""), [e.get("kernel", e.get("program")) for e in row.get("log", [])])


if __name__ == "__main__":
    main()
