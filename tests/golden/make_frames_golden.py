"""Generates tests/golden/frames_ref.json by reading the synthetic inputs of tests/golden/frames_inputs.py with the
REFERENCE's own containers (/root/reference/spimagine/models/data_model.py, utils/imgutils.py: SpimData, RawData,
RawMultipleFiles, XwingData, NumpyData, fromSpimFolder), Qt / tifffile / czifile stubbed, `spimagine` entered as a bare
namespace.  Recorded per container: size(), sizeT(), stackUnits and dtype / shape / sha1 of every time point.
np.float and np.asscalar (removed from numpy, used by parseMetaFile) are restored for the run.

    python tests/golden/make_frames_golden.py
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
REF = "/root/reference"


def import_reference():
    class _Anything(object):
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    for name in ("PyQt5", "PyQt5.QtCore", "PyQt5.QtWidgets", "PyQt5.QtGui", "tifffile", "spimagine.lib",
                 "spimagine.lib.czifile", "spimagine.gui", "spimagine.gui.shape_dtype_dialog"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    core = sys.modules["PyQt5.QtCore"]
    core.QObject = type("QObject", (object,), {"__init__": lambda self, *a, **k: None})
    core.QThread = type("QThread", (object,), {"__init__": lambda self, *a, **k: None, "LowPriority": 0,
                                               "start": lambda self, priority=None: None})   # the prefetch thread never runs
    core.pyqtSignal = core.QReadWriteLock = _Anything
    sys.modules["PyQt5"].QtCore = core
    t = sys.modules["tifffile"]
    t.TiffFile = t.imsave = t.imread = _Anything
    sys.modules["spimagine.lib.czifile"].CziFile = _Anything
    sys.modules["spimagine.gui.shape_dtype_dialog"].ShapeDtypeDialog = _Anything
    for name in ("spimagine", "spimagine.models", "spimagine.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    if not hasattr(np, "float"):
        np.float = float
    import spimagine.utils.imgutils as imgutils
    sys.modules["spimagine.utils"].imgutils = imgutils
    import spimagine.models.data_model as dm
    return dm, imgutils


def describe(c):
    size = [int(s) for s in c.size()]
    pts = []
    for t in range(int(c.sizeT())):
        a = np.ascontiguousarray(c[t])
        pts.append({"shape": list(a.shape), "dtype": a.dtype.newbyteorder("=").name if a.dtype.byteorder in "<>" else a.dtype.name,
                    "sha1": hashlib.sha1(a.tobytes()).hexdigest()})
    return {"size": size, "sizeT": int(c.sizeT()), "stackUnits": [float(u) for u in c.stackUnits], "points": pts}


def main():
    import frames_inputs
    dm, imgutils = import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as root:
        for key, (cls, args, kw) in frames_inputs.build(root).items():
            out[key] = describe(getattr(dm, cls)(*args, **kw))
        # DataModel: the container loadFromPath picks, the neighbourhood it prefetches, what model[pos] returns
        model = {}
        for key, path in (("spim", os.path.join(root, "spim")), ("xwing", os.path.join(root, "xwing"))):
            m = dm.DataModel.fromPath(path, prefetchSize=2)
            a = np.ascontiguousarray(m[1])
            model[key] = {"container": type(m.dataContainer).__name__, "prefetchSize": int(m.prefetchSize),
                          "sizeT": int(m.sizeT()), "pos_after_init": int(m.pos),
                          "neighborhood": {str(p): [int(k) for k in m.neighborhood(p)] for p in range(int(m.sizeT()))},
                          "item1_sha1": hashlib.sha1(a.tobytes()).hexdigest()}
        out["DataModel"] = model
        sub = imgutils.fromSpimFolder(os.path.join(root, "spim"), pos=2, count=2)
        out["fromSpimFolder_pos2_count2"] = {"shape": list(sub.shape), "sha1": hashlib.sha1(np.ascontiguousarray(sub).tobytes()).hexdigest()}
        for key, pos, count in (("fromSpimFolder_pos9_count5", 9, 5), ("fromSpimFolder_pos1_count0", 1, 0),
                                ("fromSpimFolder_neg_count1", -3, 1)):
            sub = imgutils.fromSpimFolder(os.path.join(root, "spim"), pos=pos, count=count)
            out[key] = {"shape": list(sub.shape), "pos": pos, "count": count,
                        "sha1": hashlib.sha1(np.ascontiguousarray(sub).tobytes()).hexdigest()}
    # the synthetic demo volume and the placeholder container (data_model.py:434-472, 518-531)
    demo = dm.DemoData(24)
    pts = {}
    for t in (0, 1, 23):
        a = np.ascontiguousarray(demo[t])
        pts[str(t)] = {"sha1": hashlib.sha1(a.tobytes()).hexdigest(), "dtype": a.dtype.name, "max": float(a.max()),
                       "sum": float(a.astype(np.float64).sum())}
    out["DemoData_24"] = {"size": [int(s) for s in demo.size()], "sizeT": int(demo.sizeT()),
                          "stackUnits": [float(u) for u in demo.stackUnits], "points": pts}
    # OverlayData (models/overlay_volumes.py): a walk back and forth along each axis
    import spimagine.models.overlay_volumes as ov
    rng = np.random.default_rng(21)
    ox, oy = rng.integers(0, 60000, (5, 6, 7)).astype(np.uint16), rng.integers(0, 60000, (5, 6, 7)).astype(np.uint16)
    walks = {}
    for axis in (-1, 0, 1):
        o = ov.OverlayData(ox, oy, axis=axis)
        walk = [0, 3, 3, 1, o.size()[0] - 1, 2, 0]
        walks[str(axis)] = {"size": [int(s) for s in o.size()], "sizeT": int(o.sizeT()), "walk": walk,
                            "sha1": [hashlib.sha1(np.ascontiguousarray(o[i]).tobytes()).hexdigest() for i in walk]}
    out["OverlayData_seed21_5x6x7"] = walks
    e = dm.EmptyData()
    out["EmptyData"] = {"size": [int(s) for s in e.size()], "sizeT": int(e.sizeT()),
                        "stackUnits": [float(u) for u in e.stackUnits], "dtype": e[0].dtype.name,
                        "shape": list(e[0].shape), "sum": int(e[0].sum())}
    with open(os.path.join(HERE, "frames_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_frames_golden.py (reference data_model / imgutils)", "containers": out}, f, indent=1)
    for k, v in out.items():
        print(k, v.get("size"), v.get("stackUnits"))


if __name__ == "__main__":
    main()
