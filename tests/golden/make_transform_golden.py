"""Generates tests/golden/transform_ref.json by driving the REFERENCE's TransformModel
(/root/reference/spimagine/models/transform_model.py) through a scripted sequence of setter calls.  PyQt5 is absent:
QObject is a plain object and pyqtSignal a stand-in that records (signal name, arguments) of every emit, `spimagine`
is entered as a bare namespace.  Recorded after every call: the emitted signals in order, the state a keyframe would
store (toTransformData), cameraZ / scaleAll, getUnscaledModelView(), getModelView() (with a data model of
128 x 64 x 30 voxels attached from step `attach` on) and getProjection().  tests/test_transform_model.py replays
the same script on spimagine_b200.transform_model.TransformModel.

    python tests/golden/make_transform_golden.py
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
LOG = []

# (method, args) -- angles are half angles, the zoom leaves its clamp on both sides, values repeat to hit the
# "only when changed" setters, the projection flips, a keyframe is loaded and taken back
SCRIPT = [
    ("reset", [10., 3000., [.2, .2, 1.]]),
    ("addRotation", [.3, 0., 1., 0.]),
    ("addRotation", [-.2, 1., 0., 0.]),
    ("addRotation", [.15, 0., 0., 1., False]),
    ("setZoom", [1.7]),
    ("setZoom", [5.]),
    ("setZoom", [.1]),
    ("setTranslate", [.1, -.2, .3]),
    ("setTranslate", [.1, -.2, .3]),
    ("addTranslate", [.05, .05, -.3]),
    ("setIso", [True]),
    ("setIso", [True]),
    ("setInterpolate", [False]),
    ("setInterpolate", [False]),
    ("setOccStrength", [.15]),
    ("setOccStrength", [.4]),
    ("setOccRadius", [11]),
    ("setOccNPoints", [31]),
    ("setOccNPoints", [50]),
    ("setPerspective", [False]),
    ("setZoom", [1.3]),
    ("setBounds", [-.5, .5, -1., 1., 0., .75]),
    ("setGamma", [.6]),
    ("setAlphaPow", [.25]),
    ("setValueScale", [0., 40000.]),
    ("setMin", [-5.]),
    ("setMax", [123.]),
    ("setStackUnits", [.16, .16, .5]),
    ("setBox", [False]),
    ("setShowSlice", [True]),
    ("setSliceDim", [2]),
    ("setSlicePos", [17]),
    ("attach", []),
    ("setPos", [3]),
    ("setRotation", [.7, 0., .6, .8]),
    ("setPerspective", [True]),
    ("setEyeDistProj", [.1]),
    ("setEyeDistCam", [.2]),
    ("keyframe", []),
    ("center", []),
    ("reset", []),
]


def import_reference():
    qt = types.ModuleType("PyQt5")
    core = types.ModuleType("PyQt5.QtCore")

    class _Signal(object):
        def __init__(self, *a):
            self.name = "?"

        def __set_name__(self, owner, name):
            self.name = name

        def emit(self, *a):
            LOG.append([self.name, [x.item() if isinstance(x, np.generic) else x for x in a]])

        def connect(self, *a):
            pass

    class QObject(object):
        def __init__(self, *a, **k):
            pass

    core.QObject, core.pyqtSignal = QObject, _Signal
    qt.QtCore = core
    sys.modules["PyQt5"], sys.modules["PyQt5.QtCore"] = qt, core
    for name in ("spimagine", "spimagine.models", "spimagine.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    sys.modules["spimagine"].models = sys.modules["spimagine.models"]
    import spimagine.models.transform_model as tm
    import spimagine.models.keyframe_model as km
    return tm, km


class FakeDataModel(object):
    def __init__(self):
        self.positions = []

    def size(self):
        return (5, 30, 64, 128)

    def setPos(self, pos):
        self.positions.append(pos)


def keyframe_of(mod):
    """the TransformData of step "keyframe", built with the module's own classes"""
    return mod.TransformData(quatRot=mod.Quaternion(.5, -.5, .5, .5), zoom=1.2, dataPos=4, minVal=0., maxVal=777.,
                             gamma=1.5, translate=[.3, 0, -.1], bounds=[-1, .2, -.3, 1, -1, 1], isBox=True,
                             isIso=False, alphaPow=.1, isSlice=False, slicePos=5, sliceDim=1)


def snapshot(m, log):
    td = m.toTransformData()
    return {"signals": log,
            "state": {"quatRot": [float(v) for v in td.quatRot.data], "zoom": float(td.zoom), "dataPos": int(td.dataPos),
                      "minVal": float(td.minVal), "maxVal": float(td.maxVal), "gamma": float(td.gamma),
                      "translate": [float(v) for v in td.translate], "bounds": [float(v) for v in td.bounds],
                      "isBox": bool(td.isBox), "isIso": bool(td.isIso), "alphaPow": float(td.alphaPow),
                      "isSlice": bool(td.isSlice), "slicePos": int(td.slicePos), "sliceDim": int(td.sliceDim),
                      "is_interpolate": bool(m.is_interpolate), "occ": [float(m.occ_strength), int(m.occ_radius),
                                                                       int(m.occ_n_points)],
                      "stackUnits": [float(u) for u in m.stackUnits], "isPerspective": bool(m.isPerspective),
                      "cameraZ": float(m.cameraZ), "scaleAll": float(m.scaleAll),
                      "eye": [float(m.eye_dist_proj), float(m.eye_dist_cam)]},
            "unscaled": np.asarray(m.getUnscaledModelView(), np.float64).tolist(),
            "modelView": np.asarray(m.getModelView(), np.float64).tolist(),
            "projection": np.asarray(m.getProjection(), np.float64).tolist()}


def run(model, mod, log, fake):
    """replays SCRIPT on `model`; `log` is the list the model's signals append to"""
    steps = []
    del log[:]
    for name, args in SCRIPT:
        if name == "attach":
            model.setModel(fake)
        elif name == "keyframe":
            model.fromTransformData(keyframe_of(mod))
        else:
            getattr(model, name)(*args)
        steps.append(snapshot(model, list(log)))
        del log[:]
    return steps


def main():
    tm, km = import_reference()
    import contextlib
    import io
    mod = types.SimpleNamespace(TransformData=tm.TransformData, Quaternion=km.Quaternion)
    with contextlib.redirect_stdout(io.StringIO()):     # setEyeDist* print
        m = tm.TransformModel()
        init = snapshot(m, list(LOG))
        fake = FakeDataModel()
        steps = run(m, mod, LOG, fake)
    with open(os.path.join(HERE, "transform_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_transform_golden.py (reference TransformModel)",
                   "script": SCRIPT, "init": init, "steps": steps, "data_positions": fake.positions}, f)
    print(len(steps), "steps;", sum(len(s["signals"]) for s in steps), "signals; positions", fake.positions)


if __name__ == "__main__":
    main()
