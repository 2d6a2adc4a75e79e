"""Generates tests/golden/keyframes_ref.json by importing the REFERENCE's keyframe / transform models from
/root/reference (only present in the build container; the GPU box and the tests read the committed JSON).

spimagine.models.keyframe_model and .transform_model need PyQt5 (absent): QObject / pyqtSignal are stubbed with
do-nothing stand-ins, and `spimagine` itself is entered as a bare namespace so that its __init__ (which pulls in
pyopencl / gputools / Qt widgets) does not run.  Everything numeric below is computed by the reference's own code:
  * a keyframe list with eased, non-monotonic and iso keyframes, written with the reference's KeyFrameEncoder;
  * KeyFrameList.getTransform at the record loop's key times (gui/keyframe_view.py:644-653) and at edge times;
  * TransformModel.fromTransformData(...).getUnscaledModelView() / getProjection() for both projections.

    python tests/golden/make_keyframe_golden.py
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    qt = types.ModuleType("PyQt5")
    core = types.ModuleType("PyQt5.QtCore")

    class _Signal(object):
        def __init__(self, *a):
            pass

        def emit(self, *a):
            pass

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
    if not hasattr(np, "asscalar"):  # removed from numpy; the reference's encoder calls it
        np.asscalar = lambda a: a.item()
    import spimagine.models.transform_model as tm
    import spimagine.models.keyframe_model as km
    sys.modules["spimagine.models"].transform_model = tm
    return km, tm


def td_dict(td):
    return {"quatRot": [float(v) for v in td.quatRot.data], "zoom": float(td.zoom), "dataPos": int(td.dataPos),
            "minVal": float(td.minVal), "maxVal": float(td.maxVal), "gamma": float(td.gamma),
            "translate": [float(v) for v in td.translate], "bounds": [float(v) for v in td.bounds],
            "isBox": bool(td.isBox), "isIso": bool(td.isIso), "alphaPow": float(td.alphaPow),
            "isSlice": bool(td.isSlice), "slicePos": int(td.slicePos), "sliceDim": int(td.sliceDim)}


def main():
    km, tm = import_reference()
    Q = km.Quaternion
    k = km.KeyFrameList()
    k.addItem(km.KeyFrame(0., km.TransformData(quatRot=Q(1, 0, 0, 0), zoom=1., dataPos=0, maxVal=60000.)))
    k.addItem(km.KeyFrame(.3, km.TransformData(quatRot=Q(.71, .71, 0, 0), zoom=1.4, dataPos=7, minVal=100.,
                                               maxVal=40000., gamma=.7, translate=[.1, -.2, .05],
                                               bounds=[-.8, .9, -1, 1, -.5, .5]), interp_elasticity=3.))
    k.addItem(km.KeyFrame(.55, km.TransformData(quatRot=Q(-.5, .5, -.5, .5), zoom=.8, dataPos=3, maxVal=30000.,
                                                isIso=True, slicePos=9, isSlice=True, sliceDim=2)))
    k.addItem(km.KeyFrame(.8, km.TransformData(quatRot=Q(-.5001, .4999, -.5, .5), zoom=2., dataPos=12,
                                               maxVal=50000., alphaPow=.4, slicePos=2), interp_elasticity=.5))
    k.addItem(km.KeyFrame(.9, km.TransformData(quatRot=Q(0, .6, .8, 0), zoom=2.6, dataPos=15, maxVal=55000.)))
    k.addItem(km.KeyFrame(1., km.TransformData(quatRot=Q(0, 0, 1, 0), zoom=1., dataPos=20, maxVal=60000.)))
    text = k._to_JSON()

    n_frames = 24
    times = [1. * r / n_frames for r in range(1, n_frames + 1)] + [-.1, 0., .3, .55, .8, .9, 1.2, .299999, .15]
    loaded = km.KeyFrameList._from_JSON(text)   # the reference's loader drops interp_elasticity
    model = tm.TransformModel()

    class _NoData(object):   # fromTransformData -> setPos -> dataModel.setPos (the time point switch)
        def setPos(self, pos):
            self.pos = pos

    model.setModel(_NoData())
    samples = []
    for t in times:
        td = k.getTransform(t)
        rec = {"t": t, "transform": td_dict(td), "transform_after_reload": td_dict(loaded.getTransform(t))}
        for persp in (True, False):
            model.setPerspective(persp)
            model.fromTransformData(td)
            rec["modelView_%s" % ("persp" if persp else "ortho")] = model.getUnscaledModelView().tolist()
            rec["projection_%s" % ("persp" if persp else "ortho")] = np.asarray(model.getProjection()).tolist()
        # what GLWidget.render hands to the renderer's setters (glwidget.py:617-621)
        rec["renderer"] = {"minVal": float(model.minVal), "maxVal": float(model.maxVal), "gamma": float(model.gamma),
                           "alphaPow": float(model.alphaPow), "bounds": [float(v) for v in model.bounds],
                           "isIso": bool(model.isIso), "dataPos": int(model.dataPos)}
        samples.append(rec)

    # slerp and easing tables
    rng = np.random.default_rng(7)
    slerp = []
    for _ in range(12):
        a, b = rng.normal(size=4), rng.normal(size=4)
        if _ % 4 == 3:
            b = a + 1e-3 * rng.normal(size=4)   # the nearly-parallel branch
        t = float(rng.uniform())
        q = km.quaternion_slerp(Q(*a), Q(*b), t)
        slerp.append({"a": a.tolist(), "b": b.tolist(), "t": t, "q": [float(v) for v in q.data],
                      "prod_ab": [float(v) for v in (Q(*a) * Q(*b)).data],
                      "rot4_a": Q(*a).toRotation4().tolist()})
    ease = [{"a": a, "x": x, "y": float(km.create_interp_func(a)(x))}
            for a in (0, .5, 3., 10.) for x in (0., .1, .5, .77, 1.)]

    out = {"generator": "tests/golden/make_keyframe_golden.py (reference keyframe_model / transform_model)",
           "keyframes_json": text, "n_frames": n_frames, "samples": samples, "slerp": slerp, "ease": ease}
    with open(os.path.join(HERE, "keyframes_ref.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "keyframes_ref.json"), len(samples), "samples")


if __name__ == "__main__":
    main()
