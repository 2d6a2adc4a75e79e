"""spimagine_b200.transform_model.TransformModel against the reference's own class: tests/golden/transform_ref.json
holds what spimagine/models/transform_model.py (Qt stubbed) stores, emits and hands out after every call of a 41-step
script (make_transform_golden.py); the same script is replayed here.  CPU only."""
import importlib.util
import json
import os
import types

import numpy as np
import pytest

from spimagine_b200 import keyframes
from spimagine_b200.transform_model import Signal, TransformModel
from spimagine_b200.utils.quaternion import Quaternion

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _generator():
    spec = importlib.util.spec_from_file_location("make_transform_golden", os.path.join(GOLDEN, "make_transform_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _recording(model):
    log = []
    for name, sig in vars(model).items():
        if isinstance(sig, Signal):
            sig.connect(lambda *a, _n=name: log.append([_n, [x.item() if isinstance(x, np.generic) else x for x in a]]))
    return log


def test_replay_of_the_reference_script():
    with open(os.path.join(GOLDEN, "transform_ref.json")) as f:
        ref = json.load(f)
    gen = _generator()
    assert json.loads(json.dumps(gen.SCRIPT)) == ref["script"]
    m = TransformModel()
    got = gen.snapshot(m, [])
    for key in ("state", "unscaled", "modelView", "projection"):
        assert got[key] == ref["init"][key], key
    log = _recording(m)
    fake = gen.FakeDataModel()
    mod = types.SimpleNamespace(TransformData=keyframes.TransformData, Quaternion=Quaternion)
    steps = gen.run(m, mod, log, fake)
    assert fake.positions == ref["data_positions"]
    assert len(steps) == len(ref["steps"]) == len(ref["script"])
    for i, (g, w) in enumerate(zip(steps, ref["steps"])):
        what = "step %d %s" % (i, ref["script"][i])
        assert g["signals"] == w["signals"], what
        assert g["state"] == w["state"], what
        for key in ("unscaled", "modelView", "projection"):           # float64, same operations: exact
            assert g[key] == w[key], (what, key)


def test_camera_equals_the_keyframe_path():
    """what the model hands to the renderer is what keyframes.camera_of computes from its TransformData"""
    m = TransformModel()
    m.addRotation(.4, 0, 1, 0)
    m.setZoom(1.6)
    m.setTranslate(.1, .2, -.1)
    for persp in (True, False):
        m.setPerspective(persp)
        mv, proj = keyframes.camera_of(m.toTransformData(), persp)
        assert np.array_equal(mv, m.getUnscaledModelView()) and np.array_equal(proj, m.getProjection())


def test_signals_and_errors():
    m = TransformModel()
    seen = []
    m._transformChanged.connect(lambda: seen.append("t"))
    m._rotationChanged.connect(lambda: seen.append("r"))
    m.setRotation(.1, 1, 0, 0)
    assert seen == ["r", "t"]
    m._rotationChanged.disconnect()
    m.setRotation(.2, 1, 0, 0)
    assert seen == ["r", "t", "t"]
    with pytest.raises(ValueError):
        m.setSliceDim(3)
    with pytest.raises(AttributeError):
        m.setPos(1)                      # no data model attached: as in the reference


def test_apply_makes_the_widgets_setter_calls():
    calls = []

    class Recorder(object):
        def __getattr__(self, name):
            return lambda *a: calls.append((name, a))

    m = TransformModel()
    m.reset(5., 900., [.2, .2, .8])
    m.setIso(True)
    m.setBounds(-1, 1, -.5, .5, -1, 1)
    assert m.apply(Recorder()) == "iso_surface"
    d = dict(calls)
    assert calls[-1][0] == "set_modelView" and np.array_equal(d["set_modelView"][0], m.getUnscaledModelView())
    assert d["set_min_val"] == (5.,) and d["set_max_val"] == (900.,) and d["set_units"] == ([.2, .2, .8],)
    assert d["set_box_boundaries"] == ([-1., 1., -.5, .5, -1., 1.],)
    assert d["set_occ_strength"] == (.15,) and d["set_occ_radius"] == (21,) and d["set_occ_n_points"] == (31,)
    assert np.array_equal(d["set_projection"][0], m.getProjection())


def test_spin_is_the_rotate_timer(tmp_path, monkeypatch):
    """gui/mainwidget.py:760-765: every tick adds a rotation of half angle -0.02 about the configured spin axis"""
    monkeypatch.setenv("SPIMAGINE_CONFIG", str(tmp_path / "none"))
    m, ref = TransformModel(), TransformModel()
    ticks = []
    m._rotationChanged.connect(lambda: ticks.append(1))
    views = list(m.spin(5))
    for v in views:
        ref.addRotation(-.02, 0, 1, 0)                 # spin_axis defaults to y
        assert np.array_equal(v, ref.getUnscaledModelView())
    assert len(ticks) == 5 and np.array_equal(m.quatRot.data, ref.quatRot.data)
    # five ticks turn the view by 5 * 0.04 rad about y
    rot = m.quatRot.toRotation4()[:3, :3]
    assert np.allclose(np.arccos(rot[0, 0]), .2, atol=1e-12) and np.allclose(rot[1], [0, 1, 0])
    (tmp_path / "cfg").write_text("spin_axis = 2\n")
    monkeypatch.setenv("SPIMAGINE_CONFIG", str(tmp_path / "cfg"))
    m = TransformModel()
    list(m.spin(3))
    assert np.allclose(m.quatRot.toRotation4()[2, :3], [0, 0, 1])           # about z now
    m = TransformModel()
    list(m.spin(2, angle=.1, axis=0))
    assert np.allclose(m.quatRot.toRotation4()[0, :3], [1, 0, 0])
