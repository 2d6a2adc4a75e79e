"""The view state in front of the renderer, without Qt: spimagine/models/transform_model.py:23-349 (TransformModel).

In the reference every setter of TransformModel emits Qt signals and GLWidget re-renders on `_transformChanged`
(gui/glwidget.py:362-366, 610-636).  A headless caller -- a rotation sweep, a batch job, a test -- needs the same
state machine (what a setter stores, clamps and resets, which camera comes out) but no event loop, so here a signal is
a list of callbacks.  The names of the setters, attributes and signals are the reference's; `apply(renderer)` is the
one addition: the setter calls GLWidget makes before `renderer.render()`.

Pinned against the reference's own class (Qt stubbed) by tests/golden/make_transform_golden.py ->
tests/golden/transform_ref.json: the state, the modelView / projection matrices and the emitted signals after a
scripted sequence of setter calls.
"""
import logging

import numpy as np

from .keyframes import TransformData
from .utils.quaternion import Quaternion
from .utils.transform_matrices import mat4_ortho, mat4_perspective, mat4_scale, mat4_translate

logger = logging.getLogger(__name__)

__all__ = ["Signal", "TransformModel"]


class Signal(object):
    """connect / disconnect / emit of a Qt signal, called synchronously in connection order"""

    def __init__(self, name=""):
        self.name = name
        self._slots = []

    def connect(self, slot):
        self._slots.append(slot)

    def disconnect(self, slot=None):
        if slot is None:
            del self._slots[:]
        else:
            self._slots.remove(slot)

    def emit(self, *args):
        for slot in list(self._slots):
            slot(*args)


# signal name -> what it carries (transform_model.py:24-45)
_SIGNALS = ("_maxChanged", "_minChanged", "_gammaChanged", "_boxChanged", "_isoChanged", "_interpChanged",
            "_perspectiveChanged", "_rotationChanged", "_translateChanged", "_slicePosChanged", "_sliceDimChanged",
            "_boundsChanged", "_transformChanged", "_stackUnitsChanged", "_alphaPowChanged")


def _same(old, new):
    if isinstance(old, np.ndarray):
        return np.array_equal(old, new)
    if isinstance(old, Quaternion):
        return np.array_equal(old.data, new.data)
    return old == new


class TransformModel(object):
    def __init__(self):
        for name in _SIGNALS:
            setattr(self, name, Signal(name))
        self.reset()

    # -- plumbing -----------------------------------------------------------------------------------------------
    def _update_value(self, name, newval):
        """store newval under `name`; True if it differs from what was there (transform_model.py:51-71)"""
        if hasattr(self, name) and _same(getattr(self, name), newval):
            return False
        setattr(self, name, newval)
        return True

    def _changed(self, signal=None, *args):
        if signal is not None:
            getattr(self, signal).emit(*args)
        self._transformChanged.emit()

    def setModel(self, dataModel):
        self.dataModel = dataModel

    # -- reset / center (transform_model.py:76-138) ------------------------------------------------------------------
    def reset(self, minVal=0., maxVal=256., stackUnits=None):
        self.dataPos = self.slicePos = self.sliceDim = 0
        self.zoom = 1.
        self.setIso(False)
        self.isPerspective = True
        self.setPerspective()
        self.setValueScale(minVal, maxVal)
        self.setGamma(1.)
        self.setAlphaPow(0)
        self.setBox(True)
        self.setInterpolate(True)
        self.setOccStrength()
        self.setOccRadius()
        self.setOccNPoints()
        self.eye_dist_proj = self.eye_dist_cam = 0
        if not hasattr(self, "isSlice"):
            self.setShowSlice(False)
        self.setStackUnits(*(stackUnits if stackUnits else [.1, .1, .1]))
        self.center()

    def center(self):
        self.quatRot = Quaternion()
        self.cameraZ, self.zoom, self.scaleAll = 5., 1., 1.
        self.setBounds(-1, 1., -1, 1, -1, 1)
        self.setTranslate(0, 0, 0)
        self.update()
        self._changed()

    # -- setters that only fire when the value moved ----------------------------------------------------------------
    def setIso(self, isIso):
        if self._update_value("isIso", isIso):
            self._changed("_isoChanged", isIso)

    def setInterpolate(self, is_interpolate):
        if self._update_value("is_interpolate", is_interpolate):
            self._changed("_interpChanged", is_interpolate)

    def setOccStrength(self, occ_strength=.15):
        if self._update_value("occ_strength", occ_strength):
            self._changed()

    def setOccRadius(self, val=21):
        if self._update_value("occ_radius", val):
            self._changed()

    def setOccNPoints(self, val=31):
        if self._update_value("occ_n_points", val):
            self._changed()

    def setTranslate(self, x, y, z):
        if self._update_value("translate", np.array([x, y, z])):
            self._changed("_translateChanged", x, y, z)

    # -- setters that always fire -----------------------------------------------------------------------------------
    def addTranslate(self, dx, dy, dz):
        self.translate = self.translate + np.array([dx, dy, dz])
        self._changed("_translateChanged", *self.translate)

    def setBounds(self, x1, x2, y1, y2, z1, z2):
        self.bounds = np.array([x1, x2, y1, y2, z1, z2])
        self._changed("_boundsChanged", x1, x2, y1, y2, z1, z2)

    def setShowSlice(self, isSlice=True):
        self.isSlice = isSlice
        self._changed()

    def setSliceDim(self, dim):
        if not 0 <= dim < 3:
            raise ValueError("dim should be in [0,1,2]!")
        self.sliceDim = dim
        self._changed("_sliceDimChanged", dim)

    def setSlicePos(self, pos):
        self.slicePos = pos
        self._changed("_slicePosChanged", pos)

    def setPos(self, pos):
        """needs a data model (setModel): AttributeError without one, as in the reference (:177-181)"""
        self.dataPos = pos
        self.dataModel.setPos(pos)
        self._changed()

    def setGamma(self, gamma):
        self.gamma = gamma
        self._changed("_gammaChanged", gamma)

    def setAlphaPow(self, alphaPow):
        self.alphaPow = alphaPow
        self._changed("_alphaPowChanged", alphaPow)

    def setValueScale(self, minVal, maxVal):
        self.setMin(minVal)
        self.setMax(maxVal)

    def setMin(self, minVal):
        self.minVal = max(1.e-6, minVal)  # never 0: the window is divided by in the kernels
        self._changed("_minChanged", self.minVal)

    def setMax(self, maxVal):
        self.maxVal = maxVal
        self._changed("_maxChanged", maxVal)

    def setStackUnits(self, px, py, pz):
        self.stackUnits = px, py, pz
        self._changed("_stackUnitsChanged", px, py, pz)

    def setBox(self, isBox=True):
        self.isBox = isBox
        self._changed("_boxChanged", isBox)

    def setZoom(self, zoom=1.):
        self.zoom = np.clip(zoom, .3, 2)
        self.update()
        self._changed()

    # -- rotation: half angles, i.e. addRotation(a, axis) turns by 2a (transform_model.py:233-248) --------------------
    @staticmethod
    def _axis_quaternion(angle, x, y, z):
        s = np.sin(angle)
        return Quaternion(np.cos(angle), s * x, s * y, s * z)

    def addRotation(self, angle, x, y, z, from_left=True):
        q = self._axis_quaternion(angle, x, y, z)
        self.setQuaternion(q * self.quatRot if from_left else self.quatRot * q)

    def setRotation(self, angle, x, y, z):
        self.setQuaternion(self._axis_quaternion(angle, x, y, z))

    def setQuaternion(self, quat):
        self.quatRot = Quaternion.copy(quat)
        self._changed("_rotationChanged")

    def setEyeDistProj(self, eye_dist_proj=0):
        self.eye_dist_proj = eye_dist_proj
        self.update()
        self._changed()

    def setEyeDistCam(self, eye_dist_cam=0.):
        self.eye_dist_cam = eye_dist_cam
        self.update()
        self._changed()

    # -- camera (transform_model.py:262-312) -----------------------------------------------------------------------
    def update(self):
        if self.isPerspective:
            self.cameraZ = 4 * (1 - np.log(self.zoom) / np.log(2.))
            self.scaleAll = 1.
        else:
            self.cameraZ = 0.
            self.scaleAll = 2.5 ** (self.zoom - 1.)

    def setPerspective(self, isPerspective=True):
        self.isPerspective = isPerspective
        self.projection = mat4_perspective(60., 1., .1, 10) if isPerspective else mat4_ortho(-2., 2., -2., 2., -1.5, 1.5)
        self.update()
        self._changed("_perspectiveChanged", isPerspective)

    def getProjection(self):
        return self.projection

    def getUnscaledModelView(self):
        """what the render kernels get (the renderer scales by the stack units itself, volumerender.py:299-325)"""
        model = np.dot(mat4_scale(self.scaleAll, self.scaleAll, self.scaleAll), self.quatRot.toRotation4())
        return np.dot(mat4_translate(0, 0, -self.cameraZ), np.dot(model, mat4_translate(*self.translate)))

    def getModelView(self):
        """with the volume's own scale: for drawing GL primitives in the rendered volume's frame"""
        modelView = self.getUnscaledModelView()
        if hasattr(self, "dataModel"):
            Nz, Ny, Nx = self.dataModel.size()[1:]
            extent = [d * N for d, N in zip(self.stackUnits, (Nx, Ny, Nz))]
            modelView = np.dot(modelView, mat4_scale(*[1. * e / max(extent) for e in extent]))
        return modelView

    # -- keyframes (transform_model.py:314-349) ---------------------------------------------------------------------
    _TD_FIELDS = ("zoom", "dataPos", "minVal", "maxVal", "gamma", "translate", "bounds", "isBox", "isIso", "alphaPow",
                  "isSlice", "slicePos", "sliceDim")

    def fromTransformData(self, transformData):
        td = transformData
        self.setQuaternion(td.quatRot)
        self.setZoom(td.zoom)
        self.setPos(td.dataPos)
        self.setBounds(*td.bounds)
        self.setBox(td.isBox)
        self.setIso(td.isIso)
        self.setAlphaPow(td.alphaPow)
        self.setTranslate(*td.translate)
        self.setValueScale(td.minVal, td.maxVal)
        self.setGamma(td.gamma)
        self.setValueScale(td.minVal, td.maxVal)
        self.setShowSlice(td.isSlice)
        self.setSlicePos(td.slicePos)
        self.setSliceDim(td.sliceDim)

    def toTransformData(self):
        return TransformData(quatRot=self.quatRot, **{k: getattr(self, k) for k in self._TD_FIELDS})

    # -- the "spin current view" button (gui/mainwidget.py:750-765) -----------------------------------------------
    def spin(self, n_frames, angle=-.02, axis=None):
        """Generator over n_frames ticks of the GUI's rotate timer: each tick is addRotation(-.02, axis) -- a turn by
        0.04 rad, the angle being a half angle -- about the axis `spin_axis` of ~/.spimagine names (0 / 1 / 2 = x / y /
        z, default y), and yields the modelView the renderer gets.  157 ticks are one revolution.
        `renderer.render_sequence(list(model.spin(n)))` renders the sweep with frames in flight."""
        if axis is None:
            from . import config
            axis = config.get_param("spin_axis")
        direction = [0, 0, 0]
        direction[int(axis)] = 1
        for _ in range(int(n_frames)):
            self.addRotation(angle, *direction)
            yield self.getUnscaledModelView()

    # -- the renderer (addition) ------------------------------------------------------------------------------------
    def apply(self, renderer):
        """The setter calls GLWidget makes on its renderer before render() (gui/glwidget.py:362-366, 610-636, 322-326):
        window, gamma, opacity, box, occlusion parameters, units, projection, modelView.  Returns the render method."""
        renderer.set_units(list(self.stackUnits))
        renderer.set_projection(self.getProjection())
        renderer.set_min_val(self.minVal)
        renderer.set_max_val(self.maxVal)
        renderer.set_gamma(self.gamma)
        renderer.set_alpha_pow(self.alphaPow)
        renderer.set_box_boundaries(list(self.bounds))
        renderer.set_occ_strength(self.occ_strength)
        renderer.set_occ_radius(self.occ_radius)
        renderer.set_occ_n_points(self.occ_n_points)
        renderer.set_modelView(self.getUnscaledModelView())
        return "iso_surface" if self.isIso else "max_project"
