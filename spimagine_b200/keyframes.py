"""Keyframe camera paths and their batch rendering -- the caller that feeds the ray-march path a whole animation
(SURVEY.md 8f-3).

Host-side mirror, without Qt, of
  * spimagine/models/keyframe_model.py:19-328  create_interp_func, TransformData (+ interp), KeyFrame, KeyFrameList
    (getTransform, the JSON encoder / decoder and its file format);
  * spimagine/models/transform_model.py:262-312  how a TransformData becomes modelView / projection
    (TransformModel.update / setPerspective / getUnscaledModelView);
  * spimagine/gui/glwidget.py:610-636 + gui/keyframe_view.py:644-668  what the "record" loop does per frame:
    keyList.getTransform(recordPos / nFrames) -> transform.fromTransformData -> renderer setters -> render.
Same names and numbers, so a keyframe file saved by spimagine's GUI plays here and yields the cameras the GUI
would have used.  The reference renders every recorded frame synchronously (render, blocking read-back, grab the
GL frame buffer) on a 50 ms timer; `render_keyframes` keeps the device busy instead: consecutive frames of one
render method go through VolumeRenderer.render_sequence (frame i+1 renders while frame i travels to pinned host
memory), time points are uploaded only when `dataPos` changes.
"""
import bisect
import json
import os

import numpy as np

from .utils.quaternion import Quaternion, quaternion_slerp
from .utils.transform_matrices import mat4_translate, mat4_scale, mat4_perspective, mat4_ortho

__all__ = ["create_interp_func", "TransformData", "KeyFrame", "KeyFrameList", "KeyFrameEncoder",
           "KeyFrameDecoder", "camera_of", "apply_transform", "render_keyframes", "record_keyframes",
           "keyframe_data_positions", "keyframe_times", "frame_name"]


def create_interp_func(a):
    """keyframe_model.py:19-39: easing on [0, 1] with f(0) = 0, f(1) = 1; a = 0 is linear, larger a is more elastic."""
    if a == 0:
        return lambda x: x
    return lambda x: .5 * (1 + np.arctan(2 * a * (x - .5)) / np.arctan(a))


_TD_DEFAULTS = (("zoom", 1), ("dataPos", 0), ("minVal", 0.), ("maxVal", 100.), ("gamma", 1.),
                ("translate", [0, 0, 0]), ("bounds", [-1, 1, -1, 1, -1, 1]), ("isBox", True), ("isIso", False),
                ("alphaPow", 0.), ("isSlice", False), ("slicePos", 0), ("sliceDim", 0))


class TransformData(object):
    """keyframe_model.py:54-170: everything a keyframe stores about the view."""

    def __init__(self, quatRot=None, **kw):
        unknown = set(kw) - set(k for k, _ in _TD_DEFAULTS)
        if unknown:
            raise TypeError("unexpected arguments %s" % sorted(unknown))
        vals = dict(_TD_DEFAULTS)
        vals.update(kw)
        self.setData(quatRot=Quaternion() if quatRot is None else quatRot, **vals)

    def setData(self, quatRot, zoom, dataPos, minVal, maxVal, gamma, translate, bounds, isBox, isIso, alphaPow,
                isSlice, slicePos, sliceDim):
        self.quatRot = Quaternion.copy(quatRot)
        self.zoom, self.dataPos = zoom, dataPos
        self.minVal, self.maxVal, self.gamma = minVal, maxVal, gamma
        self.bounds = np.array(bounds)
        self.isBox, self.isIso = isBox, isIso
        self.alphaPow = alphaPow
        self.translate = np.array(translate)
        self.isSlice, self.slicePos, self.sliceDim = isSlice, slicePos, sliceDim

    def __repr__(self):
        fields = ["quatRot = %s" % self.quatRot] + ["%s = %s" % (
            k, repr(getattr(self, k)).replace("array", "np.array")) for k, _ in _TD_DEFAULTS]
        return "TransformData(%s)" % ",\n              ".join(fields)

    @classmethod
    def interp(cls, x1, x2, lam, f=create_interp_func(0.)):
        """keyframe_model.py:130-170.  Rotation by slerp; zoom, window, gamma, bounds, alphaPow, translate linearly;
        dataPos rounded, slicePos truncated; isBox / isIso / isSlice / sliceDim are taken from the left keyframe."""
        t = f(lam)
        u = 1. - t
        # the same expressions as cls(quatRot=..., zoom=(1 - t) * x1.zoom + t * x2.zoom, ...), assigned without the
        # keyword round trip of the constructor (one TransformData per recorded frame)
        td = cls.__new__(cls)
        td.quatRot = quaternion_slerp(x1.quatRot, x2.quatRot, t)
        td.zoom = u * x1.zoom + t * x2.zoom
        td.dataPos = int(np.round(u * x1.dataPos + t * x2.dataPos))
        td.minVal = u * x1.minVal + t * x2.minVal
        td.maxVal = u * x1.maxVal + t * x2.maxVal
        td.gamma = u * x1.gamma + t * x2.gamma
        td.translate = u * x1.translate + t * x2.translate
        td.bounds = u * x1.bounds + t * x2.bounds
        td.isBox, td.isIso = x1.isBox, x1.isIso
        td.alphaPow = u * x1.alphaPow + t * x2.alphaPow
        td.isSlice, td.slicePos, td.sliceDim = x1.isSlice, int(u * x1.slicePos + t * x2.slicePos), x1.sliceDim
        return td


class KeyFrame(object):
    """keyframe_model.py:173-182.  pos in [0, 1]; interp_elasticity eases the stretch up to the next keyframe."""

    def __init__(self, pos=0, transformData=None, interp_elasticity=0.):
        self.pos = pos
        self.transformData = TransformData() if transformData is None else transformData
        self.interp_elasticity = interp_elasticity

    def __repr__(self):
        return "Keyframe at t = %.3f (elasticity = %s) \n%s" % (self.pos, self.interp_elasticity, self.transformData)


class KeyFrameList(object):
    """keyframe_model.py:188-297 without the Qt signals.  `items`: id -> KeyFrame, `posdict`: pos -> id (kept as a
    plain dict here; `_order` holds the positions sorted, which is what the reference's SortedDict provides)."""

    def __init__(self):
        self._countID = 0
        self.posdict = {}
        self.items = {}
        self._order = []

    def __repr__(self):
        return "\n".join(str(self.items[self.posdict[p]]) for p in self._order) + "\n%s\n%s\n" % (
            [(p, self.posdict[p]) for p in self._order], list(self.items.keys()))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, ID):
        return self.items[ID]

    def _getNewID(self):
        self._countID += 1
        return self._countID - 1

    def _reindex(self):
        self._order = sorted(self.posdict)

    def addItem(self, frame=None):
        """A keyframe at a position that is already taken replaces the id stored there (the old KeyFrame stays in
        `items`, unreachable by position) -- SortedDict assignment in the reference (keyframe_model.py:218-219)."""
        frame = KeyFrame() if frame is None else frame
        newID = self._getNewID()
        if newID in self.items or newID in self.posdict.values():
            raise KeyError()
        self.items[newID] = frame
        self.posdict[frame.pos] = newID
        self._reindex()
        return newID

    def removeItem(self, ID):
        self.posdict.pop(self.pos_at_id(ID))
        self.items.pop(ID)
        self._reindex()

    def item_id_at(self, index):
        return self.posdict[self._order[index]]

    def item_at(self, index):
        return self.items[self.item_id_at(index)]

    def pos_at(self, index):
        return self._order[index]

    def pos_at_id(self, ID):
        for p in self._order:
            if self.posdict[p] == ID:
                return p
        raise ValueError("%s is not in list" % ID)

    def update_pos(self, ID, pos):
        if pos in self.posdict:
            print("pos already there:", pos)
            return
        self.posdict.pop(self.pos_at_id(ID))
        self.items[ID].pos = pos
        self.posdict[pos] = ID
        self._reindex()

    def distribute(self, pos_start, pos_end):
        """keyframe_model.py:260-265: spread dataPos linearly over the keyframes' positions."""
        for it in self.items.values():
            it.transformData.dataPos = int(pos_start + (pos_end - pos_start) * it.pos)

    def getTransform(self, pos):
        """keyframe_model.py:267-291.  Clamped outside the first / last keyframe, the keyframe itself on an exact
        position, otherwise interpolated with the easing of the keyframe to the left."""
        if pos < self._order[0]:
            return self.item_at(0).transformData
        if pos > self._order[-1]:
            return self.item_at(-1).transformData
        if pos in self.posdict:
            return self.items[self.posdict[pos]].transformData
        ind = bisect.bisect_right(self._order, pos)
        left, right = self.item_at(ind - 1), self.item_at(ind)
        if np.abs(right.pos - left.pos) < 1.e-7:
            lam = 0.
        else:
            lam = (1. * pos - left.pos) / (right.pos - left.pos)
        return TransformData.interp(left.transformData, right.transformData, lam,
                                    create_interp_func(left.interp_elasticity))

    # ---- the reference's file format (keyframe_model.py:293-379, written by gui/keyframe_view.py:691-707) ----
    def _to_JSON(self):
        return json.dumps(self, indent=4, sort_keys=True, cls=KeyFrameEncoder)

    @classmethod
    def _from_JSON(cls, jsonStr):
        return json.loads(jsonStr, cls=KeyFrameDecoder)

    def save_to_JSON(self, fName):
        with open(fName, "w") as f:
            f.write(self._to_JSON())

    @classmethod
    def load_from_JSON(cls, fName):
        with open(fName, "r") as f:
            return cls._from_JSON(f.read())


class KeyFrameEncoder(json.JSONEncoder):
    """{"_countID": n, "items": {id: {"pos", "transformData": {...}, "interp_elasticity"}}, "posdict": {pos: id}}"""

    def default(self, obj):
        if isinstance(obj, np.ndarray) and obj.ndim == 1:
            return obj.tolist()
        if isinstance(obj, Quaternion):
            return obj.data.tolist()
        if isinstance(obj, KeyFrameList):
            return {"_countID": obj._countID, "items": obj.items,
                    "posdict": dict((p, obj.posdict[p]) for p in obj._order)}
        if isinstance(obj, (KeyFrame, TransformData)):
            return obj.__dict__
        if isinstance(obj, np.generic):
            return obj.item()
        return json.JSONEncoder.default(self, obj)


class KeyFrameDecoder(json.JSONDecoder):
    """Keys a file lacks keep TransformData's defaults (files of older spimagine versions).  As in the reference
    (keyframe_model.py:365-367) a keyframe's interp_elasticity is NOT restored from the file: loaded paths are
    linear between keyframes.  `KeyFrameDecoder.keep_elasticity = True` restores it instead."""
    keep_elasticity = False

    def decode(self, s, classname=""):
        if classname == "":
            dec = json.JSONDecoder.decode(self, s)
            ret = KeyFrameList()
            ret._countID = dec["_countID"]
            ret.posdict = dict((float(k), int(v)) for k, v in dec["posdict"].items())
            ret.items = dict((int(k), self._frame(v)) for k, v in dec["items"].items())
            ret._reindex()
            return ret
        raise ValueError(classname)

    def _frame(self, v):
        t = TransformData()
        t.__dict__.update(v["transformData"])
        t.quatRot = Quaternion(*t.quatRot)
        t.bounds = np.array(t.bounds)
        t.translate = np.array(t.translate)
        return KeyFrame(v["pos"], t, v.get("interp_elasticity", 0.) if self.keep_elasticity else 0.)


# ------------------------------------------------------------------ TransformData -> renderer state
def _projection_of(isPerspective=True):
    """the projection TransformModel.setPerspective installs (transform_model.py:303-312): it depends on the mode only"""
    return mat4_perspective(60., 1., .1, 10) if isPerspective else mat4_ortho(-2., 2., -2., 2., -1.5, 1.5)


def camera_of(transformData, isPerspective=True):
    """-> (modelView, projection) exactly as TransformModel hands them to the renderer
    (transform_model.py:262-312: update, setPerspective, getUnscaledModelView; glwidget.py:615-616)."""
    td = transformData
    zoom = np.clip(td.zoom, .3, 2)  # TransformModel.setZoom, transform_model.py:226-229
    if isPerspective:
        cameraZ = 4 * (1 - np.log(zoom) / np.log(2.))
        scaleAll = 1.
    else:
        cameraZ = 0.
        scaleAll = 2.5 ** (zoom - 1.)
    projection = _projection_of(isPerspective)
    model = mat4_scale(*[scaleAll] * 3)
    model = np.dot(model, td.quatRot.toRotation4())
    model = np.dot(model, mat4_translate(*td.translate))
    return np.dot(mat4_translate(0, 0, -cameraZ), model), projection


def apply_transform(renderer, transformData, isPerspective=True):
    """The setter calls of GLWidget.render / GLWidget.transform handlers for one TransformData
    (glwidget.py:362-366, 615-621); the occlusion parameters are not keyframed and stay as they are.
    Returns (modelView, method)."""
    td = transformData
    modelView, projection = camera_of(td, isPerspective)
    renderer.set_projection(projection)
    renderer.set_min_val(max(1.e-6, td.minVal))  # TransformModel.setMin, transform_model.py:205-206
    renderer.set_max_val(td.maxVal)
    renderer.set_gamma(td.gamma)
    renderer.set_alpha_pow(td.alphaPow)
    renderer.set_box_boundaries(list(td.bounds))
    renderer.set_modelView(modelView)
    return modelView, ("iso_surface" if td.isIso else "max_project")


def _static_key(td, source, isPerspective):
    """everything apply_transform / the time-point upload set from a TransformData except the modelView (the projection
    depends on the mode alone, which is the same for every frame of a record loop)"""
    pos = None if source is None else int(np.clip(td.dataPos, 0, len(source) - 1))
    return (pos, float(td.minVal), float(td.maxVal), float(td.gamma), float(td.alphaPow),
            tuple(float(b) for b in td.bounds), bool(isPerspective))


def keyframe_times(nFrames):
    """Key times of the record loop (keyframe_view.py:644-653): recordPos = 1 .. nFrames at recordPos / nFrames."""
    return [(k, 1. * k / nFrames) for k in range(1, nFrames + 1)]


def render_keyframes(renderer, keyList, nFrames, source=None, isPerspective=True, pinned=False, pipelined=True,
                     iso_planes=7, rank=0, world=1, use_source_units=True):
    """Generator over (recordPos, transformData, renderer) for the nFrames frames of the record loop;
    renderer.output / output_alpha (+ the iso planes) hold that frame when it is yielded.

    source      None: the renderer's resident volume is used for every frame; otherwise source[dataPos] is the
                (Nz, Ny, Nx) time point (a frames.FrameSource / GenericData / 4-D array): uploaded with update_data
                whenever the interpolated dataPos changes (DataModel.setPos -> GLWidget.dataModel_changed,
                glwidget.py:372-374), clipped to the source's length
    pinned      source[t] are page-locked arrays (asynchronous uploads)
    pipelined   runs of frames with the same render method go through render_sequence
    iso_planes  7: every result plane of an iso-surface frame is read back; 2: only output and output_alpha (what a
                recorded frame shows), the other planes stay on the device and read as None
    use_source_units   take the voxel size from source.stackUnits when the source has one (as the GUI does when a
                data model is loaded); False keeps what the caller set with renderer.set_units (spim_render -u)
    rank, world this process renders frames rank + 1, rank + 1 + world, ... of the loop (one process per GPU, every one
                with the whole volume or time series; no data-path collective: frames are independent)"""
    if not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    times = keyframe_times(nFrames)[rank::world]
    nFrames = len(times)
    tds = [keyList.getTransform(t) for _, t in times]
    state = {"pos": None}

    def prepare(td):
        if source is not None:
            pos = int(np.clip(td.dataPos, 0, len(source) - 1))
            if pos != state["pos"]:
                vol = source[pos]
                if not hasattr(renderer, "dataImg") or tuple(vol.shape[::-1]) != renderer.dataImg.shape \
                        or vol.dtype != renderer.dataImg.dtype:
                    renderer.set_data(vol)
                else:
                    renderer.update_data(vol, pinned=pinned)
                if use_source_units and hasattr(source, "stackUnits"):
                    units = source.stackUnits() if callable(source.stackUnits) else source.stackUnits
                    if units is not None:  # OverlayData / a bare GenericData carry none
                        renderer.set_units(units)
                state["pos"] = pos
        return apply_transform(renderer, td, isPerspective)

    i = 0
    while i < nFrames:
        method = "iso_surface" if tds[i].isIso else "max_project"
        j = i
        while j < nFrames and ("iso_surface" if tds[j].isIso else "max_project") == method:
            j += 1
        if not pipelined or tds[i].alphaPow != 0 or any(td.alphaPow != 0 for td in tds[i:j]):
            # attenuated max projections take the exact multi-pass kernel; render them one by one
            for k in range(i, j):
                prepare(tds[k])
                renderer.render(method=method)
                yield times[k][0], tds[k], renderer
        elif method == "max_project":
            # stretches of at least 4 frames in which only the camera moves (same time point, window, box, projection) go
            # to render_sequence as a LIST of modelViews: several frames per launch (spv_render_mip_batch).  Frames
            # between such stretches keep the two-frames-in-flight pipeline fed by a generator that applies each frame's
            # settings as it is pulled.
            keys_ = [_static_key(tds[k], source, isPerspective) for k in range(i, j)]
            run_end = {}
            k = i
            while k < j:
                e = k + 1
                while e < j and keys_[e - i] == keys_[k - i]:
                    e += 1
                run_end[k] = e
                k = e
            k0 = i
            while k0 < j:
                if run_end[k0] - k0 >= 4:
                    k1 = run_end[k0]
                    prepare(tds[k0])
                    views = [camera_of(tds[k], isPerspective)[0] for k in range(k0, k1)]
                else:
                    k1 = k0
                    while k1 < j and run_end[k1] - k1 < 4:
                        k1 = run_end[k1]
                    views = (prepare(tds[k])[0] for k in range(k0, k1))
                for k, r in zip(range(k0, k1), renderer.render_sequence(views, method=method)):
                    yield times[k][0], tds[k], r
                k0 = k1
        else:
            views = (prepare(tds[k])[0] for k in range(i, j))
            for k, r in zip(range(i, j), renderer.render_sequence(views, method=method, iso_planes=iso_planes)):
                yield times[k][0], tds[k], r
        i = j


def frame_name(recordPos, nFrames):
    """keyframe_view.py:653: output_<recordPos zero-filled to the digits of nFrames>.png"""
    return "output_%s.png" % str(recordPos).zfill(int(np.log10(nFrames) + 1))


def keyframe_data_positions(keyList, nFrames, n_time_points, rank=0, world=1):
    """Time points the record loop asks for, in order and without immediate repeats: the `frames` argument of a
    frames.FrameSource that feeds render_keyframes(source=...) (its reader prefetches in exactly this order)."""
    out = []
    for _, t in keyframe_times(nFrames)[rank::world]:
        pos = int(np.clip(keyList.getTransform(t).dataPos, 0, n_time_points - 1))
        if not out or out[-1] != pos:
            out.append(pos)
    return out


def record_keyframes(renderer, keyList, nFrames, dirName, lut=None, mode_black=True, **kw):
    """The GUI's record button without the GUI: every frame of the path as dirName/output_NNN.png.  The reference
    grabs the GL frame buffer (LUT-coloured, alpha-blended onto the background, texture.frag:8-38); here the same
    colouring runs on the device (set_lut / output_rgba) and the packed RGBA8 image is what crosses PCIe."""
    from PIL import Image
    os.makedirs(dirName, exist_ok=True)
    if lut is not None:
        renderer.set_lut(lut)
    names = []
    for pos, td, r in render_keyframes(renderer, keyList, nFrames, pipelined=False, **kw):  # kw: rank=, world=, source=, ...
        rgba = r.output_rgba(mode_black=mode_black)
        name = os.path.join(dirName, frame_name(pos, nFrames))
        Image.fromarray(np.ascontiguousarray(rgba[::-1]), "RGBA").save(name)
        names.append(name)
    return names
