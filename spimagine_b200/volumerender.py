"""VolumeRenderer -- drop-in for spimagine.volumerender.volumerender.VolumeRenderer on a B200.

Same class name, method names, argument meaning, defaults and error behaviour as the reference
(spimagine/volumerender/volumerender.py:59-547); the gputools / pyopencl plumbing underneath
(OCLProgram.run_kernel, OCLArray, OCLImage) is replaced by ctypes calls into libspimcuda.so
(include/spimcuda.h), hand-written CUDA for sm_100a.  There is no CPU path: constructing a renderer
without the library or without a CUDA device raises.

    rend = VolumeRenderer((400, 400))
    rend.set_data(d)                      # float32 / uint16 / uint8, shape (Nz, Ny, Nx)
    rend.set_units([1., 1., .1])
    rend.set_projection(mat4_perspective(60, 1., 1, 10))
    rend.set_modelView(np.dot(mat4_translate(0, 0, -7), mat4_scale(.7, .7, .7)))
    rend.render()                         # returns None, like the reference
    img = rend.output

Additions that the reference does not have (all keyword-only / opt-in, defaults keep reference behaviour):
    sampler="tmu" | "exact"      hardware-filtered texture fetches, or fp32 software trilinear that
                                 follows the reference loop operation by operation (parity runs)
    max_steps=200                the reference bakes config.__DEFAULTMAXSTEPS__ in at compile time
    int_filter="linear"|"nearest" what a LINEAR sampler means for integer volumes (undefined in OpenCL)
    pinned_outputs=False         True: rend.output* are views of pinned staging memory (no extra copy)
    data_min / data_max          global min / max of the resident volume, computed during upload
"""
from __future__ import absolute_import, print_function

import ctypes as C
import logging
import os
from time import time

import numpy as np
from scipy.linalg import inv
from scipy.linalg.lapack import get_lapack_funcs

from . import _lib, config
from .utils.transform_matrices import *  # noqa: F401,F403  (the reference module re-exports these too)
from .utils.transform_matrices import mat4_identity, mat4_perspective, mat4_scale

logger = logging.getLogger(__name__)

DEFAULT_MAX_STEPS = 200  # spimagine/config/config.py:27 "max_steps"


_getrf, _getri = get_lapack_funcs(("getrf", "getri"), (np.empty((4, 4), np.float64),))


def _inv4(a):
    """scipy.linalg.inv(a), which the reference calls per set_modelView (volumerender.py:312-313), without its
    per-call validation overhead (15 -> 4 us per 4x4 matrix): the same LAPACK getrf + getri on the same values.
    Inputs for which scipy takes another route (not float64, not finite, singular, lower triangular -> trtri) go
    through scipy.linalg.inv itself, so the result is bit-identical in every case (tests/test_host.py)."""
    a = np.asarray(a)
    if (a.dtype == np.float64 and a.shape == (4, 4) and np.isfinite(a).all()
            and (a[0, 1] != 0 or a[0, 2] != 0 or a[0, 3] != 0 or a[1, 2] != 0 or a[1, 3] != 0 or a[2, 3] != 0)):
        lu, piv, info = _getrf(a)
        if info == 0:
            r, info = _getri(lu, piv, lwork=64, overwrite_lu=1)
            if info == 0:
                return r
    return inv(a)


class _DataImage(object):
    """What callers read from `rend.dataImg`: .shape == (Nx, Ny, Nz) and .dtype
    (spimagine/gui/glwidget.py:333-336, volumerender.py:320)."""

    def __init__(self, shape_xyz, dtype):
        self.shape = tuple(int(s) for s in shape_xyz)
        self.dtype = np.dtype(dtype)


class VolumeRenderer(object):
    """renders a data volume by ray casting / max projection"""
    dtypes = [np.float32, np.uint16, np.uint8]
    # the reference maps these to "-D SAMPLER_FILTER=..." build options; here: texture filter mode
    interpolation_defines = {"linear": 1, "nearest": 0}

    def __init__(self, size=None, interpolation='linear', sampler="tmu", max_steps=None, device=None,
                 int_filter="linear", pinned_outputs=False):
        self._lib = _lib.load()
        self._ctx = _lib._CTX()
        if device is None:
            device = config.default_device()  # $SPIMAGINE_CUDA_DEVICE, else id_device of ~/.spimagine
        self.device = device
        self.isGPU = True
        self.max_steps = int(max_steps) if max_steps is not None else config.default_max_steps()
        self.pinned_outputs = bool(pinned_outputs)
        w, h = size if size else (200, 200)
        rc = self._lib.spv_create(int(device), int(w), int(h), C.byref(self._ctx))
        _lib.check(rc, None)
        # device memory is not the limit it was for OpenCL images; keep the reference's rule with the
        # free memory of this GPU so that set_data's stride-downsampling still exists
        self.memMax = .7 * 160e9
        self.rebuild_program(interpolation=interpolation)
        self.set_sampler(sampler)
        self.set_int_filter(int_filter)

        self._invM = np.zeros(16, np.float32)
        self._invP = np.zeros(16, np.float32)
        self._invM_ptr, self._invP_ptr = _lib.fp(self._invM), _lib.fp(self._invP)
        self.projection = np.zeros((4, 4))
        self.modelView = np.zeros((4, 4))
        self.width, self.height = int(w), int(h)
        self.reset_buffer(_realloc=False)

        self.set_dtype()
        self.set_gamma()
        self.set_max_val()
        self.set_min_val()
        self.set_occ_strength(.1)
        self.set_occ_radius(21)
        self.set_occ_n_points(30)
        self.set_alpha_pow()
        self.set_box_boundaries()
        self.set_units()
        self.set_modelView()
        self.set_projection()

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            try:
                self._fetch_iso_extras()  # results of the last iso render stay readable after close()
            except Exception:
                pass
            self._lib.spv_destroy(self._ctx)
            self._ctx = _lib._CTX()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.check(rc, self._ctx)

    # ------------------------------------------------------------------ configuration
    def rebuild_program(self, interpolation="linear"):
        """The reference rebuilds its OpenCL program with -D SAMPLER_FILTER=...; here the filter mode is a
        property of the texture object, so this only validates and records it."""
        if interpolation not in VolumeRenderer.interpolation_defines:
            raise KeyError("interpolation = '%s' not defined ,valid: %s" % (
                interpolation, list(VolumeRenderer.interpolation_defines.keys())))
        self.interpolation = interpolation
        self._check(self._lib.spv_set_interp(self._ctx, VolumeRenderer.interpolation_defines[interpolation]))
        self.proc = self  # callers only test for its presence

    def set_sampler(self, sampler="tmu"):
        codes = {"tmu": _lib.SAMPLER_TMU, "exact": _lib.SAMPLER_EXACT}
        if sampler not in codes:
            raise KeyError("sampler = '%s' not defined, valid: %s" % (sampler, list(codes.keys())))
        self.sampler = sampler
        self._check(self._lib.spv_set_sampler(self._ctx, codes[sampler]))

    def set_int_filter(self, int_filter="linear"):
        if int_filter not in ("linear", "nearest"):
            raise KeyError("int_filter = '%s' not defined, valid: ['linear', 'nearest']" % int_filter)
        self.int_filter = int_filter
        self._check(self._lib.spv_set_int_filter(self._ctx, int(int_filter == "linear")))

    def set_layout(self, layout="zpair"):
        """Storage of integer volumes from the next set_data on: "zpair" (layered {v[z], v[z+1]} texels, one
        bilinear fetch + fp32 z-lerp per sample; default) or "3d" (plain 3-D array, one trilinear fetch)."""
        codes = {"3d": 0, "zpair": 1}
        if layout not in codes:
            raise KeyError("layout = '%s' not defined, valid: %s" % (layout, list(codes.keys())))
        self.layout = layout
        self._check(self._lib.spv_set_layout(self._ctx, codes[layout]))
        self._need_alloc = True

    def set_mip_path(self, path="tmu"):
        """Kernel family of plain uint16 max projections: "tmu" (one hardware-filtered fetch per sample, default) or
        "smem" (TMA-staged shared-memory slabs, software trilinear sampling with fp32 weights; spv_set_mip_path)."""
        codes = {"tmu": 0, "smem": 1}
        if path not in codes:
            raise KeyError("mip path = '%s' not defined, valid: %s" % (path, list(codes.keys())))
        self._check(self._lib.spv_set_mip_path(self._ctx, codes[path]))
        self.mip_path = path

    def mip_path_used(self):
        """the kernel family the last max projection ran on: 'tmu' or 'smem'"""
        v = C.c_int()
        self._check(self._lib.spv_mip_path_used(self._ctx, C.byref(v)))
        return ("tmu", "smem")[v.value]

    def mip_kernel_name(self):
        return "spv::mip_smem_kernel<u16>" if self.mip_path_used() == "smem" else "spv::mip_fast_kernel<u16, linear>"

    def set_skipping(self, on=True):
        """Empty-space skipping on the min/max brick grids: True / False, or None for the default (on for
        iso surfaces, off for max projection).  Never changes the image."""
        self._check(self._lib.spv_set_skipping(self._ctx, -1 if on is None else int(bool(on))))

    def set_dtype(self, dtype=None):
        if hasattr(self, "dtype") and dtype is self.dtype:
            return
        if dtype is None:
            dtype = self.dtypes[0]
        if dtype in self.dtypes:
            self.dtype = dtype
        else:
            raise NotImplementedError("data type should be either %s not %s" % (self.dtypes, dtype))
        self.reset_buffer(_realloc=False)

    def resize(self, size):
        self.width, self.height = int(size[0]), int(size[1])
        self.reset_buffer()

    def reset_buffer(self, _realloc=True):
        self._iso_pending = False
        if _realloc:
            self._check(self._lib.spv_resize(self._ctx, self.width, self.height))
            self.__dict__.pop("_views", None)  # the pinned staging moved
        self.output = np.zeros((self.height, self.width), dtype=np.float32)
        self.output_alpha = np.zeros((self.height, self.width), dtype=np.float32)
        self.output_depth = np.zeros((self.height, self.width), dtype=np.float32)
        self.output_normals = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self.output_occlusion = np.zeros((self.height, self.width), dtype=np.float32)

    def _get_downsampled_data_slices(self, data):
        """in case data is bigger than the memory budget, returns the strided slices to render, else None"""
        nbytes = data.size * np.dtype(self.dtype).itemsize  # of the volume as it will be resident
        Nstep = int(np.ceil((1. * nbytes / self.memMax) ** (1. / 3)))
        slices = tuple(slice(0, d, Nstep) for d in data.shape)
        if Nstep > 1:
            logger.info("downsample image by factor of  %s" % Nstep)
            return slices
        return None

    def set_max_val(self, maxVal=0.):
        self.maxVal = maxVal

    def set_min_val(self, minVal=0.):
        self.minVal = minVal

    def set_gamma(self, gamma=1.):
        self.gamma = gamma

    def set_occ_strength(self, occ=.2):
        self.occ_strength = occ

    def set_occ_radius(self, rad=21):
        self.occ_radius = rad

    def set_occ_n_points(self, n_points=31):
        self.occ_n_points = n_points

    def set_alpha_pow(self, alphaPow=0.):
        self.alphaPow = alphaPow

    # ------------------------------------------------------------------ data
    def set_data(self, data, autoConvert=True, copyData=False):
        logger.debug("set_data")
        if not autoConvert and not data.dtype in self.dtypes:
            raise NotImplementedError("data type should be either %s not %s" % (self.dtypes, data.dtype))
        if data.dtype.type in self.dtypes:
            self.set_dtype(data.dtype.type)
            _data = data
        else:
            print("converting type from %s to %s" % (data.dtype.type, self.dtype))
            # the reference converts on the host here (astype); update_data converts on the device instead where
            # the element type allows it, so the array keeps its type until then
            _data = data if self._device_converts(data.dtype) else data.astype(self.dtype, copy=False)
        self.dataSlices = self._get_downsampled_data_slices(_data)
        if self.dataSlices is not None:
            self.set_shape(_data[self.dataSlices].shape[::-1])
        else:
            self.set_shape(_data.shape[::-1])
        t = time()
        self.update_data(_data, copyData=copyData)
        logger.debug("update data: %s ms" % (1000. * (time() - t)))
        self.update_matrices()

    @staticmethod
    def _device_converts(dtype):
        dtype = np.dtype(dtype)
        return dtype in _lib.SRC_CODES and dtype.isnative

    def set_shape(self, dataShape):
        """dataShape = (Nx, Ny, Nz); the device array is (re)allocated on the next update_data."""
        self.dataImg = _DataImage(dataShape, self.dtype)
        self._need_alloc = True

    def update_data(self, data, copyData=False, pinned=False):
        """pinned=True (addition): `data` lives in page-locked memory (spimagine_b200.pinned_empty): the upload is
        enqueued at PCIe rate and the call returns at once; do not touch `data` before the next render has
        returned (or sync())."""
        if self.dataSlices is not None:
            self._data = data[self.dataSlices].copy()
        else:
            self._data = data.copy() if copyData else data
        src_type = None
        if self._data.dtype != self.dtype:
            if self._device_converts(self._data.dtype):
                src_type = _lib.SRC_CODES[np.dtype(self._data.dtype)]  # converted inside the ingest pipeline
            else:
                self._data = self._data.astype(self.dtype, copy=False)
        host = np.ascontiguousarray(self._data)
        Nx, Ny, Nz = self.dataImg.shape
        if host.shape != (Nz, Ny, Nx):
            raise ValueError("data shape %s does not match the volume shape %s" % (host.shape, (Nz, Ny, Nx)))
        code = _lib.DTYPE_CODES[np.dtype(self.dtype)]
        if getattr(self, "_need_alloc", True):
            if src_type is None:
                rc = self._lib.spv_set_volume(self._ctx, host.ctypes.data, code, Nx, Ny, Nz)
            else:
                rc = self._lib.spv_set_volume_from(self._ctx, host.ctypes.data, src_type, code, Nx, Ny, Nz)
            self._need_alloc = False
        elif src_type is not None:
            rc = self._lib.spv_update_volume_from(self._ctx, host.ctypes.data, src_type)
        elif pinned and host is data:
            rc = self._lib.spv_update_volume_async(self._ctx, host.ctypes.data)
        else:
            rc = self._lib.spv_update_volume(self._ctx, host.ctypes.data)
        self._check(rc)

    def set_data_device(self, device_ptr, shape, dtype):
        """Addition: take the volume from DEVICE memory (C-order (Nz, Ny, Nx) at `device_ptr`, e.g.
        tensor.data_ptr() of a torch / cupy array on this GPU): frame sources that keep timepoints in HBM."""
        dtype = np.dtype(dtype)
        if dtype.type not in self.dtypes:
            raise NotImplementedError("data type should be either %s not %s" % (self.dtypes, dtype))
        self.set_dtype(dtype.type)
        self.dataSlices = None
        Nz, Ny, Nx = (int(s) for s in shape)
        self.set_shape((Nx, Ny, Nz))
        self._data = None
        self._check(self._lib.spv_set_volume_device(self._ctx, C.c_void_p(int(device_ptr)), _lib.DTYPE_CODES[dtype],
                                                    Nx, Ny, Nz))
        self._need_alloc = False
        self.update_matrices()

    def update_data_device(self, device_ptr, shape, dtype):
        """Addition: update_data (volumerender.py:279-294) from DEVICE memory -- a C-order (Nz, Ny, Nx) array of any
        element type the ingest path knows, of the current volume shape; converted to the volume's element type on the
        device like update_data's astype.  Whatever produced the array must have finished; the copy is enqueued on
        the renderer's stream."""
        dtype = np.dtype(dtype)
        if dtype not in _lib.SRC_CODES:
            raise NotImplementedError("element type %s cannot be converted on the device" % dtype)
        Nx, Ny, Nz = self.dataImg.shape
        if tuple(int(s) for s in shape) != (Nz, Ny, Nx) or getattr(self, "_need_alloc", True):
            raise ValueError("update_data_device needs a resident volume of shape %s (set_data first), got %s"
                             % ((Nz, Ny, Nx), tuple(shape)))
        self._data = None
        self._check(self._lib.spv_update_volume_device_from(self._ctx, C.c_void_p(int(device_ptr)),
                                                            _lib.SRC_CODES[dtype]))

    @property
    def data_min_max(self):
        """(min, max) of the resident volume: what GLWidget._get_min_max computes (gui/glwidget.py:328-344)."""
        lo, hi = C.c_float(), C.c_float()
        self._check(self._lib.spv_volume_minmax(self._ctx, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    # ------------------------------------------------------------------ camera
    def set_box_boundaries(self, boxBounds=[-1, 1, -1, 1, -1, 1]):
        self.boxBounds = np.array(boxBounds)

    def set_units(self, stackUnits=np.ones(3)):
        self.stackUnits = np.array(stackUnits)

    def set_projection(self, projection=mat4_perspective()):
        self.projection = projection
        self.update_matrices()

    def set_modelView(self, modelView=mat4_identity()):
        self.modelView = 1. * modelView
        self.update_matrices()

    def update_matrices(self):
        if hasattr(self, "dataImg"):
            # host cost matters here: this runs once per frame of a spin (set_modelView).  mScale only changes with
            # the volume shape or the units, invP only with the projection; the float32 copies the kernels take
            # live in two persistent buffers whose ctypes pointers are made once.
            # (the cache keys are the arrays' bytes: a caller may assign or modify stackUnits / projection in place)
            su = self.stackUnits
            key = (self.dataImg.shape, su.tobytes() if isinstance(su, np.ndarray) else tuple(float(u) for u in su))
            cached = getattr(self, "_mscale_of", None)
            if cached is None or cached[0] != key:
                cached = self._mscale_of = (key, self._stack_scale_mat())
            invM = _inv4(np.dot(self.modelView, cached[1]))
            proj = self.projection
            pkey = (proj.dtype, proj.tobytes()) if isinstance(proj, np.ndarray) else None
            cached = getattr(self, "_invP_of", None)
            if cached is not None and pkey is not None and cached[0] == pkey:
                invP = cached[1]
            else:  # same scipy.linalg.inv as the reference, evaluated once per distinct projection matrix
                invP = _inv4(proj)
                self._invP_of = (pkey, invP)
                self._invP[:] = invP.ravel()
            self._invM[:] = invM.ravel()  # float64 -> float32 like .astype(np.float32)
            self._check(self._lib.spv_set_matrices(self._ctx, self._invP_ptr, self._invM_ptr))

    def _stack_scale_mat(self):
        # scaling the data according to size and units
        Nx, Ny, Nz = self.dataImg.shape
        dx, dy, dz = self.stackUnits
        maxDim = max(d * N for d, N in zip([dx, dy, dz], [Nx, Ny, Nz]))
        return mat4_scale(1. * dx * Nx / maxDim, 1. * dy * Ny / maxDim, 1. * dz * Nz / maxDim)

    # ------------------------------------------------------------------ rendering
    def _box(self):
        return (C.c_float * 6)(*[float(b) for b in self.boxBounds])

    def _pinned_view(self, host, count):
        """numpy view of `count` floats of pinned staging at `host` (cached: building one costs ~15 us)."""
        key = (C.cast(host, C.c_void_p).value, count)
        cache = self.__dict__.setdefault("_views", {})
        v = cache.get(key)
        if v is None:
            if len(cache) > 8:
                cache.clear()
            v = cache[key] = np.ctypeslib.as_array(host, shape=(count,))
            # read-only: libspimcuda keeps track of which staging rows already hold the miss values and does not
            # copy them again (spv_set_tuning knob 9); callers get views (pinned_outputs=True) or copies of this
            v.flags.writeable = False
        return v

    def _fetch(self, planes):
        """One device->host transfer of the leading result planes [out | alpha | depth | occ | normals]."""
        n = self.width * self.height
        host = _lib._FP()
        self._check(self._lib.spv_read_pinned(self._ctx, planes, C.byref(host)))
        flat = self._pinned_view(host, planes * n)
        if not self.pinned_outputs:
            flat = flat.copy()
        return flat, n

    def _render_max_project(self, dtype=np.float32, numParts=1, currentPart=0):
        if dtype not in [np.uint16, np.uint8, np.float32]:
            raise NotImplementedError("wrong dtype: %s", dtype)
        p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma),
                           float(self.alphaPow), int(numParts), int(currentPart), int(self.max_steps), 0)
        # render + read-back in one call: finished bands of rows travel to pinned memory while the rest renders
        host = _lib._FP()
        self._check(self._lib.spv_render_mip_to_host(self._ctx, C.byref(p), 0, 1, C.byref(host)))
        n = self.width * self.height
        flat = self._pinned_view(host, 2 * n)
        if not self.pinned_outputs:
            flat = flat.copy()
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)
        # the reference reads back buf_depth here although max_project never writes it (garbage); zeros instead
        if self.output_depth is None or self.output_depth.shape != shape:
            self.output_depth = np.zeros(shape, np.float32)

    def _render_isosurface(self, raw_only=False):
        """iso surface with ambient occlusion: iso_surface -> normal blur -> occlusion -> blur -> shading.
        output and output_alpha are read back at once; output_depth / output_normals / output_occlusion are read
        back when they are first looked at (they stay on the device otherwise: 5 of the 7 result planes)."""
        p = _lib.IsoParams(self._box(), float(self.maxVal / 2), float(self.gamma), int(self.max_steps),
                           float(self.occ_strength), int(self.occ_radius), int(self.occ_n_points),
                           _lib.ISO_RAW_ONLY if raw_only else 0)
        # render + read-back in one call: the alpha plane travels while the screen-space passes run
        host = _lib._FP()
        self._check(self._lib.spv_render_iso_to_host(self._ctx, C.byref(p), 1, C.byref(host)))
        n = self.width * self.height
        flat = self._pinned_view(host, 2 * n)
        if not self.pinned_outputs:
            flat = flat.copy()
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)
        self._iso_pending = True

    # output_depth / output_normals / output_occlusion: plain attributes in the reference (volumerender.py:122-130);
    # here they are fetched from the device on first access after an iso-surface render
    def _fetch_iso_extras(self):
        if self.__dict__.get("_iso_pending") and getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._iso_pending = False
            flat, n = self._fetch(7)
            shape = (self.height, self.width)
            if self.pinned_outputs:  # the staging was rewritten: re-point the eager planes as well (same values)
                self.__dict__["output"] = flat[:n].reshape(shape)
                self.__dict__["output_alpha"] = flat[n:2 * n].reshape(shape)
            self.__dict__["_output_depth"] = flat[2 * n:3 * n].reshape(shape)
            self.__dict__["_output_occlusion"] = flat[3 * n:4 * n].reshape(shape)
            self.__dict__["_output_normals"] = flat[4 * n:7 * n].reshape(shape + (3,))

    def _lazy_get(self, name):
        self._fetch_iso_extras()
        return self.__dict__["_" + name]

    def _lazy_set(self, name, value):
        self._fetch_iso_extras()  # a pending read-back must not overwrite what the caller assigns
        self.__dict__["_" + name] = value

    output_depth = property(lambda self: self._lazy_get("output_depth"),
                            lambda self, v: self._lazy_set("output_depth", v))
    output_normals = property(lambda self: self._lazy_get("output_normals"),
                              lambda self, v: self._lazy_set("output_normals", v))
    output_occlusion = property(lambda self: self._lazy_get("output_occlusion"),
                                lambda self, v: self._lazy_set("output_occlusion", v))

    def render(self, data=None, stackUnits=None,
               minVal=None, maxVal=None, gamma=None,
               modelView=None, projection=None,
               boxBounds=None, return_alpha=False, method="max_project",
               numParts=1, currentPart=0):
        if data is not None:
            self.set_data(data)
        if maxVal is not None:
            self.set_max_val(maxVal)
        if minVal is not None:
            self.set_min_val(minVal)
        if gamma is not None:
            self.set_gamma(gamma)
        if stackUnits is not None:
            self.set_units(stackUnits)
        if modelView is not None:
            self.set_modelView(modelView)
        if projection is not None:
            self.set_projection(projection)
        # boxBounds and return_alpha are accepted and ignored, as in the reference (volumerender.py:508-547)
        if not hasattr(self, 'dataImg'):
            print("no data provided, set_data(data) before")
            return
        if modelView is None and not hasattr(self, 'modelView'):
            print("no modelView provided and set_modelView() not called before!")
            return
        if method == "max_project":
            self._render_max_project(self.dtype, numParts, currentPart)
        if method == "iso_surface":
            self._render_isosurface()
        if method == "iso_surface_raw":  # addition: the iso_surface kernel alone (parity tests)
            self._render_isosurface(raw_only=True)

    # ------------------------------------------------------------------ pipelined sequences (addition)
    def render_sequence(self, modelViews, method="max_project", depth=2, iso_planes=7, batch=None):
        """Generator over the frames of a camera path (a spin, a keyframe sequence): for every modelView of the
        iterable it yields `self` with output / output_alpha (and, for "iso_surface", output_depth /
        output_occlusion / output_normals) holding that frame.  Unlike calling render() per frame -- which, like
        the reference, blocks on the read-back of every frame (volumerender.py:388-390) -- frame i+1 is rendered
        while frame i is still being copied to pinned host memory (two output slots, spv_select_slot /
        spv_read_pinned_async / spv_wait_slot).  With pinned_outputs=True the yielded arrays are views of the
        slot's staging memory and are valid only until the NEXT frame is requested from the generator: resuming it
        issues frame i+2 into the slot of frame i (copy a frame that must outlive that); otherwise copies.
        iso_planes=2 reads back only what a display needs of an iso-surface frame (output, output_alpha: 8 of the
        28 bytes per pixel); output_depth / output_normals / output_occlusion are None for such frames.
        Max projections are rendered `batch` frames per launch (at most 16;
        batch=1: one launch per frame as above): the frames of a launch share the volume in L2 (spv_render_mip_batch),
        and a launch's frames are copied to pinned memory band by band while it renders.  All frames of a launch use the
        renderer's settings (projection, window, box, data) as they are when the launch is issued, so this is the
        default (self.batch_frames = 10 per launch) only when modelViews is a list / tuple / array; an iterator that
        changes the renderer as it is pulled (keyframes.render_keyframes) gets one launch per frame unless batch is
        given.  With pinned_outputs=True a yielded frame is then valid until the first frame of the launch after the
        next one is requested."""
        if not hasattr(self, 'dataImg'):
            print("no data provided, set_data(data) before")
            return
        if method not in ("max_project", "iso_surface"):
            raise KeyError("method = '%s' not defined, valid: ['max_project', 'iso_surface']" % method)
        if iso_planes not in (2, 7):
            raise ValueError("iso_planes must be 2 (output, alpha) or 7 (all result planes)")
        planes = 2 if method == "max_project" else iso_planes
        clear = method == "iso_surface" and planes == 2
        self._fetch_iso_extras()  # a deferred read-back of an earlier render() happens before the slots are reused
        # several frames per launch: only when pulling a view cannot change the renderer's state behind the frames
        # already collected for the launch -- a list / tuple / array of matrices, or a caller that asks for it (batch=k)
        materialized = isinstance(modelViews, (list, tuple, np.ndarray))
        if method == "max_project" and batch != 1 and (batch is not None or materialized):
            p = self._mip_params()
            if self._lib.spv_mip_batch_possible(self._ctx, C.byref(p)) == 1:
                for _ in self._render_sequence_batched(modelViews, p, batch or self.batch_frames):
                    yield self
                return
        if method == "iso_surface":
            for _ in self._render_sequence_iso(modelViews, planes, clear):
                yield self
            return
        # one launch per frame (float volumes with the copies switched off, skipping, the exact sampler, iterators that
        # change the renderer as they are pulled): two frames in flight over the two output slots
        pending = []  # slots in flight, oldest first
        i = 0
        # the frames of slot 1 run on a second stream: a frame starts in the tail of the one before (tuning knob 15)
        self._check(self._lib.spv_set_tuning(self._ctx, 15, 1))
        try:
            for M in modelViews:
                slot = i & 1
                self._check(self._lib.spv_select_slot(self._ctx, slot))
                self.set_modelView(M)
                p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma),
                                   float(self.alphaPow), 1, 0, int(self.max_steps), 0)
                self._check(self._lib.spv_render_mip_to_host(self._ctx, C.byref(p), 1, 0, None))
                pending.append(slot)
                i += 1
                if len(pending) == 2:
                    self._adopt_slot(pending.pop(0), planes, clear)
                    yield self
            while pending:
                self._adopt_slot(pending.pop(0), planes, clear)
                yield self
        finally:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:
                self._lib.spv_set_tuning(self._ctx, 15, 0)
                self._lib.spv_sync(self._ctx)
                self._lib.spv_select_slot(self._ctx, 0)

    def _render_sequence_iso(self, modelViews, planes, clear):
        """Iso-surface frames of a sequence: renders are issued four frames ahead of the frame that is handed out,
        read-backs two ahead.  A device slot is reused as soon as the copy out of it has been enqueued (the next search
        into it waits on the device until the frame's rectangle has left the slot -- for output + alpha that is a small
        copy into a device staging buffer, tuning knob 18, not the trip over the host link); a slot's pinned staging is
        rewritten only after the frame it held has been handed out and the consumer has come back.  The screen-space
        passes of frame i run beside the search of frame i + 1 (tuning knob 14).  With renders only three ahead the
        device ran dry for ~12 us per frame: frame i + 3 was enqueued after frame i's read-back had been waited for."""
        it = iter(modelViews)
        rendered, copied = [], []   # frame numbers: rendered but not yet read back / read-back enqueued, oldest first
        state = {"n": 0}
        lib, ctx = self._lib, self._ctx

        def render_next():
            try:
                M = next(it)
            except StopIteration:
                return False
            k = state["n"]
            self._check(lib.spv_select_slot(ctx, k & 1))
            self.set_modelView(M)
            p = _lib.IsoParams(self._box(), float(self.maxVal / 2), float(self.gamma), int(self.max_steps),
                               float(self.occ_strength), int(self.occ_radius), int(self.occ_n_points), 0)
            self._check(lib.spv_render_iso(ctx, C.byref(p)))
            rendered.append(k)
            state["n"] = k + 1
            return True

        def copy_next():
            k = rendered.pop(0)
            self._check(lib.spv_select_slot(ctx, k & 1))
            self._check(lib.spv_read_pinned_async(ctx, planes))
            copied.append(k)

        self._check(lib.spv_set_tuning(ctx, 14, 1))
        try:
            for _ in range(2):
                if render_next():
                    copy_next()
            render_next()
            if planes == 2:  # (whole frames, 7 planes, are bound by the host link: 519 us per frame three ahead, 623 four ahead)
                render_next()
            while copied:
                k = copied.pop(0)
                self._adopt_slot(k & 1, planes, clear)
                yield self
                if rendered:
                    copy_next()
                render_next()
        finally:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:
                lib.spv_set_tuning(ctx, 14, 0)
                lib.spv_sync(ctx)
                lib.spv_select_slot(ctx, 0)

    def _mip_params(self, numParts=1, currentPart=0):
        return _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma),
                              float(self.alphaPow), int(numParts), int(currentPart), int(self.max_steps), 0)

    batch_frames = 10

    def _invM_of(self, modelView):
        """float32 row-major inverse of modelView . stack scale: what update_matrices hands to the kernels"""
        return _inv4(np.dot(modelView, self._stack_scale_mat())).astype(np.float32).ravel()

    def render_batch(self, modelViews, to_host=True):
        """Enqueue ONE launch that renders a max projection for each of up to 16 modelViews (current projection, box,
        window, gamma, alphaPow, max_steps) and return the id of the set of planes it went to; collect with
        batch_frames_of(set).  Raises where spv_render_mip_batch does not apply (float volumes, exact sampler, ...)."""
        Ms = [np.asarray(M, dtype=np.float64) for M in modelViews]
        scale = self._stack_scale_mat()
        inv = np.ascontiguousarray(np.stack([_inv4(np.dot(M, scale)).ravel() for M in Ms]), dtype=np.float32)
        p = self._mip_params()
        used = C.c_int()
        self._check(self._lib.spv_render_mip_batch(self._ctx, C.byref(p), _lib.fp(inv), len(Ms), int(bool(to_host)),
                                                   C.byref(used)))
        return used.value

    def batch_frames_of(self, which, copy=None):
        """[(output, output_alpha), ...] of a set render_batch(to_host=True) returned: waits for its copies; views of
        the set's pinned planes (valid until the set is rendered into again, two render_batch calls later) or copies"""
        host, n_frames = _lib._FP(), C.c_int()
        self._check(self._lib.spv_batch_wait(self._ctx, int(which), C.byref(host), None, C.byref(n_frames)))
        n = self.width * self.height
        flat = self._pinned_view(host, n_frames.value * 2 * n)
        if copy is None:
            copy = not self.pinned_outputs
        if copy:
            flat = flat.copy()
        shape = (self.height, self.width)
        return [(flat[2 * f * n:(2 * f + 1) * n].reshape(shape), flat[(2 * f + 1) * n:(2 * f + 2) * n].reshape(shape))
                for f in range(n_frames.value)]

    # the first launch of a sequence takes half a batch: the device starts after half the host preparation (a batch's
    # matrices are inverted one by one, ~10 us each), and a sequence whose length is a multiple of the batch ends on a
    # half launch, whose last read-back band is half as long -- fill and drain of short sequences (configs[1], 20 frames:
    # 8282 -> 8606 frames/s end to end, three runs each); 5 frames per launch cost 103 instead of 98 us per frame
    # (profiles/r02_exp_axis_v4.txt), long sequences are unchanged
    first_batch_half = True

    def _render_sequence_batched(self, modelViews, p, batch):
        import itertools
        batch = max(1, min(int(batch), _lib.MAX_BATCH))
        first = max(2, (batch + 1) // 2) if (self.first_batch_half and batch >= 4) else batch
        it = iter(modelViews)
        pending = []  # (set, modelViews) in flight, oldest first
        shape = (self.height, self.width)
        last = None

        def frames_of(entry):
            which, Ms = entry
            for M, (o, a) in zip(Ms, self.batch_frames_of(which)):
                self.modelView = 1. * M
                self.output, self.output_alpha = o, a
                if self.output_depth is None or self.output_depth.shape != shape:
                    self.output_depth = np.zeros(shape, np.float32)
                yield self
        try:
            while True:
                Ms = [np.asarray(M, dtype=np.float64) for M in itertools.islice(it, first)]
                first = batch
                if not Ms:
                    break
                last = Ms[-1]
                pending.append((self.render_batch(Ms, True), Ms))
                if len(pending) == 2:
                    for _ in frames_of(pending.pop(0)):
                        yield self
            while pending:
                for _ in frames_of(pending.pop(0)):
                    yield self
        finally:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:  # (not closed under an abandoned generator)
                self._lib.spv_sync(self._ctx)
                if last is not None:
                    self.set_modelView(last)  # the context's own matrices follow the sequence, as with one launch per frame

    def set_view_copies(self, mode="auto"):
        """Which layered copies of the volume max projections sample (csrc/spv_mip_axis.cu): "auto" (default)
        = per frame the copy with pairs along x, y or z under which a texture request stays inside one layer -- the x / y
        copies cost 4 bytes per voxel each and are built on the device when a frame first wants them after an upload;
        (8 for float32 volumes, for which the z copy is an extra one as well);
        "primary" = the z copy only (time points that are uploaded, rendered once and replaced; float32: mip_fast_kernel
        on the 3-D array); "off" = mip_fast_kernel, as round 1."""
        codes = {"off": 0, "auto": 1, "primary": 2}
        if mode not in codes:
            raise KeyError("view copies = '%s' not defined, valid: %s" % (mode, sorted(codes)))
        self._check(self._lib.spv_set_tuning(self._ctx, 16, codes[mode]))

    def mip_axis_used(self):
        """(layer axis, lane map) of the last plain max projection: axis 0 x / 1 y / 2 z, lane map 0 = 2x2-pixel quads /
        1 = row quads / 2 = column quads; (-1, -1): mip_fast_kernel"""
        a, q = C.c_int(), C.c_int()
        self._check(self._lib.spv_mip_axis_used(self._ctx, C.byref(a), C.byref(q)))
        return a.value, q.value

    def _adopt_slot(self, slot, planes, clear_extras=False):
        host = _lib._FP()
        self._check(self._lib.spv_wait_slot(self._ctx, slot, C.byref(host)))
        n = self.width * self.height
        flat = self._pinned_view(host, planes * n)
        if not self.pinned_outputs:
            flat = flat.copy()
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)
        if clear_extras:
            self.output_depth = self.output_occlusion = self.output_normals = None
        if planes == 7:
            self.output_depth = flat[2 * n:3 * n].reshape(shape)
            self.output_occlusion = flat[3 * n:4 * n].reshape(shape)
            self.output_normals = flat[4 * n:7 * n].reshape(shape + (3,))

    # ------------------------------------------------------------------ diagnostics (additions)
    def last_render_ms(self):
        ms = C.c_float()
        self._check(self._lib.spv_last_timing_ms(self._ctx, C.byref(ms)))
        return ms.value

    def enable_stats(self, on=True):
        self._check(self._lib.spv_enable_stats(self._ctx, int(bool(on))))

    def last_stats(self):
        """(hit rays, texture samples issued) of the last render; needs enable_stats()."""
        v = (C.c_ulonglong * 2)()
        self._check(self._lib.spv_last_stats(self._ctx, v, 2))
        return int(v[0]), int(v[1])

    def last_warp_cycles(self):
        """(longest warp, sum over warps) of the last iso-surface search in SM cycles; needs enable_stats()."""
        v = (C.c_ulonglong * 40)()
        self._check(self._lib.spv_last_stats(self._ctx, v, 40))
        self.last_warp_histogram = [int(x) for x in v[4:40]]  # warps by floor(log2(cycles))
        return int(v[2]), int(v[3])

    def texrate_probe(self, iters=2000, footprint=None):
        """Measured samples/s of independent cache-resident filtered fetches (roofline denominator).  footprint=None:
        neighbouring lanes 0.6 texel apart (the unit's peak); otherwise (a, b, m), three vectors in texels (x, y, z):
        lane (lx, ly) of a warp's 8x4 tile fetches at base + lx*a + ly*b + j*m -- the attainable rate for the ray
        spacing (a, b) and sample spacing (m) of a camera."""
        v = C.c_double()
        if footprint is None:
            self._check(self._lib.spv_texrate_probe(self._ctx, int(iters), C.byref(v)))
        else:
            vec = np.ascontiguousarray(np.asarray(footprint, np.float32).reshape(9))
            self._check(self._lib.spv_texrate_probe_footprint(self._ctx, int(iters), _lib.fp(vec), C.byref(v)))
        return v.value

    def sample_points(self, pos):
        """Values of the resident volume at normalised positions pos (n, 3) = (x, y, z) in [0, 1], through the
        current sampler: what read_imagef(volume, sampler, pos).x returns to the reference kernels."""
        pos = np.ascontiguousarray(pos, np.float32)
        out = np.empty(len(pos), np.float32)
        self._check(self._lib.spv_sample_points(self._ctx, _lib.fp(pos), len(pos), _lib.fp(out)))
        return out

    def use_stream(self, cuda_stream=None):
        """Run on a caller-owned CUDA stream (an int cudaStream_t, e.g. torch's current stream); None restores
        the renderer's own stream."""
        self._check(self._lib.spv_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    def render_device_only(self, numParts=1, currentPart=0):
        """Enqueue one max projection with the current settings and return without reading anything back."""
        p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma),
                           float(self.alphaPow), int(numParts), int(currentPart), int(self.max_steps), 0)
        self._check(self._lib.spv_render_mip(self._ctx, C.byref(p)))

    def sync(self):
        self._check(self._lib.spv_sync(self._ctx))

    # ------------------------------------------------------------------ display hand-off (addition)
    def set_lut(self, rgb):
        """Colour map of output_rgba(): (N, 3) or (N, 4) floats in [0, 1] -- what GLWidget.set_colormap uploads as
        texture_LUT (gui/glwidget.py, arrayFromImage of the colormap png)."""
        rgb = np.ascontiguousarray(np.asarray(rgb, dtype=np.float32)[:, :3])
        self._check(self._lib.spv_set_lut(self._ctx, _lib.fp(rgb.reshape(-1)), int(rgb.shape[0])))

    def output_rgba(self, mode_black=True):
        """(height, width, 4) uint8: the current result shaded the way the reference's fragment shader does
        (gui/shaders/texture.frag:8-38: LUT colour, alpha = value, transparent where output_alpha < 0), computed on
        the device and read back as 4 bytes per pixel instead of the two float planes.  Use after render() or
        render_device_only()."""
        img = np.empty((self.height, self.width, 4), np.uint8)
        self._check(self._lib.spv_read_rgba8(self._ctx, int(bool(mode_black)), img.ctypes.data, img.nbytes))
        return img

    def launch_count(self):
        n = C.c_ulonglong()
        self._check(self._lib.spv_launch_count(self._ctx, C.byref(n)))
        return int(n.value)

    def d2h_bytes(self):
        """result bytes copied device -> host by this renderer so far (rows that cannot hold a hit are not copied)"""
        n = C.c_ulonglong()
        self._check(self._lib.spv_d2h_bytes(self._ctx, C.byref(n)))
        return int(n.value)
