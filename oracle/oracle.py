"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/spim_oracle.c).  Two interchangeable back ends
behind one ABI:

  kind="port"       oracle/libspim_oracle.so   -- the C restatement
  kind="reference"  oracle/_ref/libspim_ref.so -- the reference's own kernel text
                                                  compiled for the host (build.py)

`OracleRenderer` mirrors the call sequence of the reference's VolumeRenderer
(spimagine/volumerender/volumerender.py:59-547) so a parity test reads like a reference
test: set_data / set_units / set_modelView / set_projection / render(method=...).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DTYPE_CODES = {np.dtype(np.float32): 0, np.dtype(np.uint16): 1, np.dtype(np.uint8): 2}


class SoVolume(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("filter", C.c_int), ("int_linear", C.c_int), ("weight_bits", C.c_int)]


_FP = C.POINTER(C.c_float)


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_FP)


def lib_path(kind):
    if kind == "port":
        return os.path.join(HERE, "libspim_oracle.so")
    if kind == "reference":
        return os.path.join(HERE, "_ref", "libspim_ref.so")
    if kind == "reference_fast":  # timing only: the reference's fast-math build options (oracle/build.py)
        return os.path.join(HERE, "_ref", "libspim_ref_fast.so")
    raise ValueError(kind)


def cpu_has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2 " in f.read().replace("\n", " ")
    except OSError:
        return False


def available(kind):
    return os.path.exists(lib_path(kind))


_LIBS = {}


def load(kind="port"):
    if kind in _LIBS:
        return _LIBS[kind]
    path = lib_path(kind)
    if not os.path.exists(path):
        from . import build
        if kind == "port":
            build.build_oracle()
        else:
            build.build_ref(fast=kind == "reference_fast")
    lib = C.CDLL(path)
    VP = C.POINTER(SoVolume)
    lib.so_max_project.argtypes = [VP, C.c_int, C.c_int, _FP, _FP, _FP, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int, C.c_int, C.c_int, C.c_int, _FP, _FP]
    lib.so_iso_surface.argtypes = [VP, C.c_int, C.c_int, _FP, _FP, _FP, C.c_float, C.c_float, C.c_int,
                                   _FP, _FP, _FP, _FP]
    lib.so_convolve_scalar.argtypes = [_FP, _FP, C.c_int, C.c_int, C.c_int]
    lib.so_convolve_vec.argtypes = [_FP, _FP, C.c_int, C.c_int, C.c_int]
    lib.so_occlusion.argtypes = [_FP, C.c_int, C.c_int, C.c_int, C.c_int, _FP]
    lib.so_shading.argtypes = [_FP, C.c_int, C.c_int, _FP, _FP, C.c_float, _FP, _FP, _FP]
    lib.so_render_isosurface.argtypes = [VP, C.c_int, C.c_int, _FP, _FP, _FP, C.c_float, C.c_float, C.c_int,
                                         C.c_float, C.c_int, C.c_int, _FP, _FP, _FP, _FP, _FP, _FP, _FP]
    if kind == "port":
        lib.so_max_project_raw.argtypes = [VP, C.c_int, C.c_int, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP]
    lib.so_set_row_sampling.argtypes = [C.c_int, C.c_int]
    lib.so_set_row_sampling.restype = None
    lib.so_count_hit_rays.argtypes = [C.c_int, C.c_int, _FP, _FP, _FP]
    lib.so_count_hit_rays.restype = C.c_long
    lib.so_random.argtypes = [C.c_uint32, C.c_uint32]
    lib.so_random.restype = C.c_float
    lib.so_rand_int.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    lib.so_rand_int.restype = C.c_float
    lib.so_sample.argtypes = [VP, C.c_float, C.c_float, C.c_float]
    lib.so_sample.restype = C.c_float
    lib.so_num_threads.restype = C.c_int
    lib.so_kind.restype = C.c_char_p
    if kind == "port":
        lib.so_lcg_hash.argtypes = [C.c_uint32, C.c_uint32]
        lib.so_lcg_hash.restype = C.c_uint32
    _LIBS[kind] = lib
    return lib


def make_volume(data, filter_linear=True, int_linear=True, weight_bits=0):
    """data: C-contiguous ndarray (z,y,x) of float32 / uint16 / uint8.  Keeps `data` alive via the return."""
    data = np.ascontiguousarray(data)
    nz, ny, nx = data.shape
    v = SoVolume(data.ctypes.data, DTYPE_CODES[data.dtype], nx, ny, nz, int(bool(filter_linear)),
                 int(bool(int_linear)), int(weight_bits))
    v._keep = data
    return v


# ---- host glue: spimagine/volumerender/volumerender.py:310-325 --------------------------------
def stack_scale_mat(shape_xyz, units):
    Nx, Ny, Nz = shape_xyz
    dx, dy, dz = units
    maxDim = max(d * N for d, N in zip([dx, dy, dz], [Nx, Ny, Nz]))
    return np.diag([1. * dx * Nx / maxDim, 1. * dy * Ny / maxDim, 1. * dz * Nz / maxDim, 1.]).astype(np.float32)


def inverse_matrices(modelView, projection, shape_xyz, units=(1., 1., 1.)):
    """-> (invP, invM) as flat float32[16], row-major, like update_matrices()."""
    from scipy.linalg import inv
    invM = inv(np.dot(modelView, stack_scale_mat(shape_xyz, units)))
    invP = inv(projection)
    return invP.flatten().astype(np.float32), invM.flatten().astype(np.float32)


def perspective(fovy=45, aspect=1., z1=0.1, z2=10):
    """spimagine/utils/transform_matrices.py:45-54 (the renderer's default projection)."""
    f = 1. / np.tan(fovy / 180. * np.pi / 2.)
    return np.array([[1. * f / aspect, 0, 0, 0], [0, f, 0, 0],
                     [0, 0, -1. * (z2 + z1) / (z2 - z1), -2. * z1 * z2 / (z2 - z1)], [0, 0, -1, 0]], np.float32)


class OracleRenderer(object):
    dtypes = [np.float32, np.uint16, np.uint8]

    def __init__(self, size=None, interpolation="linear", kind="port", max_steps=200, int_linear=True,
                 weight_bits=0, pos_mode=0):
        if interpolation not in ("linear", "nearest"):
            raise KeyError(interpolation)
        self.lib = load(kind)
        self.kind = kind
        self.interpolation = interpolation
        self.max_steps = max_steps
        self.int_linear = int_linear
        self.weight_bits = weight_bits
        self.pos_mode = pos_mode
        self.width, self.height = size if size else (200, 200)
        self.gamma, self.maxVal, self.minVal, self.alphaPow = 1., 0., 0., 0.
        self.occ_strength, self.occ_radius, self.occ_n_points = .1, 21, 30
        self.boxBounds = np.array([-1, 1, -1, 1, -1, 1], np.float32)
        self.stackUnits = np.ones(3)
        self.modelView = np.identity(4)
        self.projection = perspective()
        self._alloc()

    def _alloc(self):
        h, w = self.height, self.width
        self.output = np.zeros((h, w), np.float32)
        self.output_alpha = np.zeros((h, w), np.float32)
        self.output_depth = np.zeros((h, w), np.float32)
        self.output_normals = np.zeros((h, w, 3), np.float32)
        self.output_occlusion = np.zeros((h, w), np.float32)
        self._tmp = np.zeros((h, w), np.float32)
        self._tmp_vec = np.zeros((h, w, 3), np.float32)

    def set_data(self, data):
        if data.dtype.type not in self.dtypes:
            data = data.astype(np.float32)
        self.vol = make_volume(data, self.interpolation == "linear", self.int_linear, self.weight_bits)

    def set_units(self, u): self.stackUnits = np.array(u)
    def set_modelView(self, m): self.modelView = 1. * np.asarray(m)
    def set_projection(self, p): self.projection = np.asarray(p)
    def set_box_boundaries(self, b): self.boxBounds = np.array(b, np.float32)
    def set_max_val(self, v): self.maxVal = v
    def set_min_val(self, v): self.minVal = v
    def set_gamma(self, g): self.gamma = g
    def set_alpha_pow(self, a): self.alphaPow = a
    def set_occ_strength(self, o): self.occ_strength = o
    def set_occ_radius(self, r): self.occ_radius = r
    def set_occ_n_points(self, n): self.occ_n_points = n

    def matrices(self):
        v = self.vol
        return inverse_matrices(self.modelView, self.projection, (v.nx, v.ny, v.nz), self.stackUnits)

    def render(self, data=None, maxVal=None, minVal=None, gamma=None, modelView=None, projection=None,
               method="max_project", numParts=1, currentPart=0):
        if data is not None: self.set_data(data)
        if maxVal is not None: self.maxVal = maxVal
        if minVal is not None: self.minVal = minVal
        if gamma is not None: self.gamma = gamma
        if modelView is not None: self.set_modelView(modelView)
        if projection is not None: self.set_projection(projection)
        invP, invM = self.matrices()
        box = np.ascontiguousarray(self.boxBounds, np.float32)
        V = C.byref(self.vol)
        if method == "max_project":
            rc = self.lib.so_max_project(V, self.width, self.height, _fp(invP), _fp(invM), _fp(box),
                                         self.minVal, self.maxVal, self.gamma, self.alphaPow, numParts, currentPart,
                                         self.max_steps, self.pos_mode, _fp(self.output), _fp(self.output_alpha))
        elif method == "iso_surface":
            rc = self.lib.so_render_isosurface(V, self.width, self.height, _fp(invP), _fp(invM), _fp(box),
                                               self.maxVal, self.gamma, self.max_steps, self.occ_strength,
                                               int(self.occ_radius), int(self.occ_n_points), _fp(self.output),
                                               _fp(self.output_alpha), _fp(self.output_depth),
                                               _fp(self.output_normals), _fp(self.output_occlusion),
                                               _fp(self._tmp), _fp(self._tmp_vec))
        elif method == "iso_surface_raw":  # the iso_surface kernel alone, no post passes
            rc = self.lib.so_iso_surface(V, self.width, self.height, _fp(invP), _fp(invM), _fp(box),
                                         self.maxVal / 2, self.gamma, self.max_steps, _fp(self.output),
                                         _fp(self.output_alpha), _fp(self.output_depth), _fp(self.output_normals))
        else:
            raise ValueError(method)
        if rc != 0:
            raise RuntimeError("oracle returned %d" % rc)

    def render_raw(self, z0=0, z1=None):
        """Sort-last partial: raw ray maximum over the samples owned by slices [z0, z1) (port only)."""
        invP, invM = self.matrices()
        box = np.ascontiguousarray(self.boxBounds, np.float32)
        raw = np.zeros((self.height, self.width), np.float32)
        rc = self.lib.so_max_project_raw(C.byref(self.vol), self.width, self.height, _fp(invP), _fp(invM), _fp(box),
                                         self.max_steps, int(z0), int(self.vol.nz if z1 is None else z1), _fp(raw))
        if rc != 0:
            raise RuntimeError("oracle returned %d" % rc)
        return raw

    def count_hit_rays(self):
        invP, invM = self.matrices()
        box = np.ascontiguousarray(self.boxBounds, np.float32)
        return int(self.lib.so_count_hit_rays(self.width, self.height, _fp(invP), _fp(invM), _fp(box)))


def display_rgba8(value, alpha, lut, mode_black=True):
    """The reference's display path for one frame, restated in numpy fp32 (test infrastructure, like the rest of
    oracle/): value plane uploaded as a GL_RED texture (spimagine/gui/gui_utils.py:146-149: col = (v, 0, 0), unorm
    clamp), LUT look-up and alpha of spimagine/gui/shaders/texture.frag:8-38 with the LUT a 1 x N GL_LINEAR /
    CLAMP_TO_EDGE texture (gui_utils.py:136-145), 8-bit frame buffer conversion rint(255 x)."""
    f32 = np.float32
    v = np.clip(np.nan_to_num(np.asarray(value, f32), nan=0.), f32(0), f32(1)).astype(f32)
    lut = np.asarray(lut, f32)[:, :3]
    N = lut.shape[0]
    s = v if mode_black else (f32(1) - v)
    u = s * f32(N) - f32(0.5)
    fl = np.floor(u)
    f = (u - fl).astype(f32)
    i0 = np.clip(fl.astype(np.int64), 0, N - 1)
    i1 = np.clip(fl.astype(np.int64) + 1, 0, N - 1)
    w0 = (f32(1) - f).astype(f32)
    rgb = (w0[..., None] * lut[i0]).astype(f32) + (f[..., None] * lut[i1]).astype(f32)
    out = np.empty(v.shape + (4,), np.uint8)
    out[..., :3] = np.rint(f32(255) * np.clip(rgb, f32(0), f32(1))).astype(np.uint8)
    out[..., 3] = np.rint(f32(255) * v).astype(np.uint8)
    out[np.asarray(alpha) < 0] = 0
    return out
