"""ctypes binding of libspimcuda.so (include/spimcuda.h).

There is no CPU implementation behind this module: if the shared library is missing or no CUDA
device is usable, loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SPIMCUDA_LIB: another build of the same library (kernel-tuning experiments); never a different implementation
LIB_PATH = os.environ.get("SPIMCUDA_LIB") or os.path.join(HERE, "libspimcuda.so")

SPV_F32, SPV_U16, SPV_U8 = 0, 1, 2
BUF_OUT, BUF_ALPHA, BUF_DEPTH, BUF_NORMALS, BUF_OCC, BUF_RAW, BUF_KPLANES = range(7)
SAMPLER_TMU, SAMPLER_EXACT = 0, 1
MIP_RAW_ONLY = 1
MAX_BATCH = 16
ISO_RAW_ONLY = 1

DTYPE_CODES = {np.dtype(np.float32): SPV_F32, np.dtype(np.uint16): SPV_U16, np.dtype(np.uint8): SPV_U8}
# host element types the ingest path converts on the device (SPV_SRC_*)
SRC_CODES = {np.dtype(t): i for i, t in enumerate([np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32,
                                                   np.int64, np.uint64, np.float16, np.float32, np.float64, np.bool_])}


class MipParams(C.Structure):
    _fields_ = [("box", C.c_float * 6), ("min_val", C.c_float), ("max_val", C.c_float), ("gamma", C.c_float),
                ("alpha_pow", C.c_float), ("num_parts", C.c_int), ("current_part", C.c_int),
                ("max_steps", C.c_int), ("flags", C.c_int)]


class IsoParams(C.Structure):
    _fields_ = [("box", C.c_float * 6), ("iso_val", C.c_float), ("gamma", C.c_float), ("max_steps", C.c_int),
                ("occ_strength", C.c_float), ("occ_radius", C.c_int), ("occ_n_points", C.c_int),
                ("flags", C.c_int)]


_FP = C.POINTER(C.c_float)
_CTX = C.c_void_p

# name -> (restype, argtypes); every symbol include/spimcuda.h declares
SIGNATURES = {
    "spv_version": (C.c_int, []),
    "spv_last_error": (C.c_char_p, [_CTX]),
    "spv_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_CTX)]),
    "spv_destroy": (C.c_int, [_CTX]),
    "spv_resize": (C.c_int, [_CTX, C.c_int, C.c_int]),
    "spv_set_stream": (C.c_int, [_CTX, C.c_void_p]),
    "spv_share_stream": (C.c_int, [_CTX, _CTX]),
    "spv_sync": (C.c_int, [_CTX]),
    "spv_set_volume": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spv_update_volume": (C.c_int, [_CTX, C.c_void_p]),
    "spv_update_volume_async": (C.c_int, [_CTX, C.c_void_p]),
    "spv_set_volume_from": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spv_update_volume_from": (C.c_int, [_CTX, C.c_void_p, C.c_int]),
    "spv_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "spv_host_free": (C.c_int, [C.c_void_p]),
    "spv_set_volume_device": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spv_set_volume_slab": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int]),
    "spv_set_volume_slab_halo": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int]),
    "spv_volume_minmax": (C.c_int, [_CTX, _FP, _FP]),
    "spv_set_interp": (C.c_int, [_CTX, C.c_int]),
    "spv_set_sampler": (C.c_int, [_CTX, C.c_int]),
    "spv_set_int_filter": (C.c_int, [_CTX, C.c_int]),
    "spv_set_layout": (C.c_int, [_CTX, C.c_int]),
    "spv_set_skipping": (C.c_int, [_CTX, C.c_int]),
    "spv_set_matrices": (C.c_int, [_CTX, _FP, _FP]),
    "spv_render_mip": (C.c_int, [_CTX, C.POINTER(MipParams)]),
    "spv_render_mip_to_host": (C.c_int, [_CTX, C.POINTER(MipParams), C.c_int, C.c_int, C.POINTER(_FP)]),
    "spv_mip_finish": (C.c_int, [_CTX, C.POINTER(MipParams)]),
    "spv_comp_init": (C.c_int, [_CTX, C.c_int, C.c_int]),
    "spv_comp_export": (C.c_int, [_CTX, C.c_void_p, C.c_size_t]),
    "spv_comp_import": (C.c_int, [_CTX, C.c_int, C.c_void_p, C.c_size_t]),
    "spv_comp_import_local": (C.c_int, [_CTX, C.c_int, _CTX]),
    "spv_render_mip_composite": (C.c_int, [_CTX, C.POINTER(MipParams)]),
    "spv_comp_check": (C.c_int, [_CTX]),
    "spv_set_extra_slabs": (C.c_int, [_CTX, C.POINTER(_CTX), C.c_int]),
    "spv_render_iso": (C.c_int, [_CTX, C.POINTER(IsoParams)]),
    "spv_render_iso_to_host": (C.c_int, [_CTX, C.POINTER(IsoParams), C.c_int, C.POINTER(_FP)]),
    "spv_iso_slab_search": (C.c_int, [_CTX, C.POINTER(IsoParams)]),
    "spv_iso_slab_resolve": (C.c_int, [_CTX, C.POINTER(IsoParams)]),
    "spv_iso_slab_post": (C.c_int, [_CTX, C.POINTER(IsoParams)]),
    "spv_iso_slab_check": (C.c_int, [_CTX]),
    "spv_render_iso_composite": (C.c_int, [_CTX, C.POINTER(IsoParams)]),
    "spv_read": (C.c_int, [_CTX, C.c_int, _FP, C.c_size_t]),
    "spv_read_many": (C.c_int, [_CTX, _FP, _FP, _FP, _FP, _FP]),
    "spv_read_pinned": (C.c_int, [_CTX, C.c_int, C.POINTER(_FP)]),
    "spv_device_ptr": (C.c_int, [_CTX, C.c_int, C.POINTER(C.c_void_p)]),
    "spv_set_lut": (C.c_int, [_CTX, _FP, C.c_int]),
    "spv_read_rgba8": (C.c_int, [_CTX, C.c_int, C.c_void_p, C.c_size_t]),
    "spv_select_slot": (C.c_int, [_CTX, C.c_int]),
    "spv_read_pinned_async": (C.c_int, [_CTX, C.c_int]),
    "spv_wait_slot": (C.c_int, [_CTX, C.c_int, C.POINTER(_FP)]),
    "spv_last_timing_ms": (C.c_int, [_CTX, _FP]),
    "spv_last_stats": (C.c_int, [_CTX, C.POINTER(C.c_ulonglong), C.c_int]),
    "spv_enable_stats": (C.c_int, [_CTX, C.c_int]),
    "spv_set_tuning": (C.c_int, [_CTX, C.c_int, C.c_int]),
    "spv_sample_points": (C.c_int, [_CTX, _FP, C.c_int, _FP]),
    "spv_texrate_probe": (C.c_int, [_CTX, C.c_int, C.POINTER(C.c_double)]),
    "spv_texrate_probe_footprint": (C.c_int, [_CTX, C.c_int, _FP, C.POINTER(C.c_double)]),
    "spv_launch_count": (C.c_int, [_CTX, C.POINTER(C.c_ulonglong)]),
    "spv_d2h_bytes": (C.c_int, [_CTX, C.POINTER(C.c_ulonglong)]),
    "spv_last_phases_ms": (C.c_int, [_CTX, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "spv_stream_join": (C.c_int, [_CTX]),
    "spv_set_mip_path": (C.c_int, [_CTX, C.c_int]),
    "spv_mip_path_used": (C.c_int, [_CTX, C.POINTER(C.c_int)]),
    "spv_render_mip_batch": (C.c_int, [_CTX, C.POINTER(MipParams), _FP, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "spv_mip_batch_possible": (C.c_int, [_CTX, C.POINTER(MipParams)]),
    "spv_batch_wait": (C.c_int, [_CTX, C.c_int, C.POINTER(_FP), C.POINTER(_FP), C.POINTER(C.c_int)]),
    "spv_mip_axis_used": (C.c_int, [_CTX, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spv_update_volume_device_from": (C.c_int, [_CTX, C.c_void_p, C.c_int]),
    "spv_filter_create": (C.c_int, [C.c_int, C.POINTER(_CTX)]),
    "spv_filter_destroy": (C.c_int, [_CTX]),
    "spv_filter_load": (C.c_int, [_CTX, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spv_filter_convolve_sep3": (C.c_int, [_CTX, _FP, C.c_int, _FP, C.c_int, _FP, C.c_int]),
    "spv_filter_sync": (C.c_int, [_CTX]),
    "spv_filter_result_device": (C.c_int, [_CTX, C.POINTER(_FP)]),
    "spv_filter_read": (C.c_int, [_CTX, _FP, C.c_size_t]),
    "spv_filter_last_ms": (C.c_int, [_CTX, _FP]),
    "spv_filter_last_pass_ms": (C.c_int, [_CTX, _FP, C.POINTER(C.c_int)]),
    "spv_filter_last_error": (C.c_char_p, [_CTX]),
    "spv_filter_set_tuning": (C.c_int, [_CTX, C.c_int, C.c_int]),
    "spv_filter_launch_count": (C.c_int, [_CTX, C.POINTER(C.c_ulonglong)]),
}

_lib = None


def load():
    """Load libspimcuda.so (once).  Raises ImportError with a build hint if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -m spimagine_b200.build` "
                          "(needs nvcc; there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_FP)


def pinned_empty(shape, dtype):
    """ndarray in page-locked host memory (spv_host_alloc): the source of asynchronous uploads at PCIe rate
    (VolumeRenderer.update_data(..., pinned=True), TimelapsePlayer) -- freed when the array is collected."""
    import weakref
    lib = load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.spv_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.spv_host_free, C.c_void_p(p.value))
    return arr


class SpvError(RuntimeError):
    pass


def check(rc, ctx=None):
    if rc != 0:
        msg = load().spv_last_error(ctx)
        raise SpvError("libspimcuda: %s" % (msg.decode() if msg else "error %d" % rc))
