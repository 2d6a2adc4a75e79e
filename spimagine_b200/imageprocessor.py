"""Volume processors between the data model and the renderer -- the drop-in for spimagine/models/imageprocessor.py
(ImageProcessor, CopyProcessor, BlurProcessor, BlurXYZProcessor, NoiseProcessor, FFTProcessor, LucyRichProcessor,
FuncProcessor:
same names, constructor arguments, `kwargs` attribute access and `apply(data) -> ndarray`), with the separable blur
running on the B200 through libspimcuda (spv_filter_*) instead of gputools.convolve_sep3.

The reference chain is host -> device -> host per processor and then host -> device again for the renderer
(gui/mainwidget.py:455-465: data = imp.proc.apply(data) ...; renderer.update_data(data)).  Here a chain can stay on
the device: `apply_chain(renderer, data, processors)` uploads the volume once, runs every device-capable processor
in place and hands the float32 result to the renderer's resident array (spv_update_volume_device_from), so neither
the filtered volume nor its re-upload crosses PCIe.

FFTProcessor (the Fourier spectrum of the volume) runs through libspimfft.so (include/spimfft.h): wrap-padding and
element conversion, a real-to-complex cuFFT, and a fused magnitude / fftshift / scale / crop / log pass.
There is no CPU implementation of the blur or the spectrum: without the libraries / a CUDA device, apply() raises.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


class VolumeFilter(object):
    """Device-side separable convolution (owns a stream and two float32 work volumes on `device`)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._f = C.c_void_p()
        rc = self._lib.spv_filter_create(int(device), C.byref(self._f))
        if rc != 0:
            msg = self._lib.spv_filter_last_error(None)
            raise _lib.SpvError("libspimcuda: %s" % (msg.decode() if msg else "error %d" % rc))
        self.shape = None

    def close(self):
        if getattr(self, "_f", None) is not None and self._f.value:
            self._lib.spv_filter_destroy(self._f)
            self._f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.spv_filter_last_error(self._f)
            raise _lib.SpvError("libspimcuda: %s" % (msg.decode() if msg else "error %d" % rc))

    def load(self, data):
        """data: ndarray (Nz, Ny, Nx) of any element type the ingest path knows; others go through float32."""
        data = np.asarray(data)
        if data.ndim != 3:
            raise ValueError("need a 3-D volume, got shape %s" % (data.shape,))
        if np.dtype(data.dtype) not in _lib.SRC_CODES or not data.dtype.isnative:
            data = data.astype(np.float32)
        host = np.ascontiguousarray(data)
        nz, ny, nx = host.shape
        self._check(self._lib.spv_filter_load(self._f, host.ctypes.data, 0, _lib.SRC_CODES[np.dtype(host.dtype)],
                                              nx, ny, nz))
        self.shape = (nz, ny, nx)

    def load_device(self, device_ptr, shape, dtype):
        nz, ny, nx = (int(s) for s in shape)
        self._check(self._lib.spv_filter_load(self._f, C.c_void_p(int(device_ptr)), 1,
                                              _lib.SRC_CODES[np.dtype(dtype)], nx, ny, nz))
        self.shape = (nz, ny, nx)

    def convolve_sep3(self, hx, hy, hz):
        """gputools.convolve_sep3(data, hx, hy, hz) on the loaded volume (or on the previous result)."""
        hs = [np.ascontiguousarray(np.asarray(h, dtype=np.float64).astype(np.float32)) for h in (hx, hy, hz)]
        self._check(self._lib.spv_filter_convolve_sep3(self._f, _lib.fp(hs[0]), len(hs[0]), _lib.fp(hs[1]), len(hs[1]),
                                                       _lib.fp(hs[2]), len(hs[2])))

    def result(self):
        out = np.empty(self.shape, np.float32)
        self._check(self._lib.spv_filter_read(self._f, _lib.fp(out), out.size))
        return out

    def result_device(self):
        """(device pointer, shape) of the float32 result; valid until the next load / convolution."""
        p = _lib._FP()
        self._check(self._lib.spv_filter_result_device(self._f, C.byref(p)))
        return C.cast(p, C.c_void_p).value, self.shape

    def sync(self):
        self._check(self._lib.spv_filter_sync(self._f))

    def set_tuning(self, knob, value):
        """knob 0: x and y pass as one kernel where the tap counts allow it (default 0: three passes);
        knob 1: variant of the y / z passes (1 = automatic, 16 / 32 = one column per thread with that many outputs,
        1602 / 1604 = two / four columns per thread); knob 2: x pass on row pairs, pipelined (2, default) or not (1),
        or on single rows (0).  Results are bit-identical under every knob (include/spimcuda.h)"""
        self._check(self._lib.spv_filter_set_tuning(self._f, int(knob), int(value)))

    def launch_count(self):
        n = C.c_ulonglong()
        self._check(self._lib.spv_filter_launch_count(self._f, C.byref(n)))
        return n.value

    def last_ms(self):
        ms = C.c_float()
        self._check(self._lib.spv_filter_last_ms(self._f, C.byref(ms)))
        return ms.value

    def last_pass_ms(self):
        """device time of each kernel of the last convolution: [x, y, z] (or [fused x + y, z] with tuning knob 0)"""
        ms, n = (C.c_float * 3)(), C.c_int()
        self._check(self._lib.spv_filter_last_pass_ms(self._f, ms, C.byref(n)))
        return [ms[k] for k in range(n.value)]


_shared = {}


def _shared_filter(device=0):
    f = _shared.get(device)
    if f is None:
        f = _shared[device] = VolumeFilter(device)
    return f


def convolve_sep3(data, hx, hy, hz, device=0):
    """Drop-in for gputools.convolve_sep3(data, hx, hy, hz) with a numpy volume: float32 result (Nz, Ny, Nx)."""
    f = _shared_filter(device)
    f.load(data)
    f.convolve_sep3(hx, hy, hz)
    return f.result()


class ImageProcessor(object):
    """models/imageprocessor.py:18-33"""

    def __init__(self, name="", **kwargs):
        self.name = name
        self.set_params(**kwargs)

    def set_params(self, **kwargs):
        self.kwargs = kwargs

    def apply(self, data):
        raise NotImplementedError()

    def __getattr__(self, attr):
        if attr != "kwargs" and attr in self.__dict__.get("kwargs", {}):
            return self.kwargs[attr]
        raise AttributeError(attr)


class CopyProcessor(ImageProcessor):
    """models/imageprocessor.py:37-43"""

    def __init__(self):
        super(CopyProcessor, self).__init__("copy")

    def apply(self, data):
        return data


def _gauss_taps(sigma):
    # models/imageprocessor.py:52-55 / :65-70 -- N = 2 sigma + 1, x = arange(-N, N + 1), h = exp(-x^2 / 2 sigma^2) / sum
    N = 2 * sigma + 1
    x = np.arange(-N, N + 1)
    h = np.exp(-x ** 2 / 2. / sigma ** 2)
    return 1. * h / sum(h)


class BlurProcessor(ImageProcessor):
    """models/imageprocessor.py:46-56"""

    def __init__(self, sigma=4.):
        super(BlurProcessor, self).__init__("blur", sigma=sigma)

    def _taps(self):
        # models/imageprocessor.py:52-55 normalises with `h *= 1. / sum(h)` (BlurXYZProcessor divides, :69): the two
        # differ in the last float64 bit, so each processor follows its own line
        N = 2 * self.sigma + 1
        x = np.arange(-N, N + 1)
        h = np.exp(-x ** 2 / 2. / self.sigma ** 2)
        h *= 1. / sum(h)
        return h, h, h

    def apply(self, data):
        return convolve_sep3(data, *self._taps())

    def apply_device(self, vfilter):
        """addition: blur the volume resident in `vfilter` in place (no host round trip)"""
        vfilter.convolve_sep3(*self._taps())


class BlurXYZProcessor(ImageProcessor):
    """models/imageprocessor.py:59-71 (hx from sx, hy from sy, hz from sz)"""

    def __init__(self, sx=4., sy=4., sz=4.):
        super(BlurXYZProcessor, self).__init__("blur_xyz", sx=sx, sy=sy, sz=sz)

    def _taps(self):
        return tuple(_gauss_taps(s) for s in (self.sx, self.sy, self.sz))

    def apply(self, data):
        return convolve_sep3(data, *self._taps())

    def apply_device(self, vfilter):
        """addition: blur the volume resident in `vfilter` in place (no host round trip)"""
        vfilter.convolve_sep3(*self._taps())


class NoiseProcessor(ImageProcessor):
    """models/imageprocessor.py:73-78 (host-side numpy in the reference as well)"""

    def __init__(self, sigma=10):
        super(NoiseProcessor, self).__init__("noise", sigma=sigma)

    def apply(self, data):
        return np.maximum(0, data + self.sigma * np.random.normal(0, 1, data.shape))


FFT_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libspimfft.so")
_FFT_TYPES = (np.uint8, np.int16, np.uint16, np.float32)
_fft_lib = None


def load_fft():
    """Load libspimfft.so (once) and declare the signatures of include/spimfft.h."""
    global _fft_lib
    if _fft_lib is not None:
        return _fft_lib
    if not os.path.exists(FFT_LIB_PATH):
        raise ImportError("%s not found: build it with `python -m spimagine_b200.build` "
                          "(needs nvcc and cuFFT; there is no CPU fallback)" % FFT_LIB_PATH)
    lib = C.CDLL(FFT_LIB_PATH)
    P, FP = C.c_void_p, C.POINTER(C.c_float)
    sig = {"spf_create": (C.c_int, [C.c_int, C.POINTER(P)]),
           "spf_destroy": (C.c_int, [P]),
           "spf_spectrum": (C.c_int, [P, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP]),
           "spf_result_device": (C.c_int, [P, C.POINTER(FP)]),
           "spf_read": (C.c_int, [P, FP, C.c_size_t]),
           "spf_padded_shape": (C.c_int, [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
           "spf_last_ms": (C.c_int, [P, FP]),
           "spf_launch_count": (C.c_int, [P, C.POINTER(C.c_ulonglong)]),
           "spf_last_error": (C.c_char_p, [P])}
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._signatures = sig
    _fft_lib = lib
    return lib


class SpectrumPlan(object):
    """Device side of FFTProcessor (owns a stream, the padded work arrays and the cuFFT plan on `device`)."""

    def __init__(self, device=0):
        self._lib = load_fft()
        self._p = C.c_void_p()
        rc = self._lib.spf_create(int(device), C.byref(self._p))
        if rc != 0:
            msg = self._lib.spf_last_error(None)
            raise _lib.SpvError("libspimfft: %s" % (msg.decode() if msg else "error %d" % rc))
        self.shape = None

    def close(self):
        if getattr(self, "_p", None) is not None and self._p.value:
            self._lib.spf_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.spf_last_error(self._p)
            raise _lib.SpvError("libspimfft: %s" % (msg.decode() if msg else "error %d" % rc))

    def spectrum(self, data, log=False, read_back=True):
        """data: (Nz, Ny, Nx) ndarray -> float32 spectrum of the same shape (None with read_back=False: the result
        stays on the device, see result_device)."""
        data = np.asarray(data)
        if data.ndim != 3:
            raise ValueError("need a 3-D volume, got shape %s" % (data.shape,))
        if data.dtype.type not in _FFT_TYPES or not data.dtype.isnative:
            data = data.astype(np.float32)  # the reference's astype(complex64) has float32 parts
        host = np.ascontiguousarray(data)
        nz, ny, nx = host.shape
        out = np.empty(host.shape, np.float32) if read_back else None
        self._check(self._lib.spf_spectrum(self._p, host.ctypes.data, 0, _lib.SRC_CODES[np.dtype(host.dtype)],
                                           nx, ny, nz, int(bool(log)), _lib.fp(out) if read_back else None))
        self.shape = host.shape
        return out

    def spectrum_device(self, device_ptr, shape, dtype, log=False):
        nz, ny, nx = (int(v) for v in shape)
        if np.dtype(dtype).type not in _FFT_TYPES:
            raise NotImplementedError("element type %s: uint8, int16, uint16 or float32" % np.dtype(dtype))
        self._check(self._lib.spf_spectrum(self._p, C.c_void_p(int(device_ptr)), 1, _lib.SRC_CODES[np.dtype(dtype)],
                                           nx, ny, nz, int(bool(log)), None))
        self.shape = (nz, ny, nx)

    def result(self):
        out = np.empty(self.shape, np.float32)
        self._check(self._lib.spf_read(self._p, _lib.fp(out), out.size))
        return out

    def result_device(self):
        """-> (device pointer, (Nz, Ny, Nx)) of the float32 result"""
        dev = C.POINTER(C.c_float)()
        self._check(self._lib.spf_result_device(self._p, C.byref(dev)))
        return C.cast(dev, C.c_void_p).value, self.shape

    def padded_shape(self):
        x, y, z = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.spf_padded_shape(self._p, C.byref(x), C.byref(y), C.byref(z)))
        return z.value, y.value, x.value

    def last_ms(self):
        ms = C.c_float()
        self._check(self._lib.spf_last_ms(self._p, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_ulonglong()
        self._check(self._lib.spf_launch_count(self._p, C.byref(n)))
        return int(n.value)


_spectrum_plans = {}


def _shared_plan(device=0):
    if device not in _spectrum_plans:
        _spectrum_plans[device] = SpectrumPlan(device)
    return _spectrum_plans[device]


class FFTProcessor(ImageProcessor):
    """models/imageprocessor.py:82-98: the centred, cropped magnitude spectrum of the wrap-padded volume (log2 of it
    with log=True)."""

    def __init__(self, log=False):
        super(FFTProcessor, self).__init__("fft", log=log)
        self.log = log

    def apply(self, data):
        return _shared_plan().spectrum(data, log=self.log)

    def apply_device(self, plan, device_ptr, shape, dtype):
        plan.spectrum_device(device_ptr, shape, dtype, log=self.log)


class LucyRichProcessor(ImageProcessor):
    """models/imageprocessor.py:101-120: the deconvolution is commented out in the reference; apply returns its input"""

    def __init__(self, rad=4., niter=6):
        super(LucyRichProcessor, self).__init__("RL-Deconv", rad=rad, niter=niter)
        self.rad0 = rad
        self.niter0 = niter
        self.hshape = (1,) * 3

    def reset_psf(self, dshape):
        pass

    def apply(self, data):
        if self.hshape != data.shape or self.rad != self.rad0:
            self.reset_psf(data.shape)
            self.rad0 = self.rad
        return data


class FuncProcessor(ImageProcessor):
    """models/imageprocessor.py:123-130"""

    def __init__(self, func, name="func processor", **kwargs):
        super(FuncProcessor, self).__init__(name, **kwargs)
        self.func = func

    def apply(self, data):
        return self.func(data, **self.kwargs)


def apply_chain(renderer, data, processors, device=None):
    """What MainWidget.impStateChanged does (gui/mainwidget.py:455-465) -- data through every processor, then
    renderer.update_data(result) -- without the round trips: the volume is uploaded once, the blurs and the spectrum
    run on the resident copy (a host-only processor in between gets a host array and its result is uploaded again),
    and the final float32 volume goes from the processor's memory straight into the renderer's resident array,
    converted to the renderer's element type like update_data's astype.  `renderer` must already hold a volume of
    the same shape (set_data).  Returns the device time of the processors in ms."""
    dev = device if device is not None else (renderer.device or 0)
    vf = _shared_filter(dev)
    data = np.asarray(data)
    where, ms = "host", 0.  # "host": `data`; "filter" / "plan": float32 result in that object's device memory
    plan = None

    def device_result():
        if where == "filter":
            ptr, shape = vf.result_device()
            vf.sync()
            return ptr, shape
        return plan.result_device()  # spf_spectrum returns when its stream has finished

    for p in processors:
        if isinstance(p, (BlurProcessor, BlurXYZProcessor)):
            if where == "host":
                vf.load(data)
            elif where == "plan":
                ptr, shape = device_result()
                vf.load_device(ptr, shape, np.float32)
            where = "filter"
            p.apply_device(vf)
            ms += vf.last_ms()
        elif isinstance(p, FFTProcessor):
            plan = _shared_plan(dev) if plan is None else plan
            if where == "host":
                plan.spectrum(data, log=p.log, read_back=False)
            else:
                ptr, shape = device_result()
                p.apply_device(plan, ptr, shape, np.float32)
            where = "plan"
            ms += plan.last_ms()
        elif isinstance(p, (CopyProcessor, LucyRichProcessor)):
            continue  # identity in the reference as well
        else:
            if where != "host":
                ptr, shape = device_result()
                data = plan.result() if where == "plan" else vf.result()
                where = "host"
            data = np.asarray(p.apply(data))
    if where == "host":
        renderer.update_data(data)
        return ms
    ptr, shape = device_result()
    renderer.update_data_device(ptr, shape, np.float32)
    return ms
