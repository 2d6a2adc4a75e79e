"""Generates tests/golden/processors_ref.json by running the REFERENCE's own image processors
(/root/reference/spimagine/models/imageprocessor.py, loaded by path) with `gputools` replaced by a recorder:
  * BlurProcessor / BlurXYZProcessor: the taps they hand to gputools.convolve_sep3 (float64, exact);
  * FFTProcessor.apply: the reference's expression executed with gputools.pad_to_power2 / pad_to_shape / fft standing on
    the restated helpers of oracle/filters.py (gputools itself is not available: those three stay unpinned);
  * CopyProcessor / LucyRichProcessor / FuncProcessor: identity / call-through behaviour, names and kwargs access.

    python tests/golden/make_processor_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/spimagine/models/imageprocessor.py"


def main():
    from oracle import filters as forc
    calls = []
    g = types.ModuleType("gputools")
    g.convolve_sep3 = lambda data, hx, hy, hz: (calls.append([np.asarray(h, np.float64).tolist() for h in (hx, hy, hz)]), data)[1]
    g.pad_to_power2 = lambda data, mode="constant": forc.pad_to_power2(data, mode)
    g.pad_to_shape = lambda d, shape, mode="constant": forc.pad_to_shape(d, shape, mode)
    g.fft = lambda d: np.fft.fftn(d)
    sys.modules["gputools"] = g
    spec = importlib.util.spec_from_file_location("ref_imageprocessor", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    vol = np.zeros((2, 2, 2), np.float32)
    taps = {}
    for sigma in (4., 1., 2., 2.5, .5, 7.):
        del calls[:]
        ref.BlurProcessor(sigma=sigma).apply(vol)
        taps["blur_%g" % sigma] = calls[0]
    del calls[:]
    ref.BlurXYZProcessor(sx=1., sy=2., sz=3.).apply(vol)
    taps["blur_xyz_1_2_3"] = calls[0]

    rng = np.random.default_rng(5)
    data = rng.integers(0, 4000, size=(5, 6, 7)).astype(np.uint16)
    fft = {"data": data.tolist(), "dtype": "uint16"}
    for log in (False, True):
        out = ref.FFTProcessor(log=log).apply(data)
        fft["log_%s" % log] = {"dtype": str(out.dtype), "value": np.asarray(out, np.float64).tolist()}

    iface = {}
    for name, proc in (("copy", ref.CopyProcessor()), ("blur", ref.BlurProcessor()), ("blur_xyz", ref.BlurXYZProcessor()),
                       ("noise", ref.NoiseProcessor()), ("fft", ref.FFTProcessor()), ("lucy", ref.LucyRichProcessor()),
                       ("func", ref.FuncProcessor(lambda d, k=2: d * k, "times", k=3))):
        iface[name] = {"name": proc.name, "kwargs": {k: v for k, v in proc.kwargs.items()}}
    ident = {"copy": bool(ref.CopyProcessor().apply(vol) is vol), "lucy": bool(ref.LucyRichProcessor().apply(vol) is vol),
             "func": float(ref.FuncProcessor(lambda d, k=2: d * k, "times", k=3).apply(np.ones(1))[0])}
    with open(os.path.join(HERE, "processors_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_processor_golden.py", "taps": taps, "fft": fft, "interface": iface,
                   "identity": ident}, f)
    print("taps:", sorted(taps), " fft dtypes:", fft["log_False"]["dtype"], fft["log_True"]["dtype"], " iface:", iface)


if __name__ == "__main__":
    main()
