"""Generates tests/golden/tiff_ref.json: the files of tiff_inputs.py read by the REFERENCE's vendored tifffile
(/root/reference/spimagine/lib/tifffile.py `imread`, what imgutils.read3dTiff and TiffData call), entered through a
bare `spimagine.lib` namespace.  np.fromstring's binary mode (removed from numpy, used by its strip decoders) is
restored for the run.  Recorded per file: shape, dtype and SHA-1 of the array in native byte order.

    python tests/golden/make_tiff_golden.py
"""
import hashlib
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"


def main():
    import tiff_inputs
    for name in ("spimagine", "spimagine.lib"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    warnings.simplefilter("ignore")
    old = np.fromstring
    np.fromstring = lambda s, dtype=float, count=-1, sep="": (np.frombuffer(s, dtype, count).copy() if sep == ""
                                                               else old(s, dtype, count, sep=sep))
    import spimagine.lib.tifffile as tf
    out = {}
    with tempfile.TemporaryDirectory() as root:
        for name, fn in sorted(tiff_inputs.build(root).items()):
            a = tf.imread(fn)
            a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("=")))
            out[name] = {"shape": list(a.shape), "dtype": a.dtype.name, "sha1": hashlib.sha1(a.tobytes()).hexdigest()}
            print(name, a.shape, a.dtype)
    with open(os.path.join(HERE, "tiff_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_tiff_golden.py (reference lib/tifffile.py imread)", "files": out}, f, indent=1)


if __name__ == "__main__":
    main()
