"""Deterministic TIFF files for tests/golden/make_tiff_golden.py and tests/test_tiff_codecs.py: stacks written by this
package's writer and the hand-laid strip / tile files of tests/test_tiff_codecs.py."""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(root):
    """-> {name: path} of the files written under root"""
    sys.path.insert(0, TESTS)
    try:
        import test_tiff_codecs as T
    finally:
        sys.path.remove(TESTS)
    from spimagine_b200.utils import tiffio
    rng = np.random.default_rng(33)
    out = {}

    def path(name):
        out[name] = os.path.join(root, name + ".tif")
        return out[name]

    for dt in ("uint8", "uint16", "int16", "float32"):
        a = rng.normal(500, 200, (4, 9, 11)).astype(dt)
        tiffio.write3dTiff(a, path("written_3d_" + dt))
    tiffio.write3dTiff(rng.integers(0, 60000, (2, 3, 9, 11)).astype(np.uint16), path("written_4d_uint16"))
    tiffio.write3dTiff(rng.integers(0, 60000, (9, 11)).astype(np.uint16), path("written_2d_uint16"))
    tiffio.write3dTiff(rng.integers(0, 60000, (3, 9, 11)).astype(np.uint16), path("written_bigtiff"), bigtiff=True)
    s = T._smooth((23, 41), np.uint16, seed=4)
    for bo, tag in (("<", "le"), (">", "be")):
        for comp, pred in ((5, 1), (5, 2), (8, 1), (8, 2)):
            T._tiff_with_strips(path("strips_%s_c%d_p%d" % (tag, comp, pred)), s, comp, pred, bo=bo, rows_per_strip=4,
                                pad_last=True)
    pages = list(T._smooth((3, 23, 37), np.uint16, seed=6))
    for comp, pred in ((1, 1), (8, 1), (8, 2), (5, 1), (5, 2)):
        T._tiff_with_tiles(path("tiles_c%d_p%d" % (comp, pred)), pages, (16, 16), comp, pred, "<")
    T._tiff_with_tiles(path("tiles_be_c8_p2"), pages, (32, 16), 8, 2, ">")
    return out
