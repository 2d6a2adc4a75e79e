#!/usr/bin/env python
"""Host-side throughput of the frame readers on a 64 x 512 x 512 uint16 stack (page cache warm): TiffFile.read_into
for stored / LZW / LZW + predictor / deflate / PackBits strips against PIL (libtiff) on the same file, and the CZI
reader.  No GPU involved.    python scripts/bench_readers.py > profiles/rNN_exp_readers.txt"""
import os
import sys
import tempfile
import time

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from spimagine_b200.utils import cziio, tiffio  # noqa: E402
import czi_inputs  # noqa: E402


def best(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    rng = np.random.default_rng(0)
    z, y, x = np.mgrid[:64, :512, :512]
    a = (3000 * np.exp(-((z - 32) ** 2 / 300. + (y - 256) ** 2 / 20000. + (x - 256) ** 2 / 20000.)) +
         rng.integers(0, 20, z.shape)).astype(np.uint16)
    mb = a.nbytes / 1e6
    print("frame readers, %d x %d x %d uint16 (%.0f MB), %d cores visible, best of 3, page cache warm" % (a.shape + (mb, os.cpu_count())))
    print("%-28s %10s %14s %14s %12s" % ("file", "size/raw", "ours 1 thread", "ours threads", "PIL/libtiff"))
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "a.tif")
        out = np.empty(a.shape, np.uint16)
        for label, comp, kw in (("tiff stored", None, {}), ("tiff lzw", "tiff_lzw", {}),
                                ("tiff lzw + predictor", "tiff_lzw", {"tiffinfo": {317: 2}}),
                                ("tiff deflate", "tiff_adobe_deflate", {}),
                                ("tiff deflate + predictor", "tiff_adobe_deflate", {"tiffinfo": {317: 2}}),
                                ("tiff packbits", "packbits", {})):
            pages = [Image.fromarray(p) for p in a]
            if comp:
                pages[0].save(fn, compression=comp, save_all=True, append_images=pages[1:], **kw)
            else:
                tiffio.write3dTiff(a, fn)
            t = tiffio.TiffFile(fn)
            t.decode_threads = 1
            one = best(lambda: t.read_into(out))
            assert np.array_equal(out, a)
            t.decode_threads = 0
            many = best(lambda: t.read_into(out))

            def pil():
                im = Image.open(fn)
                for i in range(im.n_frames):
                    im.seek(i)
                    np.array(im)
            ref = best(pil)
            print("%-28s %10.2f %9.0f MB/s %9.0f MB/s %7.0f MB/s" % (label, os.path.getsize(fn) / a.nbytes, mb / one, mb / many, mb / ref))
        fn = os.path.join(d, "a.czi")
        czi_inputs.write_czi(fn, a, "ZYX", "YX")
        c = cziio.CziFile(fn)
        sec = best(lambda: c.read_into(out))
        assert np.array_equal(out, a)
        print("%-28s %10.2f %9.0f MB/s" % ("czi stored, plane sub-blocks", os.path.getsize(fn) / a.nbytes, mb / sec))


if __name__ == "__main__":
    main()
