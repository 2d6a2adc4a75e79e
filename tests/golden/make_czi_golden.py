"""Generates tests/golden/czi_ref.json: the synthetic CZI files of czi_inputs.py parsed by the REFERENCE's vendored
reader (/root/reference/spimagine/lib/czifile.py, entered through a bare `spimagine.lib` namespace): shape, start,
axes, dtype of the file and, per sub-block, start / shape / SHA-1 of its pixels.  CziFile.asarray() itself indexes
with a list of slices, which numpy >= 1.23 rejects, so the assembled array is built here from the reference's own
per-sub-block results with the same index arithmetic (czifile.py:364-375) and recorded as a SHA-1 after np.squeeze
(imgutils.py:44-47 readCziFile).

    python tests/golden/make_czi_golden.py
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
    import czi_inputs
    for name in ("spimagine", "spimagine.lib"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    warnings.simplefilter("ignore")
    old = np.fromstring      # binary mode removed from numpy; czifile's LZW path uses it
    np.fromstring = lambda s, dtype=float, count=-1, sep="": (np.frombuffer(s, dtype, count).copy() if sep == ""
                                                               else old(s, dtype, count, sep=sep))
    import spimagine.lib.czifile as cz
    out = {}
    with tempfile.TemporaryDirectory() as root:
        for key, (data, axes, block_axes, starts, mosaic) in czi_inputs.cases().items():
            fn = os.path.join(root, key + ".czi")
            czi_inputs.write_case(fn, key)
            with cz.CziFile(fn) as f:
                image = np.zeros(f.shape, f.dtype)
                blocks = []
                for e in f.filtered_subblock_directory:
                    tile = e.data_segment().data(bgr2rgb=False, resize=True, order=1)
                    index = tuple(slice(i - j, i - j + k) for i, j, k in zip(e.start, f.start, tile.shape))
                    image[index] = tile
                    blocks.append({"start": [int(s) for s in e.start], "shape": [int(s) for s in tile.shape],
                                   "sha1": hashlib.sha1(np.ascontiguousarray(tile).tobytes()).hexdigest()})
                sq = np.squeeze(image)
                assert np.array_equal(sq, np.squeeze(data)), key     # the reference reads back what was written
                out[key] = {"shape": [int(s) for s in f.shape], "start": [int(s) for s in f.start],
                            "axes": f.axes.decode(), "dtype": np.dtype(f.dtype).name, "blocks": blocks,
                            "squeezed_shape": list(sq.shape),
                            "squeezed_sha1": hashlib.sha1(np.ascontiguousarray(sq).tobytes()).hexdigest()}
    with open(os.path.join(HERE, "czi_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_czi_golden.py (reference lib/czifile.py)", "files": out}, f, indent=1)
    for k, v in out.items():
        print(k, v["axes"], v["shape"], v["start"], v["dtype"], len(v["blocks"]))


if __name__ == "__main__":
    main()
