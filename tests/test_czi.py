"""utils/cziio.py and frames.CZIData against the reference's vendored czifile: tests/golden/czi_ref.json holds what
spimagine/lib/czifile.py reads from the synthetic files of tests/golden/czi_inputs.py (make_czi_golden.py): shape,
start, axes, dtype, every sub-block and the squeezed array of readCziFile.  CPU only."""
import hashlib
import json
import os
import struct
import sys

import numpy as np
import pytest

from spimagine_b200 import frames
from spimagine_b200.utils import cziio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
try:
    import czi_inputs
finally:
    sys.path.remove(GOLDEN)


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("key", sorted(czi_inputs.cases()))
def test_files_read_like_the_references_reader_reads_them(tmp_path, key):
    with open(os.path.join(GOLDEN, "czi_ref.json")) as f:
        want = json.load(f)["files"][key]
    data, axes, block_axes, starts, mosaic = czi_inputs.cases()[key]
    fn = str(tmp_path / (key + ".czi"))
    czi_inputs.write_case(fn, key)
    c = cziio.CziFile(fn)
    # czifile appends a samples axis "0" of length 1 to greyscale files
    assert c.axes + "0" == want["axes"] and list(c.shape) + [1] == want["shape"] and list(c.start) + [0] == want["start"]
    assert c.dtype.name == want["dtype"] and len(c.blocks) == len(want["blocks"])
    with open(fn, "rb") as f:
        for b, w in zip(c.blocks, want["blocks"]):
            assert list(b.start) + [0] == w["start"] and list(b.shape) + [1] == w["shape"]
            assert _sha(c._pixels(f, b)) == w["sha1"]
    got = cziio.readCziFile(fn)
    assert list(got.shape) == want["squeezed_shape"] and _sha(got) == want["squeezed_sha1"]
    assert got.dtype.isnative and np.array_equal(got, np.squeeze(data))
    # the container: 3-d -> one time point, 4-d -> time points; read_into fills caller memory
    d = frames.CZIData(fn)
    sq = np.squeeze(data)
    assert tuple(d.size()) == (sq.shape if sq.ndim == 4 else (1,) + sq.shape) and d.dtype == data.dtype
    for t in range(d.sizeT()):
        frame = sq[t] if sq.ndim == 4 else sq
        assert np.array_equal(d[t], frame)
        out = np.empty(frame.shape, d.dtype)
        d.read_into(t, out)
        assert np.array_equal(out, frame)
    m = frames.DataModel.fromPath(fn)
    try:
        assert type(m.dataContainer).__name__ == "CZIData" and np.array_equal(m[0], d[0])
    finally:
        m.close()


def test_time_points_are_read_alone(tmp_path):
    data, axes, block_axes, starts, mosaic = czi_inputs.cases()["tzyx_planes_u8"]
    fn = str(tmp_path / "t.czi")
    czi_inputs.write_czi(fn, data, axes, block_axes, starts)
    c = cziio.CziFile(fn)
    reads = []
    pixels = c._pixels
    c._pixels = lambda f, b: (reads.append(b.start), pixels(f, b))[1]
    assert np.array_equal(c.time_point(1), data[1]) and len(reads) == data.shape[1]     # one plane per z, one t
    out = np.full(data.shape[1:], 7, np.uint8)
    assert np.array_equal(c.time_point(2, out=out), data[2]) and np.array_equal(out, data[2])
    with pytest.raises(IndexError):
        c.time_point(3)
    # sub-blocks that hold whole stacks of several... one stack per t: still one block per time point
    data, axes, block_axes, starts, mosaic = czi_inputs.cases()["tzyx_stacks_f32"]
    czi_inputs.write_czi(fn, data, axes, block_axes, starts)
    c = cziio.CziFile(fn)
    assert np.array_equal(c.time_point(1), data[1])


def test_refused_files_say_why(tmp_path):
    data, axes, block_axes, starts, mosaic = czi_inputs.cases()["zyx_planes_u16"]
    fn = str(tmp_path / "x.czi")
    czi_inputs.write_czi(fn, data, axes, block_axes)
    raw = bytearray(open(fn, "rb").read())
    directory = struct.unpack("<q", raw[32 + 52:32 + 60])[0]
    first = directory + 32 + 128                      # first directory entry

    def patched(offset, fmt, value):
        b = bytearray(raw)
        b[first + offset:first + offset + struct.calcsize(fmt)] = struct.pack(fmt, value)
        p = str(tmp_path / "p.czi")
        open(p, "wb").write(b)
        return p

    with pytest.raises(cziio.CziError, match="compress"):
        cziio.CziFile(patched(18, "<i", 1))           # JPEG
    with pytest.raises(cziio.CziError, match="Bgr24"):
        cziio.CziFile(patched(2, "<i", 3))
    with pytest.raises(cziio.CziError, match="pyramid"):
        cziio.CziFile(patched(32 + 16, "<i", 3))      # stored size of X
    (tmp_path / "n.czi").write_bytes(b"not a czi file" * 10)
    with pytest.raises(cziio.CziError, match="not a CZI"):
        cziio.CziFile(str(tmp_path / "n.czi"))
    with pytest.raises(Exception, match="couldnt open .* as CZIData"):
        frames.CZIData(str(tmp_path / "n.czi"))
    # 2-d file: the squeezed array is neither a stack nor a timelapse
    czi_inputs.write_czi(fn, data[:1], axes, block_axes)
    with pytest.raises(Exception, match="couldnt open"):
        frames.CZIData(fn)
