"""utils/tiffio.py and frames.TiffData: uncompressed TIFF stacks either side of update_data (the reference goes
through tifffile, imgutils.py:18-27 / data_model.py:178-218).  Cross-checked against PIL's independent TIFF codec in
both directions.  CPU only."""
import os
import struct

import numpy as np
import pytest

from spimagine_b200 import frames
from spimagine_b200.utils import tiffio

PIL_Image = pytest.importorskip("PIL.Image")


def _stack(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return rng.normal(size=shape).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(info.min, int(info.max) + 1, size=shape, dtype=np.int64).astype(dtype)


def _pil_pages(fn):
    im = PIL_Image.open(fn)
    out = []
    for i in range(im.n_frames):
        im.seek(i)
        out.append(np.array(im))
    return np.stack(out)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.int32, np.float32, np.float64, np.uint32])
@pytest.mark.parametrize("big", [False, True])
def test_round_trip_every_type(tmp_path, dtype, big):
    a = _stack((5, 7, 9), dtype)
    fn = str(tmp_path / "a.tif")
    tiffio.write3dTiff(a, fn, bigtiff=big)
    b = tiffio.read3dTiff(fn)
    assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
    t = tiffio.TiffFile(fn)
    assert t._big == big and t._flat is not None and t.n_images == 5
    part = np.empty((2, 7, 9), t.dtype)
    t.read_into(part, first=2, count=2)
    assert np.array_equal(part, a[2:4])
    with pytest.raises(IndexError):
        t.read_into(part, first=4, count=2)
    with pytest.raises(ValueError):
        t.read_into(np.empty((2, 7, 9), np.int8), first=0, count=2)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_pil_reads_what_we_write(tmp_path, dtype):
    a = _stack((4, 6, 11), dtype, seed=1)
    fn = str(tmp_path / "a.tif")
    tiffio.imsave(fn, a)
    assert np.array_equal(_pil_pages(fn), a)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.int32])
def test_we_read_what_pil_writes(tmp_path, dtype):
    """PIL interleaves image directories and pixel data and cuts pages into several strips."""
    a = _stack((3, 150, 40), dtype, seed=2)
    fn = str(tmp_path / "p.tif")
    pages = [PIL_Image.fromarray(x) for x in a]
    pages[0].save(fn, save_all=True, append_images=pages[1:])
    t = tiffio.TiffFile(fn)
    assert t.shape == a.shape
    assert np.array_equal(t.asarray(), a)
    one = np.empty((1,) + a.shape[1:], t.dtype)
    t.read_into(one, first=1, count=1)
    assert np.array_equal(one[0], a[1])


def test_shapes_2d_4d_and_imagej_hyperstack(tmp_path):
    fn = str(tmp_path / "a.tif")
    img = _stack((6, 5), np.uint16)
    tiffio.write3dTiff(img, fn)
    assert tiffio.read3dTiff(fn).shape == (1, 6, 5)
    vol4 = _stack((3, 4, 6, 5), np.uint16, seed=3)
    tiffio.write3dTiff(vol4, fn)
    t = tiffio.TiffFile(fn)
    assert t.shape == vol4.shape and np.array_equal(t.asarray(), vol4)
    with pytest.raises(ValueError):
        tiffio.write3dTiff(np.zeros((2, 2, 2, 2, 2), np.uint8), fn)
    with pytest.raises(tiffio.TiffError):
        tiffio.write3dTiff(np.zeros((2, 2), np.complex64), fn)


def test_imagej_single_directory_layout(tmp_path):
    """ImageJ beyond 4 GB: ONE image directory whose description says images=N, all N images back to back."""
    vol4 = _stack((2, 3, 6, 5), np.uint16, seed=4)
    fn = str(tmp_path / "ij.tif")
    tiffio.write3dTiff(vol4, fn)
    raw = bytearray(open(fn, "rb").read())
    first_ifd = struct.unpack("<I", raw[4:8])[0]
    n_ent = struct.unpack("<H", raw[first_ifd:first_ifd + 2])[0]
    nxt = first_ifd + 2 + 12 * n_ent
    raw[nxt:nxt + 4] = struct.pack("<I", 0)          # cut the chain after the first directory
    open(fn, "wb").write(bytes(raw))
    t = tiffio.TiffFile(fn)
    assert len(t.pages) == 1 and t.n_images == 6 and t.shape == vol4.shape
    assert np.array_equal(t.asarray(), vol4)
    assert raw.count(b"images=6\n") == 1
    lying = bytes(raw[:nxt + 4]).replace(b"images=6\n", b"images=9\n")   # promises more images than the file holds
    open(fn, "wb").write(lying)
    with pytest.raises(tiffio.TiffError):
        tiffio.TiffFile(fn)


def _big_endian_tiff(fn, a):
    """hand-built 'MM' classic TIFF, two strips per page, no StripByteCounts shortcut"""
    n, ny, nx = a.shape
    be = a.astype(a.dtype.newbyteorder(">"))
    rows = (ny + 1) // 2
    with open(fn, "wb") as f:
        f.write(struct.pack(">2sHI", b"MM", 42, 0))
        ifd_positions = []
        strips = []
        for i in range(n):
            s0 = f.tell()
            f.write(be[i, :rows].tobytes())
            f.write(b"\0\0")                             # a gap: the strips are not contiguous
            s1 = f.tell()
            f.write(be[i, rows:].tobytes())
            strips.append((s0, s1))
        for i in range(n):
            if f.tell() % 2:
                f.write(b"\0")
            at = f.tell()
            ifd_positions.append(at)
            n_ent = 9
            extra = at + 2 + 12 * n_ent + 4
            ents = [(256, 3, 1, nx << 16), (257, 3, 1, ny << 16), (258, 3, 1, (8 * a.dtype.itemsize) << 16),
                    (259, 3, 1, 1 << 16), (262, 3, 1, 1 << 16), (273, 4, 2, extra), (277, 3, 1, 1 << 16),
                    (278, 3, 1, rows << 16), (279, 4, 2, extra + 8)]
            f.write(struct.pack(">H", n_ent))
            for e in ents:
                f.write(struct.pack(">HHII", *e))
            f.write(struct.pack(">I", 0))                # patched below
            f.write(struct.pack(">II", *strips[i]))
            f.write(struct.pack(">II", rows * nx * a.dtype.itemsize, (ny - rows) * nx * a.dtype.itemsize))
        for i, at in enumerate(ifd_positions):
            f.seek(at + 2 + 12 * 9)
            f.write(struct.pack(">I", ifd_positions[i + 1] if i + 1 < n else 0))
        f.seek(4)
        f.write(struct.pack(">I", ifd_positions[0]))


def test_big_endian_multi_strip(tmp_path):
    a = _stack((3, 7, 5), np.uint16, seed=5)
    fn = str(tmp_path / "mm.tif")
    _big_endian_tiff(fn, a)
    t = tiffio.TiffFile(fn)
    assert t.dtype == np.dtype(">u2") and t._flat is None
    got = t.asarray()
    assert got.dtype.isnative and np.array_equal(got, a)
    d = frames.TiffData(fn)
    assert d.size() == (1, 3, 7, 5) and d.dtype == np.uint16 and np.array_equal(d[0], a)


def test_unsupported_files_say_why(tmp_path):
    fn = str(tmp_path / "c.tif")
    PIL_Image.fromarray(_stack((8, 8), np.uint8)).save(fn, compression="jpeg")
    with pytest.raises(tiffio.TiffError, match="259"):
        tiffio.TiffFile(fn)
    PIL_Image.fromarray(_stack((8, 8), np.uint8)).save(fn, compression="packbits", tiffinfo={317: 2})
    with pytest.raises(tiffio.TiffError, match="317"):
        tiffio.TiffFile(fn)
    PIL_Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(fn)
    with pytest.raises(tiffio.TiffError, match="277"):
        tiffio.TiffFile(fn)
    open(fn, "wb").write(b"not a tiff at all")
    with pytest.raises(tiffio.TiffError):
        tiffio.TiffFile(fn)
    with pytest.raises(Exception, match="couldnt open"):
        frames.TiffData(fn)


def test_tiffdata_container_and_frame_source(tmp_path):
    """data_model.py:178-218 shapes: 2-d -> (1, 1, Y, X), 3-d -> (1, Z, Y, X), 4-d as it is; time points are read
    straight into the reader's ring buffers."""
    fn = str(tmp_path / "t.tif")
    data = _stack((5, 4, 6, 7), np.uint16, seed=6)
    tiffio.write3dTiff(data, fn)
    d = frames.TiffData(fn)
    assert d.size() == data.shape and d.sizeT() == 5 and d.stackUnits == [1., 1., 1.]
    assert np.array_equal(d[3], data[3])
    with pytest.raises(IndexError):
        d[5]
    src = frames.FrameSource(d, frames=[4, 0, 2], depth=3, pinned=False)
    try:
        for t in (4, 0, 2, 4):
            assert np.array_equal(src[t], data[t])
    finally:
        src.close()
    tiffio.write3dTiff(data[0], fn)
    assert frames.TiffData(fn).size() == (1, 4, 6, 7)
    tiffio.write3dTiff(data[0, 0], fn)
    assert frames.TiffData(fn).size() == (1, 1, 6, 7)
    tiffio.write3dTiff(data[:, :1], fn)               # (T, 1, Y, X) squeezes to one volume of T slices
    assert frames.TiffData(fn).size() == (1, 5, 6, 7)
    f32 = _stack((3, 4, 5), np.float32, seed=7)
    tiffio.write3dTiff(f32, fn)
    d = frames.TiffData(fn)
    assert d.dtype == np.float32 and np.array_equal(d[0], f32)


def test_one_file_per_time_point(tmp_path):
    """data_model.py:262-404: RawMultipleFiles, TiffFolderData, TiffMultipleFiles."""
    data = _stack((4, 3, 6, 7), np.uint16, seed=8)
    folder = tmp_path / "series"
    folder.mkdir()
    names = []
    for t in (2, 0, 3, 1):                                  # written out of order: the folder is sorted by name
        fn = str(folder / ("t%03d.tif" % t))
        tiffio.write3dTiff(data[t], fn)
        names.append(fn)
    (folder / "notes.txt").write_text("not an image")
    d = frames.TiffFolderData(str(folder))
    assert d.size() == data.shape and d.sizeT() == 4 and d.dtype == np.uint16
    for t in range(4):
        assert np.array_equal(d[t], data[t])
    assert d[4] is None
    with pytest.raises(IndexError):
        d.read_into(4, np.empty(data.shape[1:], np.uint16))
    m = frames.TiffMultipleFiles(sorted(names)[::-1])
    assert m.size() == data.shape and np.array_equal(m[0], data[3])
    tiffio.write3dTiff(data[0, :2], str(folder / "t001.tif"))   # a file that does not match the first one
    with pytest.raises(ValueError):
        d[1]
    with pytest.raises(Exception, match="empty"):
        frames.TiffFolderData(str(tmp_path))
    src = frames.FrameSource(d, frames=[0, 2, 3], depth=3, pinned=False)
    try:
        for t in (0, 2, 3):
            assert np.array_equal(src[t], data[t])
    finally:
        src.close()
    # raw files
    raws = []
    f32 = _stack((3, 4, 5, 6), np.float32, seed=9)
    for t in range(3):
        fn = str(tmp_path / ("r%d.raw" % t))
        f32[t].tofile(fn)
        raws.append(fn)
    r = frames.RawMultipleFiles(raws, shape=(4, 5, 6), dtype=np.float32)
    assert r.size() == f32.shape and r.dtype == np.float32 and np.array_equal(r[2], f32[2])
    r2 = frames.RawMultipleFiles(raws, shape=(20, 6), dtype=np.float32)      # 2-d shape -> one slice per file
    assert r2.size() == (3, 1, 20, 6) and np.array_equal(r2[1][0], f32[1].reshape(20, 6))
    with pytest.raises(Exception, match="couldnt open"):
        frames.RawMultipleFiles(raws, shape=(40, 5, 6), dtype=np.float32)
    with pytest.raises(ValueError):
        frames.RawMultipleFiles(raws)


def test_random_stacks_round_trip(tmp_path):
    """property test: any 2/3/4-d stack of a supported element type survives write -> read, classic and BigTIFF,
    whole and in ranges of pages"""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    fn = str(tmp_path / "h.tif")

    @settings(max_examples=40, deadline=None)
    @given(st.sampled_from(["uint8", "int8", "uint16", "int16", "uint32", "int32", "float32", "float64", "uint64"]),
           st.lists(st.integers(1, 9), min_size=2, max_size=4), st.booleans(), st.integers(0, 2 ** 31))
    def check(dtype, shape, big, seed):
        a = _stack(tuple(shape), dtype, seed=seed) if np.dtype(dtype).itemsize < 8 or np.dtype(dtype).kind == "f" else (
            np.random.default_rng(seed).integers(0, 2 ** 63, size=tuple(shape)).astype(dtype))
        tiffio.write3dTiff(a, fn, bigtiff=big)
        t = tiffio.TiffFile(fn)
        want = a.reshape((1,) + a.shape) if a.ndim == 2 else a
        if a.ndim == 4 and a.shape[0] == 1:
            want = a[0]                                    # frames = 1: a plain stack of slices
        got = t.asarray()
        assert got.dtype == a.dtype and got.shape == want.shape and np.array_equal(got, want)
        n = t.n_images
        first = seed % n
        count = 1 + (seed // 7) % (n - first)
        part = np.empty((count,) + a.shape[-2:], t.dtype)
        t.read_into(part, first=first, count=count)
        assert np.array_equal(part, a.reshape((-1,) + a.shape[-2:])[first:first + count])

    check()
