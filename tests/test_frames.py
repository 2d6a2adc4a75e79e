"""Frame sources (spimagine_b200/frames.py): the containers either side of update_data and the prefetching reader.
CPU tests use pageable ring buffers; the page-locked path is exercised in the GPU test at the bottom."""
import os
import threading
import time

import numpy as np
import pytest

import scenes
from spimagine_b200 import frames

REF_FIXTURE = "/root/reference/tests/data/spimdata"


def _timelapse(nt=7, shape=(12, 10, 14), seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 65536, size=(nt,) + shape).astype(np.uint16)


def test_spim_folder_round_trip(tmp_path):
    data = _timelapse()
    folder = str(tmp_path / "spim")
    frames.createSpimFolder(folder, data, stackUnits=(.162, .162, .5))
    d = frames.SpimData(folder)
    assert d.size() == list(data.shape) and d.sizeT() == 7 and len(d) == 7
    assert d.stackUnits[:2] == (.162, .162) and abs(d.stackUnits[2] - .5) < 1e-2   # StopZ is written with 2 decimals
    for t in (0, 3, 6):
        assert np.array_equal(d[t], data[t]) and d[t].dtype == np.uint16
    with pytest.raises(IndexError):
        d[7]
    with pytest.raises(IndexError):
        d[-1]
    out = np.empty(data.shape[1:], np.uint16)
    d.read_into(5, out)
    assert np.array_equal(out, data[5])
    with pytest.raises(Exception):
        frames.SpimData(str(tmp_path / "nothing_here"))


@pytest.mark.skipif(not os.path.isdir(REF_FIXTURE), reason="the reference tree is only present in the build container")
def test_reads_the_references_own_spim_fixture():
    """tests/data/spimdata of the reference: 10 x 32^3 uint16, read by the rules of imgutils.fromSpimFolder."""
    d = frames.SpimData(REF_FIXTURE)
    assert d.size() == [10, 32, 32, 32]
    raw = np.fromfile(os.path.join(REF_FIXTURE, "data/data.bin"), dtype="<u2").reshape(10, 32, 32, 32)
    for t in range(10):
        assert np.array_equal(d[t], raw[t])
    assert d.stackUnits == (.162, .162, (5.02 - 0) / (32 - 1.))


def test_raw_and_numpy_containers(tmp_path):
    data = _timelapse(4, (6, 5, 8), seed=1).astype(np.float32)
    fn = str(tmp_path / "stack.raw")
    data.tofile(fn)
    r = frames.RawData(fn, shape=data.shape, dtype=np.float32)
    assert r.size() == data.shape and r.dtype == np.float32
    assert np.array_equal(r[2], data[2])
    r3 = frames.RawData(fn, shape=(24, 5, 8), dtype=np.float32)      # 3-D shape -> one time point
    assert r3.size() == (1, 24, 5, 8) and np.array_equal(r3[0], data.reshape(24, 5, 8))
    with pytest.raises(ValueError):
        frames.RawData(fn, shape=(1, 1, 4, 6, 5, 8), dtype=np.float32)
    n = frames.NumpyData(data[1])
    assert n.size() == (1, 6, 5, 8) and np.array_equal(n[0], data[1])
    assert frames.NumpyData(np.zeros((3, 4))).size() == (1, 1, 3, 4)     # 2-d: one slice (data_model.py:416-418)
    with pytest.raises(TypeError):
        frames.NumpyData(np.zeros(3))


@pytest.mark.parametrize("depth", [2, 3, 5])
def test_frame_source_plays_in_order_and_keeps_frames_valid(tmp_path, depth):
    data = _timelapse(9, (8, 6, 10), seed=2)
    folder = str(tmp_path / "spim")
    frames.createSpimFolder(folder, data)
    order = [1, 4, 7]                                   # rank 1 of 3
    src = frames.FrameSource(frames.SpimData(folder), frames=order, depth=depth, pinned=False)
    try:
        assert len(src) == 9 and src.size() == list(data.shape)
        prev = None
        for lap in range(3):                            # looping playback wraps around
            for t in order:
                a = src[t]
                assert np.array_equal(a, data[t])
                if prev is not None:                    # the previous frame is still intact
                    assert np.array_equal(prev[1], data[prev[0]])
                prev = (t, a)
        with pytest.raises(IndexError):
            src[order[1]]                               # out of play order (next would be order[0])
        assert src.bytes_read >= 9 * data[0].nbytes
    finally:
        src.close()


def test_frame_source_reads_ahead_and_reports_errors(tmp_path):
    data = _timelapse(6, (8, 6, 10), seed=3)

    class Slow(frames.NumpyData):
        def __init__(self, d):
            frames.NumpyData.__init__(self, d)
            self.reads = []

        def read_into(self, pos, out):
            if pos == 4:
                raise IOError("disk on fire")
            time.sleep(0.02)
            self.reads.append(pos)
            frames.NumpyData.read_into(self, pos, out)

    c = Slow(data)
    src = frames.FrameSource(c, depth=4, pinned=False)
    try:
        assert np.array_equal(src[0], data[0])
        time.sleep(0.15)
        assert c.reads == [0, 1, 2]                     # depth - 2 = 2 frames ahead of the one in use, not more
        assert np.array_equal(src[1], data[1])
        assert np.array_equal(src[2], data[2])
        assert np.array_equal(src[3], data[3])
        with pytest.raises(IOError):
            src[4]
    finally:
        src.close()
    assert not any(t.name == "spimagine-frame-reader" and t.is_alive() for t in threading.enumerate())


@pytest.mark.gpu
def test_timelapse_player_streams_from_a_spim_folder(tmp_path):
    """disk -> page-locked ring -> asynchronous upload -> render: every owned time point of a SpimData folder shows
    the image a plain set_data + render of that time point gives."""
    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.multigpu import TimelapsePlayer
    vols = np.stack([scenes.vol_g(40, np.uint16, seed=100 + t, t=t) for t in range(6)])
    folder = str(tmp_path / "spim")
    frames.createSpimFolder(folder, vols)
    M, P = scenes.gui_camera(0.4, 3.3)
    ref = VolumeRenderer((96, 80))
    ref.set_view_copies("primary")  # as the player's renderers: time points are rendered through the z copy
    ref.set_projection(P)
    ref.set_modelView(M)
    want = {}
    for t in range(6):
        ref.set_data(vols[t])
        ref.render(maxVal=60000.)
        want[t] = ref.output.copy()
    ref.close()
    for rank in range(2):
        player = TimelapsePlayer((96, 80), rank=rank, world=2)
        src = frames.FrameSource(frames.SpimData(folder), frames=player.my_frames(6), depth=3)
        seen = []
        for t, r in player.play(src, frames=src.frames, pinned=True, projection=P, max_val=60000.,
                                modelViews={t: M for t in range(6)}):
            assert np.array_equal(r.output, want[t])
            seen.append(t)
        assert seen == list(range(rank, 6, 2))
        src.close()
        player.close()


def test_xwing_folder(tmp_path):
    """data_model.py:475-515 / imgutils.py:89-127: index, JSON metadata, one raw stack per time point."""
    data = _timelapse(3, (5, 6, 7), seed=3)
    root = tmp_path / "xw"
    (root / "stacks" / "default").mkdir(parents=True)
    for t in range(3):
        data[t].astype("<u2").tofile(str(root / "stacks" / "default" / ("%06d.raw" % t)))
    (root / "default.index.txt").write_text("0\t0.000\t7, 6, 5\n1\t0.100\t7, 6, 5\n")
    (root / "default.metadata.txt").write_text('{"VoxelDimX": 0.26, "VoxelDimY": 0.26, "VoxelDimZ": 1.5}\n{"x": 1}\n')
    d = frames.XwingData(str(root))
    assert d.size() == [3, 5, 6, 7] and d.stackUnits == (.26, .26, 1.5) and d.dtype == np.uint16
    for t in range(3):
        assert np.array_equal(d[t], data[t])
    with pytest.raises(IndexError):
        d[3]
    out = np.empty((5, 6, 7), np.uint16)
    d.read_into(1, out)
    assert np.array_equal(out, data[1])
    (root / "default.metadata.txt").write_text("not json\n")
    assert frames.XwingData(str(root)).stackUnits == (1., 1., 1.)
    with pytest.raises(Exception, match="couldnt open"):
        frames.XwingData(str(tmp_path / "missing"))


def test_containers_read_what_the_references_containers_read(tmp_path):
    """tests/golden/frames_ref.json: the reference's own SpimData / RawData / RawMultipleFiles / XwingData / NumpyData
    and fromSpimFolder run on the seeded inputs of tests/golden/frames_inputs.py (make_frames_golden.py); this
    package's containers must report the same sizes and units and return the same bytes for every time point."""
    import hashlib
    import json
    import sys
    golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden_dir)
    try:
        import frames_inputs
    finally:
        sys.path.remove(golden_dir)
    with open(os.path.join(golden_dir, "frames_ref.json")) as f:
        ref = json.load(f)["containers"]
    specs = frames_inputs.build(str(tmp_path))
    extra = {k for k in ref if k.startswith(("fromSpimFolder", "DemoData", "EmptyData", "OverlayData"))}
    assert set(specs) | extra | {"DataModel"} == set(ref)
    for key, want in ref["DataModel"].items():
        m = frames.DataModel.fromPath(os.path.join(str(tmp_path), key), prefetchSize=2)
        try:
            assert type(m.dataContainer).__name__ == want["container"] and m.prefetchSize == want["prefetchSize"]
            assert m.sizeT() == want["sizeT"] and m.pos == want["pos_after_init"]
            for p, nb in want["neighborhood"].items():
                assert m.neighborhood(int(p)).tolist() == nb
            assert hashlib.sha1(np.ascontiguousarray(m[1]).tobytes()).hexdigest() == want["item1_sha1"]
        finally:
            m.close()
    for key, (cls, args, kw) in specs.items():
        c = getattr(frames, cls)(*args, **kw)
        want = ref[key]
        assert [int(s) for s in c.size()] == want["size"] and int(c.sizeT()) == want["sizeT"] == len(c), key
        assert np.allclose([float(u) for u in c.stackUnits], want["stackUnits"], rtol=1e-12, atol=0), key
        for t, p in enumerate(want["points"]):
            a = np.ascontiguousarray(c[t])
            assert list(a.shape) == p["shape"] and a.dtype.name == p["dtype"], (key, t)
            assert hashlib.sha1(a.tobytes()).hexdigest() == p["sha1"], (key, t)
            out = np.empty(a.shape, c.dtype)
            c.read_into(t, out)
            assert np.array_equal(out, a)
    d = frames.SpimData(specs["SpimData"][1][0])
    sub = np.stack([d[2], d[3]])                       # fromSpimFolder(pos=2, count=2), imgutils.py:129-146
    want = ref["fromSpimFolder_pos2_count2"]
    assert list(sub.shape) == want["shape"] and hashlib.sha1(sub.tobytes()).hexdigest() == want["sha1"]
    # fromSpimFolder itself: a window, a window that leaves the folder, "everything behind pos", a negative pos
    got = frames.fromSpimFolder(specs["SpimData"][1][0], pos=2, count=2)
    assert got.dtype == np.dtype("<u2") and np.array_equal(got, sub)
    for key in sorted(extra):
        if key.startswith("fromSpimFolder") and "pos" in ref[key]:
            got = frames.fromSpimFolder(specs["SpimData"][1][0], pos=ref[key]["pos"], count=ref[key]["count"])
            assert list(got.shape) == ref[key]["shape"], key
            assert hashlib.sha1(np.ascontiguousarray(got).tobytes()).hexdigest() == ref[key]["sha1"], key


def test_demo_and_empty_containers_equal_the_references():
    """DemoData(24) and EmptyData of the reference (data_model.py:434-472, 518-531) recorded by make_frames_golden.py:
    same sizes, same float32 voxels bit for bit at three time points."""
    import hashlib
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_ref.json")) as f:
        ref = json.load(f)["containers"]
    want = ref["DemoData_24"]
    d = frames.DemoData(24)
    assert list(d.size()) == want["size"] and d.sizeT() == want["sizeT"] == len(d) and d.dtype == np.float32
    assert [float(u) for u in d.stackUnits] == want["stackUnits"]
    for t, p in want["points"].items():
        a = d[int(t)]
        assert a.dtype.name == p["dtype"] and a.flags.c_contiguous
        assert abs(float(a.max()) - p["max"]) <= 1e-6 * p["max"], t
        assert hashlib.sha1(a.tobytes()).hexdigest() == p["sha1"], t
    out = np.empty(d.size()[1:], np.float32)
    d.read_into(3, out)
    assert np.array_equal(out, d[3])
    big = frames.DemoData()                 # the reference ships a logo stack of this size instead
    assert big.size() == (10, 80, 80, 80) and big.sizeT() == 10 and big[9].dtype == np.float32
    e = frames.EmptyData()
    want = ref["EmptyData"]
    assert list(e.size()) == want["size"] and e.sizeT() == want["sizeT"] and [float(u) for u in e.stackUnits] == want["stackUnits"]
    assert e[0].dtype.name == want["dtype"] and list(e[0].shape) == want["shape"] and int(e[0].sum()) == want["sum"]
    import spimagine_b200
    assert spimagine_b200.DemoData is frames.DemoData


class _Counting(frames.NumpyData):
    def __init__(self, data):
        frames.NumpyData.__init__(self, data)
        self.reads = []

    def __getitem__(self, pos):
        self.reads.append(int(pos))
        return self.data[pos]


def _wait_for(cond, timeout=10.):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > timeout:
            return False
        time.sleep(.002)
    return True


def test_data_model_keeps_the_neighbourhood_loaded():
    """data_model.py:600-757: position, cache, neighbourhood pos .. pos + prefetchSize (mod sizeT) kept loaded in the
    background, everything else dropped."""
    data = _timelapse(8, (3, 4, 5), seed=5)
    c = _Counting(data)
    seen = []
    m = frames.DataModel(c, prefetchSize=2)
    try:
        m.pos_changed.append(seen.append)
        assert m.sizeT() == 8 and m.size() == (8, 3, 4, 5) and m.name() == "NumpyData" and m.stackUnits() == [1., 1., 1.]
        assert m.neighborhood(6).tolist() == [6, 7, 0] and m.pos == 0
        assert _wait_for(lambda: set(m.data) == {0, 1, 2})
        assert np.array_equal(m[1], data[1])                 # already there: no second read of time point 1
        assert _wait_for(lambda: set(m.data) == {1, 2, 3}) and c.reads.count(1) == 1
        m.setPos(6)
        assert seen == [6] and _wait_for(lambda: set(m.data) == {6, 7, 0})
        m.setPos(6)
        assert seen == [6]                                   # unchanged position: no signal
        assert np.array_equal(m[4], data[4]) and _wait_for(lambda: set(m.data) == {4, 5, 6})
        with pytest.raises(IndexError):
            m.setPos(8)
        with pytest.raises(IndexError):
            m.setPos(-1)
        # a new container: the cache starts over, the old reader is gone
        c2 = _Counting(data[:2])
        m.setContainer(c2, prefetchSize=0)
        assert m.sizeT() == 2 and m.pos == 0 and _wait_for(lambda: set(m.data) == {0}) and np.array_equal(m[1], data[1])
    finally:
        m.close()
    assert m._thread is None
    assert frames.DataModel().sizeT() is None


def test_data_model_chooses_the_container_from_the_path(tmp_path):
    """data_model.py:733-757"""
    from spimagine_b200.utils import tiffio
    data = _timelapse(3, (4, 5, 6), seed=6)
    spim = str(tmp_path / "spim")
    frames.createSpimFolder(spim, data)
    xw = tmp_path / "xw"
    (xw / "stacks" / "default").mkdir(parents=True)
    for t in range(3):
        data[t].astype("<u2").tofile(str(xw / "stacks" / "default" / ("%06d.raw" % t)))
    (xw / "default.index.txt").write_text("0\t0.0\t6, 5, 4\n")
    (xw / "default.metadata.txt").write_text('{"VoxelDimX": 1, "VoxelDimY": 1, "VoxelDimZ": 2}\n')
    tifs = tmp_path / "tifs"
    tifs.mkdir()
    names = []
    for t in range(3):
        names.append(str(tifs / ("t%d.tif" % t)))
        tiffio.write3dTiff(data[t], names[-1])
    one = str(tmp_path / "all.tiff")
    tiffio.write3dTiff(data, one)
    for path, cls, prefetch in ((spim, frames.SpimData, 1), (str(xw), frames.XwingData, 1), (str(tifs), frames.TiffFolderData, 1),
                                (names, frames.TiffMultipleFiles, 1), (one, frames.TiffData, 0)):
        m = frames.DataModel.fromPath(path, prefetchSize=1)
        try:
            assert type(m.dataContainer) is cls and m.prefetchSize == prefetch, path
            assert m.sizeT() == 3 and np.array_equal(m[2], data[2]), path
        finally:
            m.close()
    for bad in (str(tmp_path / "x.h5"), [str(tmp_path / "a.raw")]):
        with pytest.raises(ValueError):
            frames.DataModel.fromPath(bad)
    with pytest.raises(Exception, match="couldnt open .* as CZIData"):      # chosen by extension, file missing
        frames.DataModel.fromPath(str(tmp_path / "x.czi"))


def test_img2d_container(tmp_path):
    """data_model.py:150-175, :770-771: a 2-d image is a (1, 1, Y, X) stack; loadFromPath picks it by extension"""
    PIL_Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(11)
    g8 = rng.integers(0, 256, (13, 17), dtype=np.uint8)
    g16 = rng.integers(0, 65536, (13, 17), dtype=np.uint16)
    rgb = rng.integers(0, 256, (13, 17, 3), dtype=np.uint8)
    PIL_Image.fromarray(g8).save(str(tmp_path / "a.png"))
    PIL_Image.fromarray(g16).save(str(tmp_path / "b.png"))
    PIL_Image.fromarray(rgb).save(str(tmp_path / "c.bmp"))
    d = frames.Img2dData(str(tmp_path / "a.png"))
    assert d.size() == (1, 1, 13, 17) and d.sizeT() == 1 and d.dtype == np.uint8 and np.array_equal(d[0][0], g8)
    d = frames.Img2dData(str(tmp_path / "b.png"))
    assert d.dtype == np.uint16 and np.array_equal(d[0][0], g16)
    d = frames.Img2dData(str(tmp_path / "c.bmp"))
    lum = np.asarray(PIL_Image.fromarray(rgb).convert("L"))
    assert d.dtype == np.uint8 and d[0].shape == (1, 13, 17) and np.array_equal(d[0][0], lum)
    m = frames.DataModel.fromPath(str(tmp_path / "a.png"))
    try:
        assert type(m.dataContainer).__name__ == "Img2dData" and m.sizeT() == 1 and np.array_equal(m[0][0], g8)
    finally:
        m.close()
    (tmp_path / "bad.png").write_bytes(b"not an image")
    with pytest.raises(Exception, match="couldnt open .* as Img2dData"):
        frames.Img2dData(str(tmp_path / "bad.png"))


def test_overlay_container_equals_the_references():
    """models/overlay_volumes.py driven through a walk along each axis by make_frames_golden.py: same sizes, same
    bytes at every stop; data[0] is x, data[n] is y, the returned array is reused"""
    import hashlib
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_ref.json")) as f:
        ref = json.load(f)["containers"]["OverlayData_seed21_5x6x7"]
    rng = np.random.default_rng(21)
    x, y = rng.integers(0, 60000, (5, 6, 7)).astype(np.uint16), rng.integers(0, 60000, (5, 6, 7)).astype(np.uint16)
    for axis, want in ref.items():
        o = frames.OverlayData(x, y, axis=int(axis))
        assert list(o.size()) == want["size"] and o.sizeT() == want["sizeT"] == len(o) and o.dtype == np.uint16
        for i, sha in zip(want["walk"], want["sha1"]):
            assert hashlib.sha1(np.ascontiguousarray(o[i]).tobytes()).hexdigest() == sha, (axis, i)
        n = o.size()[0] - 1
        assert np.array_equal(o[0], x) and np.array_equal(o[n], y) and o[1] is o[2]
        out = np.empty(x.shape, np.uint16)
        o.read_into(2, out)
        idx = [slice(None)] * 3
        idx[int(axis)] = slice(0, 2)
        assert np.array_equal(out[tuple(idx)], y[tuple(idx)])
    with pytest.raises(ValueError):
        frames.OverlayData(x, y[:4])
    import spimagine_b200
    assert spimagine_b200.OverlayData is frames.OverlayData
