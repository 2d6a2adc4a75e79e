"""Keyframe paths (spimagine_b200/keyframes.py, utils/quaternion.py) against vectors produced by the reference's own
keyframe_model / transform_model (tests/golden/keyframes_ref.json, written by tests/golden/make_keyframe_golden.py),
the batch render loop and the headless CLI on the device."""
import json
import os

import numpy as np
import pytest

import scenes
from spimagine_b200 import keyframes as kf
from spimagine_b200.utils.quaternion import Quaternion, quaternion_slerp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keyframes_ref.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _same_transform(td, ref, tol=0.):
    assert np.allclose(td.quatRot.data, ref["quatRot"], rtol=0, atol=tol)
    for k in ("zoom", "minVal", "maxVal", "gamma", "alphaPow"):
        assert abs(getattr(td, k) - ref[k]) <= tol * max(1., abs(ref[k])), k
    for k in ("translate", "bounds"):
        assert np.allclose(getattr(td, k), ref[k], rtol=0, atol=tol), k
    for k in ("dataPos", "isBox", "isIso", "isSlice", "slicePos", "sliceDim"):
        assert getattr(td, k) == ref[k], k


# ------------------------------------------------------------------ quaternions, easing
def test_quaternion_tables(golden):
    for row in golden["slerp"]:
        a, b = Quaternion(*row["a"]), Quaternion(*row["b"])
        assert np.array_equal(quaternion_slerp(a, b, row["t"]).data, row["q"])
        assert np.array_equal((a * b).data, row["prod_ab"])
        assert np.array_equal(a.toRotation4(), np.array(row["rot4_a"]))
    q = Quaternion(.5, -.5, .5, .5)
    assert np.allclose((q * q.conj()).data, [1, 0, 0, 0])
    assert np.array_equal((q * 2.).data, 2 * q.data) and np.allclose(q.normalize().norm(), 1)
    assert np.array_equal(q.toRotation3(), q.toRotation4()[:3, :3])
    c = Quaternion.copy(q)
    c[0] = 9.
    assert q[0] == .5


def test_easing_table(golden):
    for row in golden["ease"]:
        assert kf.create_interp_func(row["a"])(row["x"]) == row["y"]


# ------------------------------------------------------------------ the reference's file, its transforms, its cameras
def test_reference_file_plays_identically(golden):
    """Loading the file the reference's encoder wrote gives the transforms the reference's loader gives (which drops
    interp_elasticity); rebuilding the list in memory gives the eased ones.  Bit for bit."""
    loaded = kf.KeyFrameList._from_JSON(golden["keyframes_json"])
    assert len(loaded) == 6 and loaded._countID == 6
    spec = json.loads(golden["keyframes_json"])
    built = kf.KeyFrameList()
    for ID in sorted(spec["items"], key=int):
        it = spec["items"][ID]
        t = dict(it["transformData"])
        td = kf.TransformData(quatRot=Quaternion(*t.pop("quatRot")), **t)
        built.addItem(kf.KeyFrame(it["pos"], td, it["interp_elasticity"]))
    for s in golden["samples"]:
        _same_transform(loaded.getTransform(s["t"]), s["transform_after_reload"])
        _same_transform(built.getTransform(s["t"]), s["transform"])
    kf.KeyFrameDecoder.keep_elasticity = True
    try:
        eased = kf.KeyFrameList._from_JSON(golden["keyframes_json"])
    finally:
        kf.KeyFrameDecoder.keep_elasticity = False
    for s in golden["samples"]:
        _same_transform(eased.getTransform(s["t"]), s["transform"])


def test_cameras_and_renderer_settings(golden):
    class Recorder(object):
        def __getattr__(self, name):
            assert name.startswith("set_")
            return lambda v: setattr(self, "_" + name[4:], v)

    built = kf.KeyFrameDecoder.keep_elasticity
    kf.KeyFrameDecoder.keep_elasticity = True
    try:
        keys = kf.KeyFrameList._from_JSON(golden["keyframes_json"])
    finally:
        kf.KeyFrameDecoder.keep_elasticity = built
    for s in golden["samples"]:
        td = keys.getTransform(s["t"])
        for persp, tag in ((True, "persp"), (False, "ortho")):
            M, P = kf.camera_of(td, persp)
            assert np.array_equal(M, np.array(s["modelView_" + tag])), (s["t"], tag)
            assert np.array_equal(P, np.array(s["projection_" + tag]))
            r = Recorder()
            M2, method = kf.apply_transform(r, td, persp)
            ref = s["renderer"]
            assert np.array_equal(M2, M) and np.array_equal(r._modelView, M) and np.array_equal(r._projection, P)
            assert (r._min_val, r._max_val, r._gamma, r._alpha_pow) == (
                ref["minVal"], ref["maxVal"], ref["gamma"], ref["alphaPow"])
            assert list(r._box_boundaries) == ref["bounds"]
            assert method == ("iso_surface" if ref["isIso"] else "max_project")


def test_json_round_trip_and_old_files():
    k = kf.KeyFrameList()
    k.addItem(kf.KeyFrame(0., kf.TransformData(zoom=np.float32(1.5), dataPos=np.int64(3))))
    k.addItem(kf.KeyFrame(1., kf.TransformData(quatRot=Quaternion(0, 1, 0, 0), bounds=[0] * 6), .7))
    again = kf.KeyFrameList._from_JSON(k._to_JSON())
    assert again._countID == 2 and again.pos_at(0) == 0. and again.pos_at(-1) == 1.
    _same_transform(again.getTransform(.25), json.loads(json.dumps(
        kf.TransformData.interp(k[0].transformData, k[1].transformData, .25), cls=kf.KeyFrameEncoder)))
    # a file of an older spimagine (keyframe_model.py:335-343 "NEW" example): missing keys keep their defaults
    old = ('{"items": {"0": {"transformData": {"dataPos": 0, "quatRot": [1.0, 0.0, 0.0, 0.0], "maxVal": 100.0, '
           '"isIso": false, "minVal": 0.0, "zoom": 1, "bounds": [-1, 1, -1, 1, -1, 1], "isBox": true, '
           '"alphaPow": 0.0, "translate": [0, 0, 0], "gamma": 1.0}, "pos": 0}}, "_countID": 1, "posdict": {"0": 0}}')
    o = kf.KeyFrameList._from_JSON(old)
    td = o.getTransform(.5)
    assert td.isSlice is False and td.slicePos == 0 and td.maxVal == 100.


def test_list_editing(tmp_path):
    """addItem / removeItem / update_pos / item order (keyframe_model.py:204-258), exercised like the reference's
    test_shuffle (:397-420)."""
    k = kf.KeyFrameList()
    with pytest.raises(IndexError):
        k.getTransform(.5)
    k.addItem(kf.KeyFrame(.5, kf.TransformData(zoom=.5)))
    assert k.getTransform(0.).zoom == .5 and k.getTransform(.5).zoom == .5 and k.getTransform(2.).zoom == .5
    rng = np.random.default_rng(0)
    for _ in range(4):
        k.addItem(kf.KeyFrame(rng.uniform(0, 1)))
    for _ in range(100):
        k.addItem(kf.KeyFrame(rng.uniform(0, 1)))
        ID = k.item_id_at(rng.integers(1, len(k.items) - 1))
        k.update_pos(ID, rng.uniform(0, 1))
        ID = k.item_id_at(rng.integers(1, len(k.items) - 1))
        k.removeItem(ID)
        order = [k.pos_at(i) for i in range(len(k))]
        assert order == sorted(order) and len(k.posdict) == len(k.items) == 5
        assert all(k.item_at(i).pos == k.pos_at(i) and k.pos_at_id(k.item_id_at(i)) == k.pos_at(i)
                   for i in range(len(k)))
    taken = k.pos_at(2)
    k.update_pos(k.item_id_at(0), taken)           # refused: position already there
    assert k.pos_at(2) == taken and len(k.posdict) == 5
    k.distribute(10, 30)
    assert all(it.transformData.dataPos == int(10 + 20 * it.pos) for it in k.items.values())
    fn = str(tmp_path / "keys.json")
    k.save_to_JSON(fn)
    assert [kf.KeyFrameList.load_from_JSON(fn).pos_at(i) for i in range(5)] == [k.pos_at(i) for i in range(5)]
    with pytest.raises(TypeError):
        kf.TransformData(zoomm=1)
    assert kf.frame_name(7, 100) == "output_007.png" and kf.frame_name(12, 99) == "output_12.png"
    assert kf.keyframe_times(4) == [(1, .25), (2, .5), (3, .75), (4, 1.)]


# ------------------------------------------------------------------ on the device
def _path(n_t=1):
    k = kf.KeyFrameList()
    k.addItem(kf.KeyFrame(0., kf.TransformData(quatRot=Quaternion(1, 0, 0, 0), zoom=1., maxVal=60000., dataPos=0)))
    k.addItem(kf.KeyFrame(.4, kf.TransformData(quatRot=Quaternion(.8, .3, .5, .1), zoom=1.3, maxVal=30000.,
                                               gamma=.8, bounds=[-.7, 1, -1, .8, -1, 1], dataPos=n_t // 2), 2.))
    k.addItem(kf.KeyFrame(.7, kf.TransformData(quatRot=Quaternion(.2, .9, -.3, .2), zoom=.9, maxVal=40000.,
                                               isIso=True, dataPos=n_t - 1)))
    k.addItem(kf.KeyFrame(1., kf.TransformData(quatRot=Quaternion(0, 0, 1, 0), zoom=1., maxVal=60000., dataPos=0)))
    return k


@pytest.mark.gpu
def test_render_keyframes_equals_the_synchronous_loop():
    """The pipelined path (render_sequence, uploads only when dataPos changes) yields, frame by frame, exactly what
    the reference's loop yields: apply the transform, update_data, render, read back."""
    from spimagine_b200 import VolumeRenderer
    n_t, n_frames = 4, 20
    source = [scenes.vol_g(48, np.uint16, seed=100 + t, t=t) for t in range(n_t)]
    keys = _path(n_t)
    rend = VolumeRenderer((160, 128))
    ref = VolumeRenderer((160, 128))
    try:
        rend.set_data(source[0])
        ref.set_data(source[0])
        n_iso = n_seen = 0
        for pos, td, r in kf.render_keyframes(rend, keys, n_frames, source=source):
            assert td is not None and pos == n_seen + 1
            want = keys.getTransform(1. * pos / n_frames)
            _same_transform(td, json.loads(json.dumps(want, cls=kf.KeyFrameEncoder)))
            ref.update_data(source[want.dataPos])
            _, method = kf.apply_transform(ref, want)
            ref.render(method=method)
            assert np.array_equal(r.output, ref.output) and np.array_equal(r.output_alpha, ref.output_alpha)
            if method == "iso_surface":
                n_iso += 1
                assert np.array_equal(r.output_depth, ref.output_depth)
                assert np.array_equal(r.output_normals, ref.output_normals)
            assert r.output.max() > 0
            n_seen += 1
        assert n_seen == n_frames and 0 < n_iso < n_frames
        # unpipelined flavour gives the same frames
        a = [np.array(r.output) for _, _, r in kf.render_keyframes(rend, keys, 6, source=source, pipelined=False)]
        b = [np.array(r.output) for _, _, r in kf.render_keyframes(rend, keys, 6, source=source)]
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        # display planes only: same images, the other iso planes are not read back
        seen_iso = False
        for (_, td, r), want in zip(kf.render_keyframes(rend, keys, 20, source=source, iso_planes=2),
                                    kf.render_keyframes(ref, keys, 20, source=source, pipelined=False)):
            assert np.array_equal(r.output, want[2].output) and np.array_equal(r.output_alpha, want[2].output_alpha)
            if td.isIso:
                seen_iso = True
                assert r.output_depth is None and r.output_normals is None and r.output_occlusion is None
        assert seen_iso
        rend.render(method="max_project")
        assert rend.output_depth.shape == (128, 160)
        with pytest.raises(ValueError):
            next(rend.render_sequence([np.eye(4)], method="iso_surface", iso_planes=3))
    finally:
        rend.close()
        ref.close()


@pytest.mark.gpu
def test_record_and_cli(tmp_path):
    from PIL import Image
    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.bin import spim_render
    from spimagine_b200.utils import tiffio
    from spimagine_b200.utils.transform_matrices import mat4_perspective
    vol = scenes.vol_g(40, np.uint16, seed=3)
    tif = str(tmp_path / "vol.tif")
    tiffio.write3dTiff(vol, tif)
    # --- single image, the reference CLI's camera
    png = str(tmp_path / "out.png")
    assert spim_render.main(["-i", tif, "-o", png, "-w", "96", "-u", "1", "1", "2", "-t", "0", "0", "-3.5",
                             "-r", ".4", "0", "1", "0"]) == 0
    r = VolumeRenderer((96, 96))
    try:
        r.set_data(vol)
        r.set_units([1, 1, 2])
        r.set_max_val(float(vol.max()))
        ns = spim_render.build_parser().parse_args(["-i", tif, "-t", "0", "0", "-3.5", "-r", ".4", "0", "1", "0"])
        r.set_modelView(spim_render.model_view(ns))
        r.set_projection(mat4_perspective(60, 1., 1, 10))
        r.render()
        assert np.array_equal(np.array(Image.open(png)), spim_render.to_uint8(r.output)) and r.output.max() > .5
        # --- 16 bit, iso, ortho, raw input
        raw = str(tmp_path / "vol.raw")
        vol.tofile(raw)
        png16 = str(tmp_path / "out16.png")
        assert spim_render.main(["-f", "raw", "--shape", "40", "40", "40", "-i", raw, "-o", png16, "-w", "64", "-O",
                                 "--16bit", "-R", "0", "65535", "--iso", "-u", "1", "1", "1"]) == 0
        im = np.array(Image.open(png16))
        assert im.shape == (64, 64) and im.max() >= 65534 and im.dtype in (np.uint16, np.int32)
        # --- the record loop
        keys = _path(1)
        kfile = str(tmp_path / "keys.json")
        keys.save_to_JSON(kfile)
        outdir = str(tmp_path / "movie")
        assert spim_render.main(["-i", tif, "-o", outdir, "-w", "96", "-u", "1", "1", "2", "--keyframes", kfile,
                                 "--frames", "12"]) == 0
        names = sorted(os.listdir(outdir))
        assert names == ["output_%02d.png" % i for i in range(1, 13)]
        lut = np.repeat(np.linspace(0, 1, 256)[:, None], 3, 1)
        r.set_lut(lut)
        loaded = kf.KeyFrameList.load_from_JSON(kfile)
        for pos in (1, 5, 9, 12):
            kf.apply_transform(r, loaded.getTransform(pos / 12.))
            r.render(method="iso_surface" if loaded.getTransform(pos / 12.).isIso else "max_project")
            frame = np.array(Image.open(os.path.join(outdir, "output_%02d.png" % pos)))
            assert frame.shape == (96, 96, 4) and np.array_equal(frame, r.output_rgba()[::-1])
    finally:
        r.close()
    with pytest.raises(ValueError):
        spim_render.main(["-f", "czi", "-i", tif])


# ------------------------------------------------------------------ frames of the record loop sharded over ranks
class _FakeRenderer(object):
    """stands in for VolumeRenderer on the CPU: records the setter calls, render() does nothing"""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name.startswith("set_"):
            return lambda v: self.calls.append((name, v))
        raise AttributeError(name)

    def update_data(self, vol, pinned=False):
        self.calls.append(("update_data", int(vol[0, 0, 0])))

    def set_data(self, vol):
        self.dataImg = type("I", (), {"shape": vol.shape[::-1], "dtype": vol.dtype})()
        self.calls.append(("set_data", int(vol[0, 0, 0])))

    def render(self, method="max_project"):
        self.calls.append(("render", method))


def _shard(rank, world, n_frames=23, n_t=6):
    keys = _path(n_t)
    source = [np.full((2, 2, 2), t, np.uint16) for t in range(n_t)]
    r = _FakeRenderer()
    got = [(pos, td.dataPos, td.isIso) for pos, td, _ in kf.render_keyframes(
        r, keys, n_frames, source=source, pipelined=False, rank=rank, world=world)]
    uploads = [v for name, v in r.calls if name in ("set_data", "update_data")]
    return got, uploads, kf.keyframe_data_positions(keys, n_frames, n_t, rank, world)


def test_record_loop_sharded_over_ranks_covers_every_frame_once():
    keys = _path(6)
    for world in (1, 2, 3, 8):
        shards = [_shard(r, world) for r in range(world)]
        frames = sorted(f for got, _, _ in shards for f in got)
        assert [f[0] for f in frames] == list(range(1, 24))
        for pos, data_pos, iso in frames:
            td = keys.getTransform(pos / 23.)
            assert (data_pos, iso) == (td.dataPos, td.isIso)
        for r, (got, uploads, order) in enumerate(shards):
            assert [g[0] for g in got] == list(range(r + 1, 24, world))
            assert uploads == order                     # a time point is uploaded only when dataPos changes
    with pytest.raises(ValueError):
        next(kf.render_keyframes(_FakeRenderer(), keys, 5, rank=2, world=2))


def test_pipelined_loop_hands_static_stretches_over_as_lists():
    """render_keyframes(pipelined=True): stretches of >= 4 frames in which only the camera moves reach render_sequence as
    a list of modelViews (several frames per launch), everything else as a generator that applies each frame's settings
    when it is pulled; every frame is rendered once, in order, with its own settings in force"""
    class Seq(_FakeRenderer):
        def __init__(self):
            _FakeRenderer.__init__(self)
            self.handed = []

        def render_sequence(self, views, method="max_project", iso_planes=7):
            rec = [isinstance(views, list), 0, method]
            self.handed.append(rec)
            for M in views:          # pulling a generator applies that frame's settings
                self.calls.append(("frame", method))
                rec[1] += 1
                yield self

    q0, q1 = kf.Quaternion(1, 0, 0, 0), kf.Quaternion(0.7, 0, 0.7, 0)
    keys = kf.KeyFrameList()
    keys.addItem(kf.KeyFrame(0., kf.TransformData(quatRot=q0, maxVal=200., gamma=1.)))
    keys.addItem(kf.KeyFrame(.5, kf.TransformData(quatRot=q1, maxVal=200., gamma=1.)))   # only the camera moves
    keys.addItem(kf.KeyFrame(1., kf.TransformData(quatRot=q0, maxVal=100., gamma=.5)))   # the window moves as well
    r = Seq()
    frames = [(pos, td.maxVal) for pos, td, _ in kf.render_keyframes(r, keys, 40)]
    assert [f[0] for f in frames] == list(range(1, 41))
    assert sum(n for _, n, _ in r.handed) == 40 and all(m == "max_project" for _, _, m in r.handed)
    lists = [n for is_list, n, _ in r.handed if is_list]
    assert lists and max(lists) >= 15          # the first half of the path is one static stretch
    gens = [n for is_list, n, _ in r.handed if not is_list]
    assert gens and sum(gens) >= 15            # the second half changes its window every frame
    # settings in force when a frame is rendered: the last set_max_val before the k-th "frame" is frame k's maxVal
    seen, cur = [], None
    for name, v in r.calls:
        if name == "set_max_val":
            cur = v
        elif name == "frame":
            seen.append(cur)
    assert np.allclose(seen, [f[1] for f in frames])


def _kf_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got, uploads, order = _shard(rank, world)
        mine = torch.zeros(24, dtype=torch.int64)
        for pos, _, _ in got:
            mine[pos] += 1
        dist.all_reduce(mine)                            # how often every frame was rendered, over all ranks
        q.put((rank, mine[1:].tolist(), uploads == order))
    finally:
        dist.destroy_process_group()


def test_record_loop_over_a_gloo_group():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_kf_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(counts == [1] * 23 and ok for _, counts, ok in results)


def test_record_writes_the_gui_file_names(tmp_path):
    """keyframe_view.py:644-653: output_<recordPos zero-filled to the digits of nFrames>.png, one per frame; images
    upright (the renderer's row 0 is the bottom row of the view)."""
    from PIL import Image

    class _Fake(_FakeRenderer):
        n = 0

        def set_lut(self, lut):
            self.lut = np.asarray(lut)

        def output_rgba(self, mode_black=True):
            _Fake.n += 1
            img = np.zeros((4, 6, 4), np.uint8)
            img[0, :, 0] = _Fake.n            # marks row 0
            img[..., 3] = 255
            return img

    keys = _path(1)
    lut = np.linspace(0, 1, 8)[:, None].repeat(3, 1)
    r = _Fake()
    names = kf.record_keyframes(r, keys, 12, str(tmp_path / "movie"), lut=lut)
    assert [os.path.basename(n) for n in names] == ["output_%02d.png" % i for i in range(1, 13)]
    assert np.array_equal(r.lut, lut)
    first = np.array(Image.open(names[0]))
    assert first.shape == (4, 6, 4) and first[-1, 0, 0] == 1 and first[0, 0, 0] == 0      # flipped: row 0 at the bottom
    mine = kf.record_keyframes(_Fake(), keys, 12, str(tmp_path / "part"), rank=1, world=3)
    assert [os.path.basename(n) for n in mine] == ["output_%02d.png" % i for i in (2, 5, 8, 11)]


def test_cli_without_a_device(tmp_path, capsys):
    """What spim_render does before it needs the GPU: help without arguments (bin/spim_render.py:96-98), the
    reference's defaults, the camera it builds, the formats it knows."""
    from spimagine_b200.bin import spim_render
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_scale, mat4_translate
    assert spim_render.main([]) == 0
    assert "renders max projections" in capsys.readouterr().out
    a = spim_render.build_parser().parse_args(["-i", "x.tif"])
    assert (a.format, a.output, a.pos, a.width, a.scale, a.units, a.translate, a.rotation, a.range, a.ortho, a.is16Bit) == (
        "tif", "out.png", 0, 400, [1.], [1., 1., 5.], [0, 0, -4], [0, 1, 0, 0], None, False, False)
    a = spim_render.build_parser().parse_args(["-i", "x", "-s", "2", "-r", ".3", "0", "1", "0", "-t", "1", "2", "-5"])
    want = np.dot(mat4_translate(1, 2, -5), np.dot(mat4_rotation(.3, 0, 1, 0), mat4_scale(2, 2, 2)))
    assert np.array_equal(spim_render.model_view(a), want)
    with pytest.raises(ValueError, match="not supported"):
        spim_render.main(["-f", "czi", "-i", str(tmp_path / "x.czi")])
    with pytest.raises(SystemExit):
        spim_render.build_parser().parse_args([])             # -i is required
    out = np.array([[0., .5], [1., 2.]], np.float32)
    assert spim_render.to_uint8(out).tolist() == [[0, 127], [255, 255]]
    assert spim_render.to_uint16(out, 0., 1000.).tolist() == [[0, 250], [500, 1000]]
    assert spim_render.to_uint16(np.ones((2, 2), np.float32), 5., 9.).tolist() == [[5, 5], [5, 5]]
