"""GPU tests of the pipelined sequence API (render_sequence: two output slots, asynchronous read-back) and of
size-independent properties at BASELINE.json's full bench size (512^3 uint16 -> 1024^2)."""
import math

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _renderer(size, **kw):
    from spimagine_b200 import VolumeRenderer
    return VolumeRenderer(size, **kw)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("n_frames", [1, 2, 5])
def test_sequence_equals_frame_by_frame_mip(pinned, n_frames):
    data = scenes.vol_g(48, np.uint16, seed=3)
    cams = [scenes.gui_camera(0.3 + 0.4 * f, 3.5) for f in range(n_frames)]
    rend = _renderer((160, 96), pinned_outputs=pinned)
    rend.set_data(data)
    rend.set_projection(cams[0][1])
    rend.set_max_val(60000.)
    want = []
    for M, _ in cams:
        rend.set_modelView(M)
        rend.render()
        want.append((rend.output.copy(), rend.output_alpha.copy()))
    got = []
    for r in rend.render_sequence([M for M, _ in cams]):
        assert r is rend
        got.append((r.output.copy(), r.output_alpha.copy()))
    assert len(got) == n_frames
    for (o, a), (wo, wa) in zip(got, want):
        assert np.array_equal(o, wo) and np.array_equal(a, wa)
    # the synchronous path still works afterwards and lands in slot 0
    rend.set_modelView(cams[0][0])
    rend.render()
    assert np.array_equal(rend.output, want[0][0])
    rend.close()


def test_sequence_equals_frame_by_frame_iso():
    data = scenes.iso_sphere(48)
    cams = [scenes.gui_camera(0.2 * f, 4.5) for f in range(4)]
    rend = _renderer((128, 128))
    rend.set_data(data)
    rend.set_projection(cams[0][1])
    rend.set_max_val(20.)
    want = []
    for M, _ in cams:
        rend.set_modelView(M)
        rend.render(method="iso_surface")
        want.append([x.copy() for x in (rend.output, rend.output_alpha, rend.output_depth, rend.output_normals,
                                        rend.output_occlusion)])
    k = 0
    for r in rend.render_sequence([M for M, _ in cams], method="iso_surface"):
        for g, w in zip((r.output, r.output_alpha, r.output_depth, r.output_normals, r.output_occlusion), want[k]):
            assert np.array_equal(g, w)
        k += 1
    assert k == len(cams)
    with pytest.raises(KeyError):
        list(rend.render_sequence([cams[0][0]], method="nope"))
    rend.close()


def test_iso_sequence_reads_back_only_rows_that_can_hold_a_surface():
    """render_sequence(method="iso_surface", iso_planes=2): output and alpha of rows the projected box cannot touch are
    not copied (they hold out 0 / alpha 0 for every element type); cameras that move the box across the image, max
    projections and full read-backs in between, float32 and uint16 volumes -- every frame equals the blocking render"""
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate
    P = scenes.gui_camera(0., 4.)[1]
    cams = [np.dot(mat4_translate(0.3 * math.sin(i), 1.4 * math.cos(0.9 * i), -4.5 - 0.5 * (i % 3)), mat4_rotation(0.4 * i, 0., 1., 0.))
            for i in range(9)]
    cams.append(np.dot(mat4_translate(0., 9., -4.), mat4_rotation(0.3, 0., 1., 0.)))   # the box is off screen
    for vol, maxval in ((scenes.iso_sphere(48), 20.), (scenes.vol_g(56, np.uint16, seed=2), 30000.)):
        rend = _renderer((144, 176), pinned_outputs=True)
        rend.set_data(vol)
        rend.set_projection(P)
        rend.set_max_val(maxval)
        try:
            want = []
            for M in cams:
                rend.render(modelView=M, method="iso_surface")
                want.append((rend.output.copy(), rend.output_alpha.copy()))
            assert any(w[0].max() > 0 for w in want) and want[-1][0].max() == 0
            b0 = rend.d2h_bytes()
            got = [(r.output.copy(), r.output_alpha.copy()) for r in
                   rend.render_sequence(cams, method="iso_surface", iso_planes=2)]
            moved = rend.d2h_bytes() - b0
            assert 0 < moved < 0.9 * len(cams) * 2 * 144 * 176 * 4        # rows were left out ...
            for i, ((o, a), (wo, wa)) in enumerate(zip(got, want)):        # ... and the frames read the same
                assert np.array_equal(o, wo) and np.array_equal(a, wa), i
            # a max projection and a full iso read-back dirty the staging in between
            rend.render(modelView=cams[2], method="max_project")
            full = [r.output_depth.copy() for r in rend.render_sequence(cams[:3], method="iso_surface")]
            assert np.isfinite(full[0]).any()
            got = [(r.output.copy(), r.output_alpha.copy()) for r in
                   rend.render_sequence(cams[::-1], method="iso_surface", iso_planes=2)]
            for i, ((o, a), (wo, wa)) in enumerate(zip(got, want[::-1])):
                assert np.array_equal(o, wo) and np.array_equal(a, wa), i
        finally:
            rend.close()


def test_iso_sequence_on_a_camera_path_equals_unclipped_reads():
    """A slowly turning, drifting camera: render_sequence(iso_planes=2) moves each frame's rectangle through the device
    staging and, where the previous frame's rectangle sticks out by little, copies the union of the two instead of
    clearing the pinned planes on the host.  Every frame must equal a blocking render whose planes are read back
    whole (tuning knob 9 = 0), for both settings of the staging knob (18)."""
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate
    P = scenes.gui_camera(0., 4.)[1]
    cams = [np.dot(mat4_translate(0.02 * i - 0.3, 0.25 * math.sin(0.2 * i), -4.2 + 0.01 * i), mat4_rotation(0.05 * i, 0.3, 1., 0.))
            for i in range(24)]
    rend = _renderer((208, 176), pinned_outputs=True)
    try:
        rend.set_data(scenes.vol_g(56, np.uint16, seed=2))
        rend.set_projection(P)
        rend.set_max_val(30000.)
        rend._check(rend._lib.spv_set_tuning(rend._ctx, 9, 0))
        want = []
        for M in cams:
            rend.render(modelView=M, method="iso_surface")
            want.append((rend.output.copy(), rend.output_alpha.copy()))
        rend._check(rend._lib.spv_set_tuning(rend._ctx, 9, 1))
        assert sum(w[0].max() > 0 for w in want) == len(cams)
        for stage in (1, 0, 1):
            rend._check(rend._lib.spv_set_tuning(rend._ctx, 18, stage))
            b0 = rend.d2h_bytes()
            got = [(r.output.copy(), r.output_alpha.copy()) for r in rend.render_sequence(cams, method="iso_surface", iso_planes=2)]
            assert rend.d2h_bytes() - b0 < 0.9 * len(cams) * 2 * 208 * 176 * 4
            for i, ((o, a), (wo, wa)) in enumerate(zip(got, want)):
                assert np.array_equal(o, wo) and np.array_equal(a, wa), (stage, i)
    finally:
        rend.close()


def test_sequence_abandoned_midway_leaves_renderer_usable():
    data = scenes.vol_g(32, np.uint16, seed=1)
    cams = [scenes.gui_camera(0.1 * f, 3.5) for f in range(6)]
    rend = _renderer((64, 64))
    rend.set_data(data)
    rend.set_projection(cams[0][1])
    rend.set_max_val(60000.)
    it = rend.render_sequence([M for M, _ in cams])
    next(it)
    it.close()  # generator closed with frames in flight
    rend.set_modelView(cams[2][0])
    rend.render()
    a = rend.output.copy()
    rend.resize((80, 48))
    rend.set_modelView(cams[2][0])
    rend.render()
    assert rend.output.shape == (48, 80)
    rend.resize((64, 64))
    rend.set_modelView(cams[2][0])
    rend.render()
    assert np.array_equal(rend.output, a)
    rend.close()


# ----------------------------------------------------------------------------- full bench size (configs[1])
@pytest.fixture(scope="module")
def full_c2():
    vol = scenes.vol_g(512, np.uint16, seed=0)
    rend = _renderer((1024, 1024))
    rend.set_data(vol)
    rend.set_max_val(60000.)
    M, P = scenes.gui_camera(2 * math.pi * 40 / 360, 4.0)
    rend.set_projection(P)
    rend.set_modelView(M)
    yield vol, rend, M, P
    rend.close()


def test_full_size_properties(full_c2):
    vol, rend, M, P = full_c2
    assert rend.data_min_max == (float(vol.min()), float(vol.max()))
    rend.render()
    auto = rend.output.copy()
    assert rend.mip_axis_used() == (1, 1)
    # multi-pass renders and empty-space skipping run mip_fast_kernel on the z copy: bit-level identities are stated
    # against the plain render through that copy; the default render (pairs along y for this camera) is the same image
    # up to the texture unit's weight rounding
    rend.set_view_copies("primary")
    rend.render()
    img, alpha = rend.output.copy(), rend.output_alpha.copy()
    assert np.abs(img - auto).max() < 2e-4
    hit = alpha > 0  # uint16 path: alpha = tnear on hit, 0 on miss (camera outside the box)
    assert 0.25 < hit.mean() < 0.45
    assert np.all(img[~hit] == 0) and img.min() >= 0 and img.max() <= 1
    # idempotence
    rend.render()
    assert np.array_equal(rend.output, img)
    # window linearity for alpha_pow == 0: the raw maximum does not depend on the window
    rend.render(maxVal=30000.)
    assert np.allclose(np.minimum(2 * img, 1), rend.output, atol=2e-7)
    rend.set_max_val(60000.)
    # multi-pass rendering covers more sample positions: fmax-merged result never drops below a single part
    rend.render(numParts=2, currentPart=0)
    p0 = rend.output.copy()
    rend.render(numParts=2, currentPart=1)
    assert np.all(rend.output >= p0)
    # empty-space skipping and the plain 3-D layout agree with the default path
    rend.set_skipping(True)
    rend.render()
    assert np.array_equal(rend.output, img)
    rend.set_skipping(None)
    rend.set_view_copies("auto")


def test_full_size_slabs_are_bit_exact(full_c2):
    """encode -> split -> composite round trip at the full bench size: 4 z-slabs rendered separately and merged
    with max equal the monolithic render bit for bit."""
    import ctypes as C
    from spimagine_b200 import _lib
    from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs, slab_with_halo
    vol, rend, M, P = full_c2
    rend.set_view_copies("primary")  # the slabs hold pairs along z: bit for bit against the render through the z copy
    rend.render()
    rend.set_view_copies("auto")
    want = rend.output.copy()
    acc = None
    for rank, (z0, z1) in enumerate(partition_slabs(512, 4)):
        s = SlabMaxProjector((1024, 1024), rank=rank, world=4)
        lo, hi = slab_with_halo(z0, z1, 512)
        s.set_slab(vol[lo:hi], 512, z0, z1)
        s.set_projection(P)
        s.set_modelView(M)
        p = _lib.MipParams(s._box(), 0., 60000., 1., 0., 1, 0, 200, _lib.MIP_RAW_ONLY)
        s._check(s._lib.spv_render_mip(s._ctx, C.byref(p)))
        raw = np.empty((1024, 1024), np.float32)
        s._check(s._lib.spv_read(s._ctx, _lib.BUF_RAW, _lib.fp(raw), raw.size))
        acc = raw if acc is None else np.maximum(acc, raw)
        s.close()
    got = np.where(acc < 0, 0, np.clip(acc / np.float32(60000.), 0, 1)).astype(np.float32)
    assert np.array_equal(got, want)


def test_full_size_against_oracle_rows(full_c2, oracle_mod):
    """The oracle on every 64th row of the full-size frame (seconds on the CPU): 1e-3 of the range."""
    vol, rend, M, P = full_c2
    rend.render()
    o = oracle_mod.OracleRenderer((1024, 1024), kind="port")
    o.set_data(vol)
    o.set_modelView(M)
    o.set_projection(P)
    o.lib.so_set_row_sampling(0, 64)
    o.render(maxVal=60000.)
    o.lib.so_set_row_sampling(0, 1)
    rows = slice(0, 1024, 64)
    assert np.abs(o.output[rows] - rend.output[rows]).max() < 1e-3
    assert np.array_equal(o.output_alpha[rows], rend.output_alpha[rows])


# ----------------------------------------------------------------------------- ingest paths
@pytest.mark.parametrize("dtype,layout", [(np.uint16, "zpair"), (np.uint16, "3d"), (np.float32, "3d"), (np.uint8, "zpair")])
def test_ingest_paths_agree(dtype, layout):
    """pageable (threaded staging ring), page-locked (direct DMA, asynchronous) and device sources fill the volume
    identically; volumes larger than one 32 MiB chunk, so the two-deep pipeline wraps around several times."""
    import torch
    from spimagine_b200 import pinned_empty
    shape = (45, 512, 1024) if np.dtype(dtype).itemsize < 4 else (37, 512, 512)
    if np.dtype(dtype) == np.uint8:
        shape = (150, 512, 1024)
    a = scenes.random_vol(shape, dtype, seed=11)
    b = np.ascontiguousarray(a[::-1, ::-1])
    M, P = scenes.gui_camera(0.7, 3.6)
    rend = _renderer((192, 160))
    rend.set_layout(layout)
    rend.set_projection(P)
    rend.set_modelView(M)
    peak = float(a.max())

    def image():
        rend.render(maxVal=peak)
        return rend.output.copy()

    rend.set_data(a)                      # pageable, allocating
    img_a = image()
    assert rend.data_min_max == (float(a.min()), float(a.max()))
    rend.update_data(b)                   # pageable, same shape
    img_b = image()
    assert not np.array_equal(img_a, img_b)
    pa, pb = pinned_empty(shape, dtype), pinned_empty(shape, dtype)
    pa[...] = a
    pb[...] = b
    for k in range(3):                    # asynchronous uploads back to back, renders in between
        rend.update_data(pa, pinned=True)
        assert np.array_equal(image(), img_a)
        rend.update_data(pb, pinned=True)
        assert np.array_equal(image(), img_b)
    da = torch.from_numpy(a.view(np.int16) if np.dtype(dtype) == np.uint16 else a).cuda()
    rend.set_data_device(da.data_ptr(), shape, dtype)
    assert np.array_equal(image(), img_a)
    # point samples straight from the array: every voxel arrived where it belongs
    rng = np.random.default_rng(0)
    idx = np.stack([rng.integers(0, s, 4000) for s in shape], 1)  # (z, y, x)
    pos = ((idx[:, ::-1] + 0.5) / np.array(shape[::-1])).astype(np.float32)
    rend2 = _renderer((64, 64), interpolation="nearest", sampler="exact")
    rend2.set_layout(layout)
    rend2.set_data(a)
    assert np.array_equal(rend2.sample_points(pos), a[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32))
    rend2.update_data(pb, pinned=True)
    assert np.array_equal(rend2.sample_points(pos), b[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32))
    rend.close()
    rend2.close()


@pytest.mark.parametrize("target,layout", [(np.float32, "3d"), (np.uint16, "zpair"), (np.uint16, "3d"), (np.uint8, "zpair")])
def test_device_conversion_equals_host_astype(target, layout):
    """set_data / update_data of an array whose element type is not a texel type: the reference converts on the host
    (`astype(self.dtype)`, volumerender.py:245-246, 290-291); here the source bytes travel and are converted on the
    device.  Every voxel must come out as numpy's astype would have made it -- integers wrap, floats round to
    nearest (float target) or truncate (integer targets) -- for sources larger than one pipeline chunk."""
    from spimagine_b200 import pinned_empty
    rng = np.random.default_rng(5)
    shape = (45, 384, 512)
    hi = 250 if np.dtype(target) == np.uint8 else 60000
    sources = {
        np.int8: rng.integers(-128, 128, shape).astype(np.int8),
        np.int16: rng.integers(-32768, 32768, shape).astype(np.int16),
        np.int32: rng.integers(-70000, 200000, shape).astype(np.int32),
        np.uint32: rng.integers(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32),
        np.int64: rng.integers(-2 ** 40, 2 ** 40, shape).astype(np.int64),
        np.uint64: rng.integers(0, 2 ** 63, shape, dtype=np.uint64),
        np.float16: (rng.random(shape) * min(hi, 2000)).astype(np.float16),
        np.float64: rng.random(shape) * hi,
        np.bool_: rng.random(shape) > .5,
    }
    if np.dtype(target) == np.float32:   # a float source with more mantissa than the texel: rounding matters
        sources[np.float64] = rng.random(shape) * 1e6 - 5e5
    idx = np.stack([rng.integers(0, s, 6000) for s in shape], 1)  # (z, y, x)
    pos = ((idx[:, ::-1] + 0.5) / np.array(shape[::-1])).astype(np.float32)
    rend = _renderer((64, 64), interpolation="nearest", sampler="exact")
    rend.set_layout(layout)
    rend.set_dtype(target)
    first = True
    for st, a in sources.items():
        want = a.astype(target)
        if first:
            rend.set_data(a)                       # allocating path: spv_set_volume_from
            first = False
        else:
            rend.update_data(a)                    # spv_update_volume_from
        assert rend.dtype == target and rend.dataImg.dtype == np.dtype(target)
        got = rend.sample_points(pos)
        assert np.array_equal(got, want[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32)), st
        assert rend.data_min_max == (float(want.min()), float(want.max())), st
    # a page-locked source of another element type is converted as well (synchronously)
    p = pinned_empty(shape, np.int16)
    p[...] = sources[np.int16]
    rend.update_data(p, pinned=True)
    want = sources[np.int16].astype(target)
    assert np.array_equal(rend.sample_points(pos), want[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32))
    # unsupported element types still raise when autoConvert is off, like the reference
    with pytest.raises(NotImplementedError):
        rend.set_data(sources[np.int32], autoConvert=False)
    rend.close()


# ----------------------------------------------------------------------------- timelapse playback
def test_timelapse_player_resident_and_streamed():
    """Frame sharding for 3D+t playback: every rank's player shows exactly its own time points, resident and
    streamed (pageable and page-locked sources) playback equal a plain per-frame render."""
    from spimagine_b200 import pinned_empty
    from spimagine_b200.multigpu import TimelapsePlayer
    T, shape = 7, (24, 40, 48)
    source = [scenes.vol_g(0, np.uint16, seed=100, shape=shape, t=t) for t in range(T)]
    cams = [scenes.gui_camera(0.2 * t, 3.6) for t in range(T)]
    P = cams[0][1]
    want = []
    ref = _renderer((96, 80))
    ref.set_view_copies("primary")  # as the player's renderers: time points are rendered through the z copy
    for t in range(T):
        ref.set_data(source[t])
        ref.set_projection(P)
        ref.set_modelView(cams[t][0])
        ref.render(maxVal=60000.)
        want.append(ref.output.copy())
    ref.close()
    seen = []
    for rank in range(3):
        pl = TimelapsePlayer((96, 80), rank=rank, world=3)
        mine = pl.my_frames(T)
        assert mine == list(range(rank, T, 3))
        assert pl.preload(source) == mine
        for t in mine:
            r = pl.render_resident(t, cams[t][0], projection=P, max_val=60000.)
            assert np.array_equal(r.output, want[t])
        got = {t: r.output.copy() for t, r in pl.play(source, [c[0] for c in cams], projection=P, max_val=60000.)}
        assert sorted(got) == mine and all(np.array_equal(got[t], want[t]) for t in mine)
        pinned = []
        for t in range(T):
            a = pinned_empty(shape, np.uint16)
            a[...] = source[t]
            pinned.append(a)
        got = {t: r.output.copy() for t, r in pl.play(pinned, [c[0] for c in cams], pinned=True, projection=P, max_val=60000.)}
        assert all(np.array_equal(got[t], want[t]) for t in mine)
        seen += mine
        pl.close()
    assert sorted(seen) == list(range(T))


# ----------------------------------------------------------------------------- display hand-off
@pytest.mark.parametrize("dtype,n_lut", [(np.float32, 256), (np.uint16, 37)])
def test_display_pass_matches_the_shader_restatement(dtype, n_lut, oracle_mod):
    """output_rgba(): LUT colour + alpha computed on the device (texture.frag:8-38) equals the numpy restatement
    applied to the float planes byte for byte, in black and white mode, after max projection and iso surface; the
    float32 kernel's misses (alpha = -1) come out transparent."""
    rng = np.random.default_rng(3)
    lut = rng.random((n_lut, 3)).astype(np.float32)
    lut[0] = 0
    data = scenes.vol_g(64, dtype, seed=2)
    peak = float(data.max())
    rend = _renderer((200, 144))
    rend.set_data(data)
    M, P = scenes.gui_camera(0.5, 3.4)
    rend.set_projection(P)
    rend.set_modelView(M)
    rend.set_lut(lut)
    for method, maxVal in (("max_project", .6 * peak), ("iso_surface", .5 * peak)):
        rend.render(maxVal=maxVal, method=method)
        for black in (True, False):
            got = rend.output_rgba(mode_black=black)
            want = oracle_mod.display_rgba8(rend.output, rend.output_alpha, lut, mode_black=black)
            assert got.shape == (144, 200, 4) and got.dtype == np.uint8
            assert np.array_equal(got, want)
        if method == "max_project":
            miss = rend.output_alpha < 0 if np.dtype(dtype) == np.float32 else rend.output_alpha == 0
            assert 0 < miss.sum() < miss.size
            if np.dtype(dtype) == np.float32:
                assert not got[miss].any()
            assert got[..., 3].max() == 255          # values above maxVal saturate
    # the device-only path feeds the display pass without the float read-back
    rend.render(maxVal=.6 * peak)
    want = rend.output_rgba()
    rend.resize((96, 80))
    rend.resize((200, 144))
    rend.set_modelView(M)
    rend.render_device_only()
    assert np.array_equal(rend.output_rgba(), want)
    with pytest.raises(Exception):
        _renderer((32, 32)).output_rgba()            # no colour map set
    rend.close()


@pytest.mark.parametrize("size", [(256, 200), (1024, 1024)])
def test_read_back_row_orders_and_band_counts_give_the_same_frame(size):
    """spv_render_mip_to_host deals its tile rows in an order chosen for the read-back overlap (tuning knob 8) and cuts
    the frame into bands (knob 2): scheduling only -- every combination returns the planes of the plain render, for
    boxes in the middle of the image, off centre, filling it, partly behind the eye and off screen."""
    import ctypes as C
    from spimagine_b200 import _lib
    from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_rotation, mat4_translate
    data = scenes.vol_g(64, np.uint16, seed=5)
    rend = _renderer(size)
    rend.set_data(data)
    rend.set_max_val(60000.)
    P = mat4_perspective(60, 1., .1, 10)
    views = [np.dot(mat4_translate(0, 0, -4), mat4_rotation(0.4, 0, 1, 0)),          # centred
             np.dot(mat4_translate(0.3, 1.4, -4), mat4_rotation(1.1, 1, 1, 0)),     # near the top edge, partly outside
             np.dot(mat4_translate(0, 0, -1.6), mat4_rotation(0.2, 0, 1, 0)),       # fills the image
             np.dot(mat4_translate(0, 0, -0.5), mat4_rotation(0.2, 0, 1, 0)),       # corners behind the eye
             np.dot(mat4_translate(0, 9., -4), mat4_rotation(0.2, 0, 1, 0))]        # off screen
    rend.set_projection(P)
    n = size[0] * size[1]
    for M in views:
        rend.set_modelView(M)
        rend.render_device_only()
        rend.sync()
        want, _ = rend._fetch(2)
        want = want.copy()
        for mode in (0, 1):
            for bands in (1, 5, 12, 32):
                rend._check(rend._lib.spv_set_tuning(rend._ctx, 8, mode))
                rend._check(rend._lib.spv_set_tuning(rend._ctx, 2, bands))
                rend.render()
                got = np.concatenate([rend.output.ravel(), rend.output_alpha.ravel()])
                assert got.size == 2 * n and np.array_equal(got, want), (mode, bands)
    rend.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_rows_that_cannot_hold_hits_are_not_copied_but_read_the_same(dtype):
    """spv_render_mip_to_host leaves out the rows the box cannot project to (tuning knob 9): the pinned staging holds
    their miss values already.  Every frame must equal the frame with all rows copied -- across cameras that move the
    box around the image, box boundaries, both projections, interleaved iso-surface frames and other read paths
    that scribble over the same staging, both output slots, and a change of the volume's element type."""
    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.utils.transform_matrices import mat4_ortho, mat4_translate, mat4_rotation, mat4_perspective
    vol = scenes.vol_g(48, dtype, seed=4)
    peak = float(vol.max())
    a = VolumeRenderer((256, 384), pinned_outputs=True)      # clipped copies (default)
    b = VolumeRenderer((256, 384))
    try:
        b._check(b._lib.spv_set_tuning(b._ctx, 9, 0))         # every row copied
        for r in (a, b):
            r.set_data(vol)
            r.set_max_val(peak)
        rng = np.random.default_rng(0)
        cams = []
        for i in range(40):
            M = np.dot(mat4_translate(rng.uniform(-1.5, 1.5), rng.uniform(-2.5, 2.5), rng.uniform(-9, -2.2)),
                       mat4_rotation(rng.uniform(0, 6.3), *rng.normal(size=3)))
            cams.append((M, mat4_perspective(60, 1., .1, 20) if i % 5 else mat4_ortho(-2, 2, -3, 3, -10, 10)))
        cams.append((mat4_translate(0, 0, -.5), mat4_perspective(60, 1., .1, 20)))      # inside the volume
        cams.append((mat4_translate(0, 40., -4), mat4_perspective(60, 1., .1, 20)))     # box off screen
        n_clipped = 0
        for i, (M, P) in enumerate(cams):
            box = [-1, 1, -1, 1, -1, 1] if i % 3 else [-.5, .8, -.3, .4, -1, 1]
            for r in (a, b):
                r.set_projection(P)
                r.set_modelView(M)
                r.set_box_boundaries(box)
            if i % 7 == 3:                                     # other writers of the same staging in between
                a.render(method="iso_surface", maxVal=peak)
                b.render(method="iso_surface", maxVal=peak)
                assert np.array_equal(a.output, b.output) and np.array_equal(a.output_alpha, b.output_alpha), i
                if i % 2:
                    assert np.array_equal(a.output_depth, b.output_depth)
                    assert np.array_equal(a.output, b.output) and np.array_equal(a.output_alpha, b.output_alpha), i
                a.render(method="iso_surface_raw", maxVal=peak)
                b.render(method="iso_surface_raw", maxVal=peak)
                assert np.array_equal(a.output, b.output) and np.array_equal(a.output_alpha, b.output_alpha), i
            if i % 11 == 5:
                for _ in a.render_sequence([M, M, M], method="iso_surface" if i % 2 else "max_project"):
                    pass
            a.render()
            b.render()
            assert np.array_equal(a.output, b.output), i
            assert np.array_equal(a.output_alpha, b.output_alpha), i
            assert not a.output.flags.writeable
            n_clipped += int(not (a.output_alpha[0] != a.output_alpha[0, 0]).any())
        assert n_clipped > 10
        seq = [M for M, _ in cams[:12]]
        a.set_box_boundaries([-1, 1, -1, 1, -1, 1])
        b.set_box_boundaries([-1, 1, -1, 1, -1, 1])
        got = [(np.array(r.output), np.array(r.output_alpha)) for r in a.render_sequence(seq)]
        for M, (o, al) in zip(seq, got):
            b.set_modelView(M)
            b.render()
            assert np.array_equal(o, b.output) and np.array_equal(al, b.output_alpha)
        # another element type: the alpha plane's miss value changes (0 <-> -1)
        other = scenes.vol_g(48, np.float32 if dtype == np.uint16 else np.uint16, seed=5)
        for r in (a, b):
            r.set_data(other)
            r.set_max_val(float(other.max()))
            r.set_modelView(cams[1][0])
            r.render()
        assert np.array_equal(a.output, b.output) and np.array_equal(a.output_alpha, b.output_alpha)
    finally:
        a.close()
        b.close()


def test_iso_sequence_with_a_new_volume_between_frames():
    """render_sequence(method="iso_surface") runs a frame's screen-space passes and every other frame's search on
    streams of their own (tuning knob 14); an update_data issued from the generator between two frames must not reach
    the array while a search on the second stream is still reading it, and a max projection right after the sequence
    must see finished slots.  Every frame equals the frame-by-frame render."""
    vols = [scenes.vol_g(0, np.uint16, seed=s, shape=(72, 80, 88)) for s in (1, 2, 3)]
    cams = [scenes.gui_camera(0.25 * f, 3.2) for f in range(9)]
    rend = _renderer((192, 160))
    rend.set_data(vols[0])
    rend.set_projection(cams[0][1])
    rend.set_max_val(24000.)
    want = []
    for f, (M, _) in enumerate(cams):
        if f % 2 == 0:
            rend.update_data(vols[(f // 2) % 3])
        rend.set_modelView(M)
        rend.render(method="iso_surface")
        want.append([x.copy() for x in (rend.output, rend.output_alpha, rend.output_depth, rend.output_normals,
                                        rend.output_occlusion)])

    def views():
        for f, (M, _) in enumerate(cams):
            if f % 2 == 0:
                rend.update_data(vols[(f // 2) % 3])
            yield M

    for rep in range(2):
        k = 0
        for r in rend.render_sequence(views(), method="iso_surface"):
            got = (r.output, r.output_alpha, r.output_depth, r.output_normals, r.output_occlusion)
            for g, w in zip(got, want[k]):
                assert np.array_equal(g, w), (rep, k)
            k += 1
        assert k == len(cams)
        rend.update_data(vols[0])
        rend.set_modelView(cams[0][0])
        rend.render()                       # a max projection into slot 0 right behind the sequence
        mip = rend.output.copy()
        rend.render(method="iso_surface")
        assert np.array_equal(rend.output, want[0][0]) and np.array_equal(rend.output_depth, want[0][2])
        rend.render()
        assert np.array_equal(rend.output, mip)
    rend.close()
