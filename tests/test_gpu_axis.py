"""View-aligned layered copies and several frames per launch (csrc/spv_mip_axis.cu, spv_render_mip_batch):
max_project_short of volume_kernel.cl:185-335 through pairs along x, y or z with three lane-to-pixel maps.

Every combination is compared with the CPU oracle (north_star: within 1e-3 of the dynamic range per pixel; the alpha plane
-- tnear / hit mask -- bit for bit), the z copy with 2x2 quads with mip_fast_kernel bit for bit, and launches of several
frames with the same frames rendered one by one bit for bit.  "Oracle": see tests/test_gpu_configs.py."""
import ctypes as C
import math

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _renderer(size, **kw):
    from spimagine_b200 import VolumeRenderer
    return VolumeRenderer(size, **kw)


def _knob(rend, k, v):
    assert rend._lib.spv_set_tuning(rend._ctx, k, v) == 0


def _oracle_image(oracle_mod, vol, size, M, P, **kw):
    o = oracle_mod.OracleRenderer(size, kind="port")
    o.set_data(vol)
    o.set_modelView(M)
    o.set_projection(P)
    for k, v in kw.items():
        getattr(o, "set_" + k)(v)
    o.render()
    return o.output.copy(), o.output_alpha.copy()


CAMS = [("gui 0.3", lambda: scenes.gui_camera(0.3, 3.4)), ("gui 1.9", lambda: scenes.gui_camera(1.9, 3.0)),
        ("tilted", lambda: scenes.tilted_camera(3.2))]


# The texture unit's 8-bit filter weights cost up to (steepest voxel-to-voxel step) / 512: within north_star's 1e-3 of the
# range on volumes whose features span a few voxels (asserted on 96^3 here and on configs[1] at its own size below), up to
# 4e-3 on the small non-cubic volumes whose blobs are one or two voxels wide -- the bound tests/test_gpu_parity.py states
# for mip_fast_kernel on its 24^3..32^3 scenes.
@pytest.mark.parametrize("dtype,shape,tol", [(np.uint16, (40, 56, 72), 4e-3), (np.uint8, (64, 33, 47), 4e-3),
                                             (np.uint16, (96, 96, 96), 1e-3), (np.float32, (44, 60, 52), 4e-3),
                                             (np.float32, (96, 96, 96), 1e-3)])
def test_every_axis_and_lane_map_against_the_oracle(oracle_mod, dtype, shape, tol):
    """non-cubic volumes (the three copies have three different extents), every forced (axis, lane map), three cameras"""
    vol = scenes.vol_g(max(shape), dtype, seed=5, shape=shape)
    peak = float(vol.max())
    size = (208, 152)
    rend = _renderer(size)
    rend.set_data(vol)
    try:
        for name, cam in CAMS:
            M, P = cam()
            ref, ref_a = _oracle_image(oracle_mod, vol, size, M, P, max_val=peak)
            rend.set_projection(P)
            rend.set_modelView(M)
            _knob(rend, 16, 0)
            rend.render(maxVal=peak)
            assert rend.mip_axis_used() == (-1, -1)
            fast = rend.output.copy()
            assert np.abs(fast - ref).max() < tol
            for axis in range(3):
                for quad in range(3):
                    _knob(rend, 16, 10 + 3 * axis + quad)
                    rend.render(maxVal=peak)
                    assert rend.mip_axis_used() == (axis, quad)
                    err = float(np.abs(rend.output - ref).max())
                    assert err < tol, (name, axis, quad, err)
                    assert np.array_equal(rend.output_alpha, ref_a), (name, axis, quad)
                    if axis == 2 and np.dtype(dtype) != np.float32:  # the lane map only changes which thread renders a pixel
                        assert np.array_equal(rend.output, fast), (name, quad)  # (float32: `fast` is the 3-D trilinear fetch)
            assert (ref_a > 0).mean() > 0.05
    finally:
        rend.close()


def test_window_gamma_box_and_ragged_image_sizes(oracle_mod):
    """windows, gamma, reduced boxes; image widths that are not multiples of 4 / 16 (per-pixel stores at the edge)"""
    vol = scenes.vol_g(48, np.uint16, seed=3)
    M, P = scenes.gui_camera(0.8, 3.1)
    for size in [(203, 149), (64, 40), (17, 9)]:
        rend = _renderer(size)
        rend.set_data(vol)
        rend.set_projection(P)
        rend.set_modelView(M)
        try:
            for kw in [dict(max_val=40000., min_val=3000., gamma=0.7), dict(max_val=60000., box_boundaries=[-.4, .7, -.9, .2, -.5, .5])]:
                ref, ref_a = _oracle_image(oracle_mod, vol, size, M, P, **kw)
                rend.set_max_val(kw["max_val"])
                rend.set_min_val(kw.get("min_val", 0.))
                rend.set_gamma(kw.get("gamma", 1.))
                rend.set_box_boundaries(kw.get("box_boundaries", [-1, 1, -1, 1, -1, 1]))
                for mode in (1, 10 + 3 * 1 + 1, 10 + 3 * 0 + 2, 10 + 3 * 2 + 0):
                    _knob(rend, 16, mode)
                    rend.render()
                    assert rend.mip_axis_used()[0] >= 0
                    assert np.abs(rend.output - ref).max() < 4e-3, (size, kw, mode)  # 48^3: see above
                    assert np.array_equal(rend.output_alpha, ref_a), (size, kw, mode)
        finally:
            rend.close()


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
def test_attenuated_projection_through_every_copy(oracle_mod, dtype):
    """alpha_pow != 0 (volume_kernel.cl:300-318) through the layered copies: sub-critical attenuation against the oracle,
    the z copy against mip_alpha_kernel bit for bit"""
    vol = scenes.vol_g(96, dtype, seed=8)
    peak = float(vol.max())
    size = (176, 144)
    M, P = scenes.gui_camera(0.7, 3.3)
    rend = _renderer(size)
    rend.set_data(vol)
    rend.set_projection(P)
    rend.set_modelView(M)
    try:
        for alpha_pow, gamma in [(0.3, 1.), (1.0, 0.7)]:
            ref, ref_a = _oracle_image(oracle_mod, vol, size, M, P, max_val=peak, alpha_pow=alpha_pow, gamma=gamma)
            rend.set_alpha_pow(alpha_pow)
            rend.set_gamma(gamma)
            _knob(rend, 16, 0)
            rend.render(maxVal=peak)
            fast = rend.output.copy()
            assert rend.mip_axis_used() == (-1, -1) and np.abs(fast - ref).max() < 1e-3
            for axis in range(3):
                for quad in range(3):
                    _knob(rend, 16, 10 + 3 * axis + quad)
                    rend.render()
                    assert rend.mip_axis_used() == (axis, quad)
                    assert np.abs(rend.output - ref).max() < 1e-3, (alpha_pow, axis, quad)
                    assert np.array_equal(rend.output_alpha, ref_a)
                    if axis == 2:
                        assert np.array_equal(rend.output, fast), (alpha_pow, quad)
            assert ref.max() > 0.05
    finally:
        rend.close()


def test_choice_follows_the_camera():
    """the GUI camera spins about y: pairs along y, row quads, at every angle, from the first frame after an upload on
    (the choice is a function of the camera: a view renders to the same bits whatever came before); seen along y (from
    above): not the y copy; "primary": only the lane map is chosen"""
    vol = scenes.vol_g(64, np.uint16, seed=1)
    rend = _renderer((256, 256))
    rend.set_data(vol)
    try:
        M, P = scenes.gui_camera(0.4, 4.0)
        rend.set_projection(P)
        for th in (0.4, 0.0, 1.2, 2.0, 3.5, 5.5):
            rend.set_modelView(scenes.gui_camera(th, 4.0)[0])
            rend.render(maxVal=60000.)
            assert rend.mip_axis_used() == (1, 1), th
        first = rend.output.copy()
        from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate
        top = np.dot(mat4_translate(0, 0, -4.), mat4_rotation(math.pi / 2, 1., 0., 0.))  # looking along y
        rend.set_modelView(top)
        rend.render()
        assert rend.mip_axis_used()[0] != 1
        rend.update_data(vol)
        rend.set_modelView(scenes.gui_camera(5.5, 4.0)[0])
        rend.render()
        assert rend.mip_axis_used() == (1, 1)
        assert np.array_equal(rend.output, first)
        rend.set_view_copies("primary")
        rend.render()
        assert rend.mip_axis_used()[0] == 2
        rend.set_view_copies("off")
        rend.render()
        assert rend.mip_axis_used() == (-1, -1)
        with pytest.raises(KeyError):
            rend.set_view_copies("nope")
    finally:
        rend.close()


def test_update_data_rebuilds_the_copies():
    a = scenes.random_vol((50, 60, 70), np.uint16, seed=1)
    b = scenes.random_vol((50, 60, 70), np.uint16, seed=2)
    M, P = scenes.gui_camera(0.9, 3.3)
    size = (160, 128)
    rend = _renderer(size)
    rend.set_projection(P)
    rend.set_modelView(M)
    try:
        _knob(rend, 16, 10 + 3 * 1 + 1)
        rend.set_data(a)
        rend.render(maxVal=65535.)
        img_a = rend.output.copy()
        rend.update_data(b)
        rend.render()
        img_b = rend.output.copy()
        assert not np.array_equal(img_a, img_b)
        _knob(rend, 16, 0)
        rend.render()
        assert np.abs(rend.output - img_b).max() < 8e-3   # two texture-unit paths on white noise: twice the weight error
        _knob(rend, 16, 10 + 3 * 0 + 2)
        rend.update_data(a)
        rend.render()
        assert np.abs(rend.output - img_a).max() < 8e-3
    finally:
        rend.close()


def test_frames_of_a_launch_equal_single_frames():
    """launches of 1..16 frames with cameras that pick different copies and lane maps, to the host (only the rectangle
    the projected box can touch travels) and on the device; rectangles that shrink and move between launches into the
    same set of planes; a box that is partly and wholly off screen"""
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate
    vol = scenes.vol_g(72, np.uint16, seed=4)
    size = (240, 176)
    rend = _renderer(size)
    rend.set_data(vol)
    rend.set_max_val(60000.)
    P = scenes.gui_camera(0., 4.)[1]
    rend.set_projection(P)
    cams = [scenes.gui_camera(0.37 * i, 2.6 + 0.35 * (i % 5))[0] for i in range(16)]
    cams[3] = scenes.tilted_camera(3.2)[0]
    cams[5] = np.dot(mat4_translate(0, 0, -4.), mat4_rotation(math.pi / 2, 1., 0., 0.))
    cams[7] = np.dot(mat4_translate(1.6, 0.3, -4.), mat4_rotation(0.5, 0., 1., 0.))   # partly off screen
    cams[9] = np.dot(mat4_translate(9., 0., -4.), mat4_rotation(0.5, 0., 1., 0.))     # wholly off screen
    cams[11] = np.dot(mat4_translate(0, 0, -0.5), mat4_rotation(0.2, 0., 1., 0.))     # the eye inside the box
    try:
        singles = []
        for M in cams:
            rend.render(modelView=M)
            singles.append((rend.output.copy(), rend.output_alpha.copy(), rend.mip_axis_used()))
        assert len(set(s[2] for s in singles)) >= 2   # the cameras do not all pick the same copy / lane map
        assert singles[9][1].max() == 0 and singles[0][1].max() > 0
        for order in ([16], [5, 16, 1, 7, 16, 3, 2, 16]):
            for n in order:
                which = rend.render_batch(cams[:n] if n != 7 else cams[9:16], True)
                frames = rend.batch_frames_of(which, copy=True)
                want = singles[:n] if n != 7 else singles[9:16]
                assert len(frames) == len(want)
                for f, ((o, a), (so, sa, _)) in enumerate(zip(frames, want)):
                    assert np.array_equal(o, so), (n, f)
                    assert np.array_equal(a, sa), (n, f)
        # device planes
        which = rend.render_batch(cams[4:12], False)
        dev, nf = C.POINTER(C.c_float)(), C.c_int()
        assert rend._lib.spv_batch_wait(rend._ctx, which, None, C.byref(dev), C.byref(nf)) == 0 and nf.value == 8
        import torch
        n = size[0] * size[1]
        host = np.empty(8 * 2 * n, np.float32)
        assert torch.cuda.is_available()
        from cuda import cudart  # cuda-python: a plain device -> host copy of the set's planes
        err, = cudart.cudaMemcpy(host.ctypes.data, C.cast(dev, C.c_void_p).value, host.nbytes,
                                 cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0
        for f in range(8):
            assert np.array_equal(host[2 * f * n:(2 * f + 1) * n].reshape(size[1], size[0]), singles[4 + f][0]), f
            assert np.array_equal(host[(2 * f + 1) * n:(2 * f + 2) * n].reshape(size[1], size[0]), singles[4 + f][1]), f
    finally:
        rend.close()


def test_float_volumes_through_the_copies(oracle_mod):
    """max_project_float (volume_kernel.cl:97-185): float32 volumes live in a 3-D array, every layered pair copy is built
    from it; misses read alpha -1 (also in the tiles and rows that are neither rendered nor copied), launches of several
    frames equal single frames, attenuation follows the float law, and a renderer that changes element type between
    launches keeps the right miss values"""
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate
    vol = scenes.vol_g(80, np.float32, seed=2)
    size = (208, 160)
    rend = _renderer(size)
    rend.set_data(vol)
    rend.set_max_val(1.)
    P = scenes.gui_camera(0., 4.)[1]
    rend.set_projection(P)
    cams = [scenes.gui_camera(0.5 * i, 2.8 + 0.3 * (i % 4))[0] for i in range(9)]
    cams[4] = np.dot(mat4_translate(1.5, 0.2, -4.), mat4_rotation(0.5, 0., 1., 0.))   # partly off screen
    cams[6] = np.dot(mat4_translate(9., 0., -4.), mat4_rotation(0.5, 0., 1., 0.))     # wholly off screen
    try:
        singles = []
        for M in cams:
            rend.render(modelView=M)
            assert rend.mip_axis_used()[0] >= 0
            singles.append((rend.output.copy(), rend.output_alpha.copy()))
            ref, ref_a = _oracle_image(oracle_mod, vol, size, M, P, max_val=1.)
            assert np.abs(singles[-1][0] - ref).max() < 1e-3
            assert np.array_equal(singles[-1][1], ref_a)
        assert set(np.unique(singles[0][1])) == {-1., 1.} and np.all(singles[6][1] == -1.)
        for n in (9, 3, 9):
            frames = rend.batch_frames_of(rend.render_batch(cams[:n]), copy=True)
            for f, ((o, a), (so, sa)) in enumerate(zip(frames, singles)):
                assert np.array_equal(o, so) and np.array_equal(a, sa), (n, f)
        # attenuated, float law
        rend.set_alpha_pow(0.8)
        ref, ref_a = _oracle_image(oracle_mod, vol, size, cams[1], P, max_val=1., alpha_pow=0.8)
        att = rend.batch_frames_of(rend.render_batch(cams[:3]), copy=True)
        assert np.abs(att[1][0] - ref).max() < 1e-3 and np.array_equal(att[1][1], ref_a)
        rend.set_alpha_pow(0.)
        # the same planes after an integer volume (miss alpha 0) and back
        u16 = scenes.vol_g(64, np.uint16, seed=3)
        rend.set_data(u16)
        rend.set_max_val(60000.)
        want = []
        for M in cams[:5]:
            rend.render(modelView=M)
            want.append((rend.output.copy(), rend.output_alpha.copy()))
        got = rend.batch_frames_of(rend.render_batch(cams[:5]), copy=True)
        assert all(np.array_equal(o, wo) and np.array_equal(a, wa) for (o, a), (wo, wa) in zip(got, want))
        assert got[4][1].min() == 0.
        rend.set_data(vol)
        rend.set_max_val(1.)
        for _ in range(2):   # both sets of planes
            frames = rend.batch_frames_of(rend.render_batch(cams), copy=True)
            assert all(np.array_equal(o, so) and np.array_equal(a, sa) for (o, a), (so, sa) in zip(frames, singles))
        rend.set_view_copies("primary")   # float32 volumes have no primary pair copy: mip_fast_kernel
        rend.render(modelView=cams[0])
        assert rend.mip_axis_used() == (-1, -1) and np.abs(rend.output - singles[0][0]).max() < 1e-3
    finally:
        rend.close()


def test_render_sequence_in_launches_of_several_frames():
    """render_sequence(batch=...) yields the frames render() gives, in order, for every launch size; settings that the
    multi-frame launch does not cover (attenuation, float volumes, the exact sampler) fall back to one launch per frame"""
    vol = scenes.vol_g(64, np.uint16, seed=6)
    size = (192, 160)
    rend = _renderer(size)
    rend.set_data(vol)
    rend.set_max_val(50000.)
    rend.set_gamma(0.8)
    rend.set_box_boundaries([-.9, .8, -1, 1, -.7, 1])
    cams = [scenes.gui_camera(0.29 * i, 3.4) for i in range(23)]
    rend.set_projection(cams[0][1])
    try:
        want = []
        for M, _ in cams:
            rend.render(modelView=M)
            want.append((rend.output.copy(), rend.output_alpha.copy()))
        # a list of views is rendered several frames per launch by default; an iterator (which may change the renderer as
        # it is pulled) one launch per frame unless the caller asks for more
        for views, batch in [([M for M, _ in cams], None), ((M for M, _ in cams), None), ((M for M, _ in cams), 1),
                             ((M for M, _ in cams), 2), ([M for M, _ in cams], 7), ((M for M, _ in cams), 16),
                             ([M for M, _ in cams], 50)]:
            launches0 = rend.launch_count()
            got = [(r.output.copy(), r.output_alpha.copy(), r.modelView.copy()) for r in
                   rend.render_sequence(views, batch=batch)]
            launches = rend.launch_count() - launches0
            per_launch = 1 if batch == 1 or (batch is None and not isinstance(views, list)) else min(batch or 10, 16)
            # (the first launch of a sequence takes half a batch from 4 frames per launch up)
            first = (per_launch + 1) // 2 if per_launch >= 4 else per_launch
            assert launches == 1 + -(-(len(cams) - first) // per_launch), (batch, launches)
            assert len(got) == len(want)
            for i, ((o, a, M), (wo, wa)) in enumerate(zip(got, want)):
                assert np.array_equal(o, wo) and np.array_equal(a, wa), (batch, i)
                if per_launch > 1:  # (one launch per frame: the next frame has been issued -- its modelView set -- already)
                    assert np.allclose(M, cams[i][0])
            assert np.allclose(rend.modelView, cams[-1][0])
        # attenuated projections take the same launches
        rend.set_alpha_pow(0.5)
        want_att = []
        for M, _ in cams[:12]:
            rend.render(modelView=M)
            want_att.append(rend.output.copy())
        assert rend.mip_axis_used() == (1, 1)
        assert not np.array_equal(want_att[3], want[3][0])
        for views in ([M for M, _ in cams[:12]], (M for M, _ in cams[:12])):
            att = [r.output.copy() for r in rend.render_sequence(views)]
            assert len(att) == 12 and all(np.array_equal(x, y) for x, y in zip(att, want_att))
        rend.set_alpha_pow(0.)
        with pytest.raises(Exception):
            rend.render_batch([M for M, _ in cams[:17]])
        # settings the launches of several frames do not cover fall back to one launch per frame
        rend.set_skipping(True)
        sk = [r.output.copy() for r in rend.render_sequence([M for M, _ in cams[:5]])]
        assert rend.mip_axis_used() == (-1, -1)
        assert all(np.abs(x - w[0]).max() < 4e-3 for x, w in zip(sk, want))
        with pytest.raises(Exception):
            rend.render_batch([M for M, _ in cams[:3]])
        rend.set_skipping(None)
    finally:
        rend.close()


def test_c2_sweep_in_launches_against_oracle_rows(oracle_mod):
    """configs[1] at its own size the way bench.py renders it: 20 frames 18 degrees apart, 10 per launch, every 64th row
    of 6 of them against the oracle"""
    vol = scenes.vol_g(512, np.uint16, seed=0)
    rend = _renderer((1024, 1024), max_steps=200)
    rend.set_data(vol)
    rend.set_max_val(60000.)
    cams = [scenes.gui_camera(2 * math.pi * j / 20., 4.0) for j in range(20)]
    rend.set_projection(cams[0][1])
    o = oracle_mod.OracleRenderer((1024, 1024), kind="port")
    o.set_data(vol)
    o.set_projection(cams[0][1])
    rows = slice(0, 1024, 64)
    worst = 0.
    try:
        o.lib.so_set_row_sampling(0, 64)
        for i, r in enumerate(rend.render_sequence((M for M, _ in cams), batch=10)):
            if i % 4 != 1:
                continue
            o.set_modelView(cams[i][0])
            o.render(maxVal=60000.)
            err = float(np.abs(o.output[rows] - r.output[rows]).max())
            worst = max(worst, err)
            assert err < 1e-3, (i, err)
            assert np.array_equal(o.output_alpha[rows], r.output_alpha[rows]), i
        assert rend.mip_axis_used() == (1, 1)
    finally:
        o.lib.so_set_row_sampling(0, 1)
        rend.close()
    print("C2 in launches of 10: worst max |gpu - oracle| = %.3g" % worst)
