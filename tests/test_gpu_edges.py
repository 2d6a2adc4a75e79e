"""Edge cases of the hot path on the GPU, against the oracle: degenerate and ragged volume extents, image sizes
that do not fill a tile, the layered-array limit, empty volumes, cameras that see nothing or sit inside the volume."""
import numpy as np
import pytest

import scenes
from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_rotation, mat4_translate

pytestmark = pytest.mark.gpu


def _renderer(size, **kw):
    from spimagine_b200 import VolumeRenderer
    return VolumeRenderer(size, **kw)


def _pair(oracle_mod, size, data, M, P, interp="linear", **kw):
    o = oracle_mod.OracleRenderer(size, interpolation=interp, kind="port")
    g = _renderer(size, interpolation=interp, sampler="exact", **kw)
    for r in (o, g):
        r.set_data(data)
        r.set_modelView(M)
        r.set_projection(P)
    return o, g


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 7, 5), (9, 1, 4), (6, 5, 1), (2, 2, 2), (3, 300, 2), (33, 8, 9)])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_degenerate_and_ragged_extents_are_bit_exact(oracle_mod, shape, dtype):
    """Axes of one texel (every sample clamps to the edge), extents that are no multiple of the 8^3 bricks, a volume
    that is one voxel thick: the exact sampler equals the oracle bit for bit; the texture-unit path agrees on the hit
    mask, stays inside the value range and finds the same surface pixels with and without empty-space skipping."""
    data = scenes.random_vol(shape, dtype, seed=sum(shape))
    M, P = scenes.gui_camera(0.7, 3.1)
    o, g = _pair(oracle_mod, (61, 43), data, M, P)
    peak = float(data.max()) if float(data.max()) > 0 else 1.
    for r in (o, g):
        r.render(maxVal=peak)
    assert np.array_equal(g.output, o.output) and np.array_equal(g.output_alpha, o.output_alpha)
    mip_alpha = o.output_alpha.copy()
    for r in (o, g):
        r.render(maxVal=peak * .7, method="iso_surface_raw")
    assert np.array_equal(g.output_depth, o.output_depth)
    assert np.array_equal(g.output_normals, o.output_normals)
    t = _renderer((61, 43))
    t.set_data(data)
    t.set_modelView(M)
    t.set_projection(P)
    t.render(maxVal=peak)
    assert np.array_equal(t.output_alpha, mip_alpha)
    assert t.output.min() >= 0 and t.output.max() <= 1 + 1e-6
    assert t.data_min_max == (float(data.min()), float(data.max()))
    planes = []
    for skip in (False, True):
        t.set_skipping(skip)
        t.render(maxVal=peak * .7, method="iso_surface")
        planes.append([a.copy() for a in (t.output, t.output_depth, t.output_normals, t.output_occlusion)])
    for a, b in zip(*planes):
        assert np.array_equal(a, b)
    for r in (g, t):
        r.close()


@pytest.mark.parametrize("size", [(1, 1), (3, 5), (17, 9), (8, 4), (130, 66)])
def test_image_sizes_that_do_not_fill_a_tile(oracle_mod, size):
    data = scenes.vol_g(24, np.uint16, seed=4)
    M, P = scenes.gui_camera(0.3, 2.6)
    o, g = _pair(oracle_mod, size, data, M, P)
    for r in (o, g):
        r.render(maxVal=60000.)
    assert g.output.shape == size[::-1]
    assert np.array_equal(g.output, o.output) and np.array_equal(g.output_alpha, o.output_alpha)
    t = _renderer(size)
    t.set_data(data)
    t.set_modelView(M)
    t.set_projection(P)
    t.render(maxVal=60000.)
    assert np.abs(t.output - o.output).max() < 4e-3 and np.array_equal(t.output_alpha, o.output_alpha)
    o.render(maxVal=30000., method="iso_surface")
    g.render(maxVal=30000., method="iso_surface")
    assert np.array_equal(g.output_depth, o.output_depth) and np.array_equal(g.output_normals, o.output_normals)
    t.render(maxVal=30000., method="iso_surface")           # tile flags, blur and occlusion queue at ragged sizes
    both = np.isfinite(t.output_depth) & np.isfinite(o.output_depth)
    assert (np.isfinite(t.output_depth) != np.isfinite(o.output_depth)).mean() < 0.05
    if both.any():
        assert np.abs(t.output_depth[both] - o.output_depth[both]).max() < 0.05
    seq = [r.output.copy() for r in t.render_sequence([M, M])]
    t.render(maxVal=30000.)
    assert np.array_equal(seq[0], t.output) and np.array_equal(seq[1], t.output)
    for r in (g, t):
        r.close()


def test_more_slices_than_a_layered_array_holds(oracle_mod):
    """2048 layers is the limit of a layered CUDA array: a taller integer volume silently uses the plain 3-D layout;
    rows of the frame against the oracle."""
    data = scenes.random_vol((2051, 6, 10), np.uint16, seed=9)
    M = np.dot(mat4_translate(0, 0, -3.2), mat4_rotation(1.2, 1, 0.2, 0))
    P = mat4_perspective(50, 1., .1, 10)
    o, g = _pair(oracle_mod, (48, 40), data, M, P, interp="nearest")
    for r in (o, g):
        r.set_units([40., 40., 1.])          # a column, not a needle: enough pixels see it
        r.set_modelView(M)                   # units take effect with the next matrix update, as in the reference
        r.render(maxVal=65535.)
    assert np.array_equal(g.output, o.output) and np.array_equal(g.output_alpha, o.output_alpha)
    assert (o.output_alpha > 0).sum() > 50
    t = _renderer((48, 40), interpolation="nearest")
    t.set_data(data)
    t.set_units([40., 40., 1.])
    t.set_modelView(M)
    t.set_projection(P)
    t.render(maxVal=65535.)
    assert np.array_equal(t.output_alpha, o.output_alpha)
    # nearest sampling of white noise: the texture-unit path places sample k at pos0 + k * delta (one fma) instead of
    # accumulating, which picks a neighbouring voxel now and then -- most ray maxima agree, all are voxel values
    assert np.mean(np.abs(t.output - o.output) > 1e-6) < 0.25
    vals = np.unique(np.rint(t.output[t.output_alpha > 0] * 65535.).astype(np.int64))
    assert np.isin(vals, np.unique(data)).all()
    t3 = _renderer((48, 40), interpolation="nearest")
    t3.set_layout("3d")
    t3.set_data(data)
    t3.set_units([40., 40., 1.])
    t3.set_modelView(M)
    t3.set_projection(P)
    t3.render(maxVal=65535.)
    assert np.array_equal(t3.output, t.output)               # the fallback IS the 3-D layout
    for r in (g, t, t3):
        r.close()


def test_empty_volume_and_cameras_that_see_nothing(oracle_mod):
    zeros = np.zeros((20, 24, 28), np.uint16)
    M, P = scenes.gui_camera(0.2, 3.)
    t = _renderer((64, 48))
    t.set_data(zeros)
    t.set_modelView(M)
    t.set_projection(P)
    t.render(maxVal=100.)
    assert not t.output.any() and (t.output_alpha > 0).any()          # the box is hit, the volume is empty
    t.render(maxVal=100., method="iso_surface")
    assert not t.output.any() and np.isinf(t.output_depth).all() and not t.output_normals.any()
    assert not t.output_occlusion.any()
    assert t.data_min_max == (0., 0.)
    # looking away from the volume: every pixel misses
    away = np.dot(mat4_rotation(np.pi, 0, 1, 0), mat4_translate(0, 0, -3.))
    data = scenes.vol_g(24, np.uint16, seed=1)
    o, g = _pair(oracle_mod, (64, 48), data, away, P)
    for r in (o, g, t):
        r.set_data(data) if r is t else None
        r.set_modelView(away)
        r.render(maxVal=60000.)
    # The volume is BEHIND the camera: the slab test still reports tfar > tnear (both negative), tnear is clamped to 0
    # and the reference marches |tfar - tnear| forward from the eye through clamp-to-edge texels (volume_kernel.cl:
    # 270-300) -- a quirk every implementation of this path has to share.  alpha = tnear = 0 everywhere.
    assert not o.output_alpha.any()
    assert np.array_equal(g.output, o.output) and np.array_equal(g.output_alpha, o.output_alpha)
    assert np.abs(t.output - o.output).max() < 4e-3 and not t.output_alpha.any()
    # looking sideways: the rays really miss the box
    side = np.dot(mat4_rotation(np.pi / 2, 0, 1, 0), mat4_translate(0, 0, -3.))
    for r in (o, g, t):
        r.set_modelView(side)
        r.render(maxVal=60000.)
    assert not o.output.any() and not o.output_alpha.any()
    assert not g.output.any() and not t.output.any() and not t.output_alpha.any()
    # camera inside the volume: tnear is clamped to 0 (alpha = 0 on the integer path although the ray hits)
    inside = mat4_translate(0.05, -0.02, -0.3)
    for r in (o, g, t):
        r.set_modelView(inside)
        r.render(maxVal=60000.)
    assert np.array_equal(g.output, o.output) and np.array_equal(g.output_alpha, o.output_alpha)
    assert np.array_equal(t.output_alpha, o.output_alpha) and np.abs(t.output - o.output).max() < 4e-3
    assert o.output.max() > 0.1
    for r in (o, g):
        r.render(maxVal=30000., method="iso_surface_raw")
    assert np.array_equal(g.output_depth, o.output_depth)
    g.close()
    t.close()


@pytest.mark.gpu
def test_texture_rate_probes():
    """Roofline calibration (bench.py): the peak probe and the footprint probe return plausible rates, a footprint
    that spreads a quad over several layers is slower than one inside a layer, bad vectors are refused."""
    from spimagine_b200 import VolumeRenderer, _lib
    r = VolumeRenderer((64, 64))
    try:
        with pytest.raises(_lib.SpvError):
            r.texrate_probe(10)                                  # no volume yet
        r.set_data(scenes.vol_g(64, np.uint16, seed=0))
        peak = r.texrate_probe(400)
        in_layer = r.texrate_probe(400, footprint=[[1.15, 0, 0], [0, 1.15, 0], [0, 0, 1.6]])
        across = r.texrate_probe(400, footprint=[[0, 0, 1.15], [0, 1.15, 0], [1.6, 0, 0]])
        assert 1e11 < across < in_layer <= 1.1 * peak < 3e12, (peak, in_layer, across)
        with pytest.raises(_lib.SpvError):
            r.texrate_probe(400, footprint=[[100., 0, 0], [0, 1, 0], [0, 0, 1]])
        with pytest.raises(_lib.SpvError):
            r.texrate_probe(0)
    finally:
        r.close()
