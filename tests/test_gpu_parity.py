"""GPU parity tests: libspimcuda (through the drop-in VolumeRenderer, i.e. through the C ABI) against
  * the committed golden vectors rendered with the reference's own kernel text,
  * the CPU oracle on seeded inputs,
  * analytic known answers and size-independent properties.
Bars: the EXACT sampler is bit-exact (it evaluates the reference's fp32 expressions in the same order); the
TMU sampler is within 1e-3 of the dynamic range per pixel for max projection, within one ray step for the
iso-surface hit depth and 1e-2 for normals (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import golden_cases
import scenes
from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_rotation, mat4_translate

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _renderer(size, **kw):
    from spimagine_b200 import VolumeRenderer
    return VolumeRenderer(size, **kw)


def _maxdiff(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d = np.where(np.isfinite(d), d, np.where(a == b, 0, np.inf))
    return float(d.max())


# ---------------------------------------------------------------------------------------------- golden vectors
# pow(), cos()/sin() are library functions whose last bits differ between glibc and CUDA: those outputs get a
# tolerance, everything else must be identical.
_POW_CASES = {"mip_f32_gamma"}


@pytest.mark.parametrize("name", sorted(golden_cases.CASES))
def test_exact_sampler_matches_golden(name):
    case = golden_cases.CASES[name]
    rend = _renderer(golden_cases.SIZE, interpolation=case.get("interpolation", "linear"), sampler="exact")
    res = golden_cases.run_case(rend, name)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    iso = case.get("method") == "iso_surface"
    for k in gold.files:
        g, r = gold[k], res[k]
        assert g.shape == r.shape
        if iso and k == "output":          # Phong: pow(x, 10)
            assert _maxdiff(r, g) < 2e-5, k
        elif iso and k == "occlusion":     # disc sampling: cos/sin last-bit differences can move a tap one pixel
            assert np.mean(np.abs(r - g) > 1e-6) < 0.01 and _maxdiff(r, g) < 0.1, k
        elif iso and k == "normals" and case.get("render", {}).get("gamma", 1.) != 1.:
            assert _maxdiff(r, g) < 1e-5, k  # h = dt * pow(gamma, 2)
        elif name in _POW_CASES and k == "output":
            assert _maxdiff(r, g) < 1e-6, k
        else:
            assert np.array_equal(r, g), "%s/%s: max |d| = %g, %d pixels differ" % (
                name, k, _maxdiff(r, g), int((r != g).sum()))


@pytest.mark.parametrize("name", sorted(golden_cases.CONFIG_CASES))
def test_config_case_matches_golden(name):
    """BASELINE.json configs[0] at its own size: exact sampler bit for bit, texture-unit sampler within 1e-3 of the
    dynamic range (north_star), hit mask identical."""
    case = golden_cases.CONFIG_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    rend = _renderer(case["size"], sampler="exact")
    res = golden_cases.run_config_case(rend, name)
    assert np.array_equal(res["output"], gold["output"]) and np.array_equal(res["alpha"], gold["alpha"])
    rend.close()
    rend = _renderer(case["size"], sampler="tmu")
    res = golden_cases.run_config_case(rend, name)
    assert _maxdiff(res["output"], gold["output"]) < 1e-3 * case["render"]["maxVal"]
    assert np.array_equal(res["alpha"], gold["alpha"])
    rend.close()


@pytest.mark.parametrize("name", [n for n in sorted(golden_cases.CASES)
                                  if "random" not in n and "nearest" not in n and "alpha" not in n])
def test_tmu_sampler_close_to_golden(name):
    """Hardware filtering (8-bit weights) on the smooth golden scenes: north_star tolerances."""
    case = golden_cases.CASES[name]
    rend = _renderer(golden_cases.SIZE, interpolation=case.get("interpolation", "linear"), sampler="tmu")
    res = golden_cases.run_case(rend, name)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    if case.get("method") == "iso_surface":
        g_hit, r_hit = np.isfinite(gold["depth"]), np.isfinite(res["depth"])
        both = g_hit & r_hit
        assert (g_hit != r_hit).mean() < 0.02          # silhouette pixels may flip
        dt = 4. / (200 - 1)                            # generous bound on one ray step for these cameras
        err = np.abs(res["depth"][both] - gold["depth"][both])
        assert np.percentile(err, 95) <= dt and np.median(err) <= dt / 4
        assert np.median(np.abs(res["normals"][both] - gold["normals"][both])) < 2e-2
    else:
        # these 24^3..32^3 scenes change by a third of their range from one voxel to the next; the 8-bit weights
        # of the texture unit are worth up to 3/512 of the local neighbour difference (the smooth-volume bar of
        # 1e-3 is tested on Vol-G below)
        for k in gold.files:
            if k.startswith("output"):
                assert _maxdiff(res[k], gold[k]) < 4e-3, k   # outputs are normalised to [0, 1]
            else:
                assert np.array_equal(res[k], gold[k]), k      # alpha depends on the box test only


# ---------------------------------------------------------------------------------------------- oracle, seeded
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("interp", ["linear", "nearest"])
def test_exact_sampler_bitwise_on_random_volumes(oracle_mod, dtype, interp):
    data = scenes.random_vol((17, 23, 29), dtype, seed=11)
    M, P = scenes.gui_camera(0.8, 3.3)
    o = oracle_mod.OracleRenderer((50, 38), interpolation=interp, kind="port")
    g = _renderer((50, 38), interpolation=interp, sampler="exact")
    for r in (o, g):
        r.set_data(data)
        r.set_units([1., .8, 2.])
        r.set_modelView(M)
        r.set_projection(P)
    for alpha_pow in (0., .7):
        for r in (o, g):
            r.set_alpha_pow(alpha_pow)
            r.render(maxVal=float(data.max()), minVal=1.)
        assert np.array_equal(g.output, o.output), "alpha_pow=%g: max |d| = %g" % (alpha_pow, _maxdiff(g.output, o.output))
        assert np.array_equal(g.output_alpha, o.output_alpha)
    for r in (o, g):
        r.set_alpha_pow(0.)
        r.render(maxVal=float(data.max()) * .6, method="iso_surface_raw")
    assert np.array_equal(g.output_depth, o.output_depth)
    assert np.array_equal(g.output_alpha, o.output_alpha)
    assert np.array_equal(g.output_normals, o.output_normals)
    assert _maxdiff(g.output, o.output) < 2e-5


@pytest.mark.parametrize("layout", ["zpair", "3d"])
@pytest.mark.parametrize("dtype,peak", [(np.float32, 1.), (np.uint16, 60000.), (np.uint8, 250.)])
def test_tmu_sampler_within_tolerance_of_reference_semantics(oracle_mod, dtype, peak, layout):
    """Vol-G (smooth blobs + 1 % noise), 96^3 -> 256x192: TMU path vs the fp32 oracle, 1e-3 of the range."""
    data = scenes.vol_g(96, dtype, seed=0)
    M, P = scenes.gui_camera(0.5, 3.4)
    o = oracle_mod.OracleRenderer((256, 192), kind="port")
    g = _renderer((256, 192))
    g.set_layout(layout)
    for r in (o, g):
        r.set_data(data)
        r.set_modelView(M)
        r.set_projection(P)
        r.render(maxVal=peak)
    assert np.array_equal(g.output_alpha, o.output_alpha)
    assert _maxdiff(g.output, o.output) < 1e-3
    # and much closer to the oracle's model of the texture unit (8-bit weights, fma positions)
    o8 = oracle_mod.OracleRenderer((256, 192), kind="port", weight_bits=8, pos_mode=2)
    o8.set_data(data)
    o8.set_modelView(M)
    o8.set_projection(P)
    o8.render(maxVal=peak)
    assert _maxdiff(g.output, o8.output) < 5e-4


def test_tmu_nearest_is_bit_exact(oracle_mod):
    """Point sampling has no weights to quantise: the TMU path must return exact voxel values."""
    for dtype in (np.uint16, np.float32, np.uint8):
        data = scenes.random_vol((21, 18, 30), dtype, seed=4)
        M, P = scenes.tilted_camera()
        o = oracle_mod.OracleRenderer((64, 48), interpolation="nearest", kind="port", pos_mode=2)
        g = _renderer((64, 48), interpolation="nearest")
        for r in (o, g):
            r.set_data(data)
            r.set_modelView(M)
            r.set_projection(P)
            r.render(maxVal=float(data.max()))
        # positions differ from the oracle's by < 1 ulp of the coordinate: a sample exactly on a voxel boundary
        # could pick the neighbour; allow a handful of pixels, demand exact voxel values everywhere
        assert (g.output != o.output).mean() < 0.002
        vals = np.unique(data).astype(np.float32) / np.float32(data.max())
        assert np.isin(g.output[g.output > 0], np.clip(vals, 0, 1)).all()


@pytest.mark.parametrize("layout", ["zpair", "3d"])
def test_skipping_does_not_change_the_image(layout):
    for dtype, peak in ((np.uint16, 60000.), (np.float32, 1.)):
        data = scenes.vol_g(80, dtype, seed=7)
        g = _renderer((200, 168))
        g.set_layout(layout)
        g.set_data(data)
        g.enable_stats(True)
        outs = []
        for theta in (0.2, 1.9):
            M, P = scenes.gui_camera(theta, 2.8)
            g.set_modelView(M)
            g.set_projection(P)
            pair = []
            for skip in (False, True):
                g.set_skipping(skip)
                g.render(maxVal=peak)
                pair.append((g.output.copy(), g.output_alpha.copy(), g.last_stats()))
            assert np.array_equal(pair[0][0], pair[1][0])
            assert np.array_equal(pair[0][1], pair[1][1])
            hits, fetched = pair[0][2]
            assert fetched == hits * 208            # (200/16+1)*16 samples per hit ray
            assert pair[1][2][0] == hits and pair[1][2][1] < fetched
            outs.append(pair)


def test_constant_volume_and_misses():
    g = _renderer((31, 29))
    g.set_data(np.full((9, 9, 9), 123., np.float32))
    M, P = scenes.gui_camera(0.3, 4.)
    g.set_modelView(M)
    g.set_projection(P)
    for sampler in ("tmu", "exact"):
        g.set_sampler(sampler)
        g.render(maxVal=200.)
        hit = g.output_alpha > 0
        assert 0 < hit.sum() < hit.size
        np.testing.assert_allclose(g.output[hit], 123. / 200., rtol=1e-6)
        assert (g.output[~hit] == 0).all() and (g.output_alpha[~hit] == -1).all()
    g.set_modelView(mat4_translate(30., 0, -4.))
    g.render(maxVal=200.)
    assert (g.output == 0).all() and (g.output_alpha == -1).all()
    lo, hi = g.data_min_max
    assert lo == 123. and hi == 123.


def test_data_min_max_and_update_data():
    g = _renderer((16, 16))
    a = scenes.random_vol((9, 20, 33), np.uint16, 1)
    g.set_data(a)
    assert g.data_min_max == (float(a.min()), float(a.max()))
    b = (a // 3).astype(np.uint16)
    g.update_data(b)
    assert g.data_min_max == (float(b.min()), float(b.max()))
    assert g.dataImg.shape == (33, 20, 9) and g.dataImg.dtype == np.uint16


def test_multipass_rendering_matches_oracle(oracle_mod):
    data = scenes.two_blobs(40)
    M, P = scenes.tilted_camera()
    for sampler, kw in (("exact", dict()), ("tmu", dict(pos_mode=2, weight_bits=8))):
        o = oracle_mod.OracleRenderer((72, 60), kind="port", **kw)
        g = _renderer((72, 60), sampler=sampler)
        for r in (o, g):
            r.set_data(data)
            r.set_modelView(M)
            r.set_projection(P)
        for part in range(4):
            for r in (o, g):
                r.render(maxVal=255., numParts=4, currentPart=part)
            if sampler == "exact":
                assert np.array_equal(g.output, o.output)
            else:
                assert _maxdiff(g.output, o.output) < 2e-3   # two_blobs(40) is steep at voxel scale
            assert np.array_equal(g.output_alpha, o.output_alpha)


def test_api_errors_like_the_reference():
    from spimagine_b200 import VolumeRenderer
    with pytest.raises(KeyError):
        VolumeRenderer((8, 8), interpolation="cubic")
    g = VolumeRenderer((8, 8))
    with pytest.raises(NotImplementedError):
        g.set_data(np.zeros((4, 4, 4), np.float64), autoConvert=False)
    with pytest.raises(NotImplementedError):
        g.set_dtype(np.int32)
    assert g.render() is None              # no data: prints and returns None
    g.set_data(np.ones((4, 4, 4), np.float64))   # autoConvert -> float32
    assert g.dtype == np.float32
    assert g.render(maxVal=1.) is None     # the reference's render() returns None too
    assert g.output.shape == (8, 8) and g.output.dtype == np.float32
    g.resize((12, 10))
    g.render(maxVal=1.)
    assert g.output.shape == (10, 12)


# ---------------------------------------------------------------------------------------------- the sampler alone
def _points(n, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-0.05, 1.05, (n, 3)).astype(np.float32)      # includes clamp-to-edge positions
    p[:64] = rng.integers(0, 2, (64, 3)).astype(np.float32)        # exact corners
    return p


@pytest.mark.parametrize("layout", ["zpair", "3d"])
@pytest.mark.parametrize("interp", ["linear", "nearest"])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
def test_exact_sampler_is_the_opencl_sampler_bitwise(oracle_mod, dtype, interp, layout):
    """read_imagef(volume, sampler, pos).x for 20000 positions: libspimcuda's exact sampler vs the oracle."""
    data = scenes.random_vol((13, 22, 31), dtype, seed=8)
    pos = _points(20000)
    g = _renderer((8, 8), interpolation=interp, sampler="exact")
    g.set_layout(layout)
    g.set_data(data)
    got = g.sample_points(pos)
    V = oracle_mod.make_volume(data, filter_linear=(interp == "linear"))
    lib = oracle_mod.load("port")
    want = np.array([lib.so_sample(__import__("ctypes").byref(V), float(x), float(y), float(z)) for x, y, z in pos[:4000]],
                    np.float32)
    assert np.array_equal(got[:4000], want)


@pytest.mark.parametrize("layout", ["zpair", "3d"])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
def test_tmu_sampler_error_is_bounded_by_weight_quantisation(dtype, layout):
    """|tex3D - fp32 trilinear| <= (axes / 512 + eps) * (max - min of the 8 texels): 8 fractional weight bits on
    every axis the texture unit filters (3 for the 3-D layout, 2 for the z-paired layout)."""
    data = scenes.random_vol((24, 28, 36), dtype, seed=12)
    pos = np.random.default_rng(5).uniform(0.03, 0.97, (50000, 3)).astype(np.float32)
    g = _renderer((8, 8))
    g.set_layout(layout)
    g.set_data(data)
    tmu = g.sample_points(pos).astype(np.float64)
    g.set_sampler("exact")
    ex = g.sample_points(pos).astype(np.float64)
    nz, ny, nx = data.shape
    i0 = np.floor(pos[:, 0] * np.float32(nx) - np.float32(.5)).astype(int)
    j0 = np.floor(pos[:, 1] * np.float32(ny) - np.float32(.5)).astype(int)
    k0 = np.floor(pos[:, 2] * np.float32(nz) - np.float32(.5)).astype(int)
    lo, hi = np.full(len(pos), np.inf), np.full(len(pos), -np.inf)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                v = data[np.clip(k0 + dz, 0, nz - 1), np.clip(j0 + dy, 0, ny - 1), np.clip(i0 + dx, 0, nx - 1)].astype(np.float64)
                lo, hi = np.minimum(lo, v), np.maximum(hi, v)
    axes = 2 if (layout == "zpair" and dtype != np.float32) else 3
    full = 1. if dtype == np.float32 else float(np.iinfo(dtype).max)
    bound = (axes / 512. + 1e-3) * (hi - lo) + 2. * full / 65536.   # + the unit's output rounding for integer formats
    assert ((np.abs(tmu - ex) <= bound).mean()) > 0.999
    assert np.abs(tmu - ex).max() <= 1.5 * bound.max()
    assert (tmu >= lo - 1e-3 * full).all() and (tmu <= hi + 1e-3 * full).all()   # a convex combination


# ---------------------------------------------------------------------------------------------- iso surface
@pytest.mark.parametrize("layout", ["zpair", "3d"])
def test_iso_sphere_tmu_within_north_star_tolerance(oracle_mod, layout):
    """uint16 sphere 45000*exp(-10 R), iso at 500: hit depth within one ray step, normals within 1e-2."""
    N = 96
    Z, Y, X = scenes.grid(N)
    data = (45000. * np.exp(-10. * np.sqrt(X ** 2 + Y ** 2 + Z ** 2))).astype(np.uint16)
    o = oracle_mod.OracleRenderer((160, 160), kind="port")
    g = _renderer((160, 160))
    g.set_layout(layout)
    for r in (o, g):
        r.set_data(data)
        r.set_modelView(np.dot(mat4_translate(0, 0, -5), mat4_rotation(.4, 1, .5, 0)))
        r.set_projection(mat4_perspective())
        r.render(maxVal=1000., method="iso_surface_raw")
    oh, gh = np.isfinite(o.output_depth), np.isfinite(g.output_depth)
    assert (oh != gh).mean() < 0.005
    both = oh & gh
    dt = 2. * np.sqrt(3) / 199
    err = np.abs(g.output_depth[both] - o.output_depth[both])
    assert err.max() <= dt * 1.01
    assert np.percentile(err, 99) <= dt / 5          # nearly always the same refinement sub-step or the next
    with np.errstate(invalid="ignore"):
        same = both & (np.abs(g.output_depth - o.output_depth) < 1e-5)
    assert same.sum() > 0.8 * both.sum()
    assert np.percentile(np.abs(g.output_normals[same] - o.output_normals[same]), 99) < 1e-2


@pytest.mark.parametrize("layout", ["zpair", "3d"])
def test_iso_skipping_does_not_change_the_image(layout):
    """Empty-space skipping in the iso-surface search is exact: identical buffers with it on and off, from
    outside the surface and from a camera that starts inside it (isGreater == true)."""
    for dtype, peak in ((np.uint16, 60000.), (np.float32, 1.)):
        data = scenes.vol_g(112, dtype, seed=5)
        g = _renderer((240, 200))
        g.set_layout(layout)
        g.set_data(data)
        g.enable_stats(True)
        for theta, dist, frac in ((0.3, 3.0, .3), (1.4, 2.4, .1), (2.0, 0.2, .05)):
            M, P = scenes.gui_camera(theta, dist)
            g.set_modelView(M)
            g.set_projection(P)
            res = []
            for skip in (False, True):
                g.set_skipping(skip)
                g.render(maxVal=2 * frac * peak, method="iso_surface")
                res.append((g.output.copy(), g.output_depth.copy(), g.output_normals.copy(), g.output_alpha.copy(),
                            g.output_occlusion.copy(), g.last_stats()))
            for a, b in zip(res[0][:5], res[1][:5]):
                assert np.array_equal(a, b)
            assert res[1][5][1] < res[0][5][1]          # fewer texture samples issued
            assert np.isfinite(res[0][1]).sum() > 500   # the scene does show a surface


@pytest.mark.parametrize("shape", [(300, 300, 300), (150, 330, 270)])
def test_iso_cell_traversal_is_exact_on_large_and_anisotropic_volumes(shape):
    """The hierarchical traversal (128^3 / 32^3 cells crossed in one step) only pays off on volumes with several
    top-level cells per axis; it must stay invisible there too, for non-cubic grids, anisotropic voxels, a reduced
    box, a box larger than the volume (clamp-to-edge samples) and tilted cameras."""
    data = scenes.vol_g(0, np.uint16, seed=7, shape=shape)
    g = _renderer((256, 192))
    g.set_data(data)
    g.set_units([1., .7, 2.1])
    g.enable_stats(True)
    M2 = np.dot(mat4_translate(0.1, -0.05, -3.1), np.dot(mat4_rotation(0.8, 0.3, 1, 0.25), mat4_rotation(0.5, 1, 0, 0)))
    cams = [scenes.gui_camera(0.4, 3.4), (M2, mat4_perspective(50, 1., .1, 10)), scenes.gui_camera(2.2, 0.3)]
    boxes = [[-1, 1, -1, 1, -1, 1], [-.6, .9, -.8, .5, -1, .7], [-1.3, 1.3, -1.2, 1.2, -1.4, 1.4]]
    n_hit = 0
    for (M, P), box in zip(cams, boxes):
        g.set_modelView(M)
        g.set_projection(P)
        g.set_box_boundaries(box)
        for frac in (.08, .5):
            res = []
            for skip in (False, True):
                g.set_skipping(skip)
                g.render(maxVal=2 * frac * 60000., method="iso_surface")
                res.append((g.output.copy(), g.output_depth.copy(), g.output_normals.copy(), g.output_alpha.copy(),
                            g.output_occlusion.copy(), g.last_stats()))
            for a, b in zip(res[0][:5], res[1][:5]):
                assert np.array_equal(a, b)
            assert res[1][5][1] < res[0][5][1]
            n_hit += int(np.isfinite(res[0][1]).sum())
    assert n_hit > 5000


def test_iso_tuning_knobs_do_not_change_the_image():
    """Ray segments (1 / 2 / 4 warps share a ray), CTA order, occlusion queue width, the occlusion's tap table, shading in
    the epilogue of the occlusion blur or in its own launch: scheduling only."""
    data = scenes.vol_g(0, np.uint16, seed=3, shape=(90, 120, 100))
    g = _renderer((200, 152))
    g.set_data(data)
    M, P = scenes.gui_camera(0.9, 3.0)
    g.set_modelView(M)
    g.set_projection(P)
    ref = None
    for segments, centre, occ_ctas, table, fused_shading in ((1, 1, 10, 1, 0), (2, 1, 10, 0, 1), (4, 0, 3, 0, 0), (1, 0, 16, 1, 1)):
        for knob, value in ((4, segments), (5, centre), (6, occ_ctas), (17, table), (20, fused_shading)):
            g._check(g._lib.spv_set_tuning(g._ctx, knob, value))
        g.render(maxVal=26000., method="iso_surface")
        planes = [a.copy() for a in (g.output, g.output_alpha, g.output_depth, g.output_normals, g.output_occlusion)]
        if ref is None:
            ref = planes
            assert np.isfinite(planes[2]).sum() > 1000
        for a, b in zip(planes, ref):
            assert np.array_equal(a, b), (segments, centre, occ_ctas, table, fused_shading)
    g._check(g._lib.spv_set_tuning(g._ctx, 6, 10))
    g.close()


def test_occlusion_tap_table_does_not_change_the_image():
    """The ambient-occlusion pass with its taps' pixel offsets read from the per-image table and depths gathered from a
    shared-memory tile (tuning knob 17, default) against the form that hashes every tap in every frame: all planes bit
    for bit, across changes of the radius, the tap count (more than one 32-tap chunk, a radius beyond the table's
    range) and the image size (the table is rebuilt), on images whose extents are not multiples of the 32 x 8 groups."""
    data = scenes.vol_g(0, np.uint16, seed=3, shape=(90, 120, 100))
    M, P = scenes.gui_camera(0.9, 3.0)
    n_hit = 0
    g = _renderer((203, 149))
    g.set_data(data)
    g.set_modelView(M)
    g.set_projection(P)
    for size in ((203, 149), (96, 64)):
        g.resize(size)
        for radius, n_points in ((21, 31), (21, 30), (7, 5), (44, 70), (60, 12), (1, 1)):
            g.set_occ_radius(radius)
            g.set_occ_n_points(n_points)
            res = []
            for table in (1, 0, 1):
                g._check(g._lib.spv_set_tuning(g._ctx, 17, table))
                g.render(maxVal=26000., method="iso_surface")
                res.append([a.copy() for a in (g.output, g.output_alpha, g.output_depth, g.output_normals, g.output_occlusion)])
            for other in res[1:]:
                for a, b in zip(res[0], other):
                    assert np.array_equal(a, b), (size, radius, n_points)
            n_hit += int((res[0][4] > 0).sum())
    assert n_hit > 5000
    g._check(g._lib.spv_set_tuning(g._ctx, 17, 1))
    g.close()


def test_iso_post_passes_on_the_tmu_path(oracle_mod):
    """The texture-unit path skips the occlusion hashing where no surface pixel is within reach and shades only
    surface pixels; both shortcuts must be invisible: recompute occlusion -> blur -> shading with the oracle from
    the GPU's own depth / normal buffers."""
    data = scenes.vol_g(96, np.uint16, seed=3)
    M, P = scenes.gui_camera(0.6, 4.5)     # small object in a large image: most tiles are out of reach
    g = _renderer((320, 256))
    g.set_data(data)
    g.set_modelView(M)
    g.set_projection(P)
    g.set_occ_strength(.4)
    g.set_occ_radius(13)
    g.set_occ_n_points(25)
    g.render(maxVal=40000., method="iso_surface")
    hit = np.isfinite(g.output_depth)
    assert 1000 < hit.sum() < 0.5 * hit.size
    lib = oracle_mod.load("port")
    fp = oracle_mod._fp
    H, W = g.output_depth.shape
    depth = np.ascontiguousarray(g.output_depth)
    occ = np.zeros((H, W), np.float32)
    tmp = np.zeros((H, W), np.float32)
    lib.so_occlusion(fp(occ), W, H, 13, 25, fp(depth))
    lib.so_convolve_scalar(fp(occ), fp(tmp), W, H, 5)
    assert np.mean(np.abs(occ - g.output_occlusion) > 1e-6) < 0.02
    assert _maxdiff(occ, g.output_occlusion) < 0.1
    assert (g.output_occlusion[~hit] >= 0).all()
    out = np.zeros((H, W), np.float32)
    invP, invM = g._invP, g._invM
    lib.so_shading(fp(out), W, H, fp(invP), fp(invM), .4, fp(np.ascontiguousarray(g.output_normals)), fp(depth),
                   fp(np.ascontiguousarray(g.output_occlusion)))
    assert _maxdiff(out, g.output) < 2e-5
    assert (g.output[~hit] == 0).all()


def test_iso_full_pipeline_exact(oracle_mod):
    """iso_surface -> blur(7) -> occlusion -> blur(5) -> shading against the oracle, stage by stage."""
    data = scenes.two_blobs(48)
    M, P = scenes.tilted_camera()
    o = oracle_mod.OracleRenderer((120, 96), kind="port")
    g = _renderer((120, 96), sampler="exact")
    for r in (o, g):
        r.set_data(data)
        r.set_modelView(M)
        r.set_projection(P)
        r.set_occ_strength(.5)
        r.set_occ_radius(15)
        r.set_occ_n_points(40)
        r.render(maxVal=100., method="iso_surface")
    assert np.array_equal(g.output_depth, o.output_depth)
    assert np.array_equal(g.output_alpha, o.output_alpha)
    assert np.array_equal(g.output_normals, o.output_normals)       # blur weights come from the same libm
    assert np.mean(np.abs(g.output_occlusion - o.output_occlusion) > 1e-6) < 0.02
    assert _maxdiff(g.output_occlusion, o.output_occlusion) < 0.1
    assert np.percentile(np.abs(g.output - o.output), 99) < 1e-4
    assert _maxdiff(g.output, o.output) < 0.05


# ---------------------------------------------------------------------------------------------- sort-last slabs
@pytest.mark.parametrize("layout", ["zpair", "3d"])
@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("dtype,peak", [(np.uint16, 60000.), (np.float32, 1.)])
def test_slab_decomposition_is_bit_exact(world, dtype, peak, layout):
    """max over per-slab partial renders == the monolithic render, bitwise (KAT 9); skipping on and off."""
    from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs
    nz = 61
    data = scenes.vol_g(0, dtype, seed=9, shape=(nz, 70, 83))
    M, P = scenes.gui_camera(0.9, 2.7)
    mono = _renderer((136, 104))
    mono.set_layout(layout)
    mono.set_view_copies("primary")  # the slabs hold pairs along z: "bit for bit" is against the single-GPU render through the same (z) copy
    mono.set_data(data)
    mono.set_modelView(M)
    mono.set_projection(P)
    mono.render(maxVal=peak)
    for skip in (True, False):
        raws = []
        for rank in range(world):
            s = SlabMaxProjector((136, 104), rank=rank, world=world)
            s.set_layout(layout)
            s.set_skipping(skip)
            s.set_data(data)
            s.set_modelView(M)
            s.set_projection(P)
            s.set_max_val(peak)
            s.render()
            raw = np.zeros((104, 136), np.float32)
            s._check(s._lib.spv_read(s._ctx, 5, raw.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)), raw.size))
            raws.append(raw)
            alpha = s.output_alpha.copy()
            s.close()
        comp = np.maximum.reduce(raws)
        hit = comp >= 0
        expect = np.where(hit, np.clip(np.maximum(comp, 0) / np.float32(peak), 0, 1), 0).astype(np.float32)
        assert np.array_equal(expect, mono.output), "world=%d skip=%s max |d| = %g" % (world, skip, _maxdiff(expect, mono.output))
        assert np.array_equal(alpha, mono.output_alpha)


def test_slab_against_oracle_partials(oracle_mod):
    """Each slab's partial image owns exactly the samples the specification (oracle so_max_project_raw) says."""
    from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs
    import ctypes as C
    data = scenes.vol_g(0, np.uint16, seed=3, shape=(40, 48, 56))
    M, P = scenes.gui_camera(2.2, 3.)
    o = oracle_mod.OracleRenderer((96, 80), kind="port", pos_mode=2, weight_bits=8)
    o.set_data(data)
    o.set_modelView(M)
    o.set_projection(P)
    for rank, (z0, z1) in enumerate(partition_slabs(40, 3)):
        s = SlabMaxProjector((96, 80), rank=rank, world=3)
        s.set_data(data)
        s.set_modelView(M)
        s.set_projection(P)
        s.render(maxVal=60000.)
        raw = np.zeros((80, 96), np.float32)
        s._check(s._lib.spv_read(s._ctx, 5, raw.ctypes.data_as(C.POINTER(C.c_float)), raw.size))
        want = o.render_raw(z0, z1)
        assert np.array_equal(raw < 0, want < 0)
        assert _maxdiff(raw, want) < 60000. * 2e-3
        # a pixel is zero in one iff it is zero in the other: same ownership of samples
        assert np.mean((raw == 0) != (want == 0)) < 0.01
        s.close()
