"""Pins the CPU oracle (oracle/spim_oracle.c, a C restatement of the reference's OpenCL kernels):
  * against the committed golden vectors, which were rendered with the reference's own kernel text
    (tests/golden/make_golden.py -> oracle/_ref), bit for bit;
  * against that reference build directly when it is present (this container; it also travels to the GPU box);
  * against analytic known answers that do not depend on any implementation.
The oracle is test infrastructure; nothing in spimagine_b200/ uses it."""
import os

import numpy as np
import pytest

import golden_cases
import scenes
from spimagine_b200.utils.transform_matrices import mat4_ortho, mat4_perspective, mat4_rotation, mat4_translate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", sorted(golden_cases.CASES))
def test_port_matches_golden_bitwise(oracle_mod, name):
    case = golden_cases.CASES[name]
    rend = oracle_mod.OracleRenderer(golden_cases.SIZE, interpolation=case.get("interpolation", "linear"),
                                     kind="port")
    res = golden_cases.run_case(rend, name)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert sorted(gold.files) == sorted(res)
    for k in gold.files:
        assert _eq(res[k], gold[k]), "%s/%s differs from the reference build (max |d| = %g)" % (
            name, k, np.nanmax(np.abs(np.where(np.isfinite(gold[k]), res[k] - gold[k], 0))))


@pytest.mark.parametrize("name", sorted(golden_cases.CONFIG_CASES))
def test_port_matches_config_golden_bitwise(oracle_mod, name):
    """BASELINE.json configs[0] at its own size (128^3 float32 -> 512x512)."""
    case = golden_cases.CONFIG_CASES[name]
    res = golden_cases.run_config_case(oracle_mod.OracleRenderer(case["size"], kind="port"), name)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    for k in gold.files:
        assert _eq(res[k], gold[k]), k


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("interp", ["linear", "nearest"])
def test_port_matches_reference_build_on_random_volumes(oracle_mod, have_ref, dtype, interp):
    if not have_ref:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    data = scenes.random_vol((17, 23, 29), dtype, seed=5)
    M, P = scenes.gui_camera(0.8, 3.3)
    outs = []
    for kind in ("port", "reference"):
        r = oracle_mod.OracleRenderer((50, 38), interpolation=interp, kind=kind)
        r.set_data(data)
        r.set_units([1., .8, 2.])
        r.set_modelView(M)
        r.set_projection(P)
        r.set_alpha_pow(.3)
        r.render(maxVal=float(data.max()), minVal=1.)
        a = (r.output.copy(), r.output_alpha.copy())
        r.set_alpha_pow(0.)
        r.render(maxVal=float(data.max()) * .6, method="iso_surface")
        outs.append(a + (r.output.copy(), r.output_depth.copy(), r.output_normals.copy(), r.output_occlusion.copy()))
    for a, b in zip(*outs):
        assert _eq(a, b)


def test_constant_volume_known_answer(oracle_mod):
    """tests/test_rendering/test_simple_rendering.py:85-95: every hit pixel is 123/200; misses are 0 / alpha -1."""
    r = oracle_mod.OracleRenderer((31, 29))
    r.set_data(np.full((9, 9, 9), 123., np.float32))
    M, P = scenes.gui_camera(0.3, 4.)
    r.set_modelView(M)
    r.set_projection(P)
    r.render(maxVal=200.)
    hit = r.output_alpha > 0
    assert 0 < hit.sum() < hit.size
    # the eight fp32 weights of the spec's trilinear sum do not add up to exactly 1: a few ulp around 0.615
    np.testing.assert_allclose(r.output[hit], 123. / 200., rtol=1e-6)
    np.testing.assert_array_equal(r.output[~hit], 0.)
    np.testing.assert_array_equal(r.output_alpha[~hit], -1.)
    assert r.count_hit_rays() == hit.sum()
    # integer volume: alpha is tnear on hits and 0 on misses
    r.set_data(np.full((9, 9, 9), 123, np.uint16))
    r.render(maxVal=200., gamma=2.)
    np.testing.assert_allclose(r.output[hit], (123. / 200.) ** 2, rtol=1e-6)
    assert (r.output_alpha[hit] > 2.).all() and (r.output_alpha[~hit] == 0).all()


def test_camera_looking_away_misses_everything(oracle_mod):
    r = oracle_mod.OracleRenderer((16, 16))
    r.set_data(scenes.gaussian(16))
    r.set_modelView(mat4_translate(0, 0, 4.))  # volume behind the camera... rays still hit going backwards?
    r.set_projection(mat4_perspective(60, 1., .1, 10))
    r.set_modelView(mat4_translate(30., 0, -4.))  # far off to the side: outside the frustum
    r.render(maxVal=200.)
    assert (r.output == 0).all() and (r.output_alpha == -1).all()


def test_ortho_ramp_known_answer(oracle_mod):
    """Orthographic view straight down z onto a volume that is a ramp along z: the ray maximum is the far-face
    value wherever the ray is inside the box; linear in x for a ramp along x."""
    N = 16
    z = np.linspace(0, 1, N, dtype=np.float32)
    vol = np.broadcast_to(z[:, None, None], (N, N, N)).copy()
    r = oracle_mod.OracleRenderer((32, 32))
    r.set_data(vol)
    r.set_projection(mat4_ortho(-1.3, 1.3, -1.3, 1.3, -3, 3))
    r.set_modelView(np.dot(mat4_translate(0, 0, 0), mat4_rotation(1e-3, 0, 1, 0)))
    r.render(maxVal=1.)
    hit = r.output_alpha > 0
    assert hit.sum() > 300
    inner = np.zeros_like(hit)
    inner[10:22, 10:22] = True
    np.testing.assert_allclose(r.output[inner & hit], 1., atol=1e-6)


def test_nearest_outputs_are_voxel_values(oracle_mod):
    data = scenes.random_vol((12, 13, 14), np.uint16, 3)
    r = oracle_mod.OracleRenderer((40, 40), interpolation="nearest")
    r.set_data(data)
    M, P = scenes.tilted_camera()
    r.set_modelView(M)
    r.set_projection(P)
    r.render(maxVal=0.)  # maxVal == 0: no window, clamp to [0,1] only ... so use the raw render
    raw = r.render_raw()
    vals = set(np.unique(data).astype(np.float32).tolist()) | {-1.0, 0.0}
    assert set(np.unique(raw).tolist()) <= vals


def test_multipass_parts_are_monotone(oracle_mod):
    r = oracle_mod.OracleRenderer((40, 32))
    r.set_data(scenes.two_blobs(24))
    M, P = scenes.tilted_camera()
    r.set_modelView(M)
    r.set_projection(P)
    prev = None
    for part in range(4):
        r.render(maxVal=255., numParts=4, currentPart=part)
        if prev is not None:
            assert (r.output >= prev).all()
        prev = r.output.copy()


def test_lcg_hash_table(oracle_mod):
    """utils.cl:10-24 in exact uint32 arithmetic, checked against Python integers."""
    lib = oracle_mod.load("port")

    def ref(x, y):
        a = (4421 + (1 + x) * (1 + y) + x + y) & 0xffffffff
        for _ in range(10):
            a = ((1664525 * a + 1013904223) & 0xffffffff) % 79197919
        return a

    for x, y in [(0, 0), (1, 2), (200, 200), (1023, 7), (65535, 65535), (4000000000, 3), (123456, 654321),
                 (0xffffffff, 0xffffffff), (31, 961), (999, 97700), (5, 0), (0, 5), (77777, 1), (2, 3), (3, 2),
                 (1000, 1000)]:
        assert lib.so_lcg_hash(x, y) == ref(x, y)
        assert lib.so_random(x, y) == np.float32(ref(x, y)) / np.float32(79197919)


def test_blur_impulse_response_is_the_asymmetric_table(oracle_mod):
    """convolve_2d.cl: weights exp(coef*(ht-Nh/2.f)^2/Nh^2) sit at integer offsets ht-Nh/2: centred at Nh/2.f but
    applied around Nh//2, so the response to an impulse is not symmetric."""
    lib = oracle_mod.load("port")
    W = H = 21
    for nh, coef, fn in ((5, -10., lib.so_convolve_scalar), (7, -5., None)):
        w = np.exp(np.float32(coef) * (np.arange(nh, dtype=np.float32) - np.float32(nh / 2.)) ** 2 / nh / nh)
        w = (w / w.sum()).astype(np.float32)
        if fn is None:
            buf = np.zeros((H, W, 3), np.float32)
            buf[10, 10] = (1, 2, 3)
            tmp = np.zeros_like(buf)
            lib.so_convolve_vec(oracle_mod._fp(buf), oracle_mod._fp(tmp), W, H, nh)
            got = buf[10, :, 0]
        else:
            buf = np.zeros((H, W), np.float32)
            buf[10, 10] = 1
            tmp = np.zeros_like(buf)
            fn(oracle_mod._fp(buf), oracle_mod._fp(tmp), W, H, nh)
            got = buf[10, :]
        # output pixel i reads input i + (ht - nh//2): the impulse at 10 reaches i = 10 - (ht - nh//2)
        expect = np.zeros(W, np.float32)
        for ht in range(nh):
            expect[10 - (ht - nh // 2)] = w[ht] * w[nh // 2]
        np.testing.assert_allclose(got, expect, rtol=2e-6, atol=1e-8)
        assert abs(got[9] - got[11]) > 1e-3  # asymmetric


def test_sphere_iso_surface_known_answer(oracle_mod):
    """tests/test_rendering/test_simple_rendering.py:55-68: 900*exp(-10 R) thresholded at 10 is a sphere of radius
    ln(90)/10 = 0.45; seen from z = -5 the centre pixel's depth is 4.9 - 0.45 plus at most one coarse and one fine
    step, its normal points back at the camera."""
    N = 64
    r = oracle_mod.OracleRenderer((65, 65))
    r.set_data(scenes.iso_sphere(N))
    r.set_modelView(mat4_translate(0, 0, -5))
    r.set_projection(mat4_perspective())
    r.render(maxVal=20., method="iso_surface_raw")
    d = r.output_depth[32, 32]
    dt = 2. / 199
    # t runs from the near plane (0.1 in front of the eye)
    assert 4.9 - .45 - 2 * dt < d < 4.9 - .45 + 2 * dt + .05
    n = r.output_normals[32, 32]
    assert abs(np.linalg.norm(n) - 1) < 1e-5 and abs(n[2]) > .99
    hit = np.isfinite(r.output_depth)
    # visible disc radius in pixels: r_pix = 0.45/sqrt(25-0.45^2) / tan(22.5 deg) * 32.5
    rp = .45 / np.sqrt(25 - .45 ** 2) / np.tan(np.pi / 8) * 32.5
    assert abs(np.sqrt(hit.sum() / np.pi) - rp) < 1.5
    assert (r.output[~hit] == 0).all() and (r.output_alpha[~hit] == 0).all()


def test_position_modes_agree_within_drift(oracle_mod):
    """pos += delta (reference) vs pos0 + k*delta (what a sample-skipping kernel needs): rounding drift only."""
    data = scenes.two_blobs(48)
    M, P = scenes.tilted_camera()
    outs = []
    for pm in (0, 1, 2):
        r = oracle_mod.OracleRenderer((64, 48), pos_mode=pm)
        r.set_data(data)
        r.set_modelView(M)
        r.set_projection(P)
        r.render(maxVal=255.)
        outs.append(r.output.copy())
    assert np.abs(outs[0] - outs[1]).max() < 2e-4
    assert np.abs(outs[1] - outs[2]).max() < 2e-4


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sort_last_partition_is_bit_exact(oracle_mod, world):
    from spimagine_b200.multigpu import partition_slabs
    data = scenes.vol_g(40, np.uint16, seed=2)
    M, P = scenes.gui_camera(1.1, 3.)
    r = oracle_mod.OracleRenderer((48, 40), pos_mode=2, weight_bits=8)
    r.set_data(data)
    r.set_modelView(M)
    r.set_projection(P)
    full = r.render_raw()
    parts = [r.render_raw(z0, z1) for z0, z1 in partition_slabs(40, world)]
    np.testing.assert_array_equal(np.maximum.reduce(parts), full)


def test_display_restatement_known_answers():
    """oracle.display_rgba8 (texture.frag:8-38): LUT end points, linear LUT filtering between texel centres, alpha
    = value, transparent misses, inverted look-up in white mode."""
    from oracle import oracle
    lut = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]], np.float32)          # N = 4: texel centres at 1/8, 3/8, ...
    v = np.array([[0., 1 / 8, 2 / 8, 3 / 8, 1., 2., -1., 0.5]], np.float32)
    a = np.array([[1., 1., 1., 1., 1., 1., 1., -1.]], np.float32)
    img = oracle.display_rgba8(v, a, lut)
    assert img.shape == (1, 8, 4) and img.dtype == np.uint8
    assert img[0, 0].tolist() == [0, 0, 0, 0]            # below the first texel centre: clamp to edge
    assert img[0, 1].tolist() == [0, 0, 0, 32]           # exactly texel 0
    assert img[0, 2].tolist() == [128, 0, 0, 64]         # half way between texels 0 and 1
    assert img[0, 3].tolist() == [255, 0, 0, 96]         # texel 1
    assert img[0, 4].tolist() == [255, 255, 255, 255]
    assert img[0, 5].tolist() == [255, 255, 255, 255]    # values are clamped like a unorm texture
    assert img[0, 6].tolist() == [0, 0, 0, 0]
    assert img[0, 7].tolist() == [0, 0, 0, 0]            # tnear < 0: transparent
    white = oracle.display_rgba8(v, a, lut, mode_black=False)
    assert white[0, 0].tolist() == [255, 255, 255, 0] and white[0, 4].tolist() == [0, 0, 0, 255]
