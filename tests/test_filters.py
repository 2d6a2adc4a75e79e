"""Volume filters: the oracle's restatement of gputools.convolve_sep3 against analytic known answers
and scipy; the processor classes against the reference's interface (spimagine/models/imageprocessor.py); and, on the
GPU, libspimcuda's spv_filter_* against the oracle bit for bit."""
import numpy as np
import pytest

import scenes


@pytest.fixture(scope="module")
def forc():
    from oracle import filters
    filters.build()
    return filters


ASYM = (np.array([.1, .2, .3, .4, .5]), np.array([1., -2., 3.]), np.array([.5, .25, .125, .0625, .03125, .015625, 1.]))


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_impulse_response_is_the_outer_product_of_the_taps(forc):
    hx, hy, hz = ASYM
    d = np.zeros((21, 19, 23), np.float32)
    c = (10, 9, 11)
    d[c] = 1.
    r = forc.convolve_sep3(d, hx, hy, hz)
    # out[i] = sum h[ht] in[i + Nh/2 - ht]  =>  an impulse at c puts tap ht at c + ht - Nh/2 (true convolution)
    want = np.zeros_like(d)
    for a, ha in enumerate(hz):
        for b, hb in enumerate(hy):
            for e, he in enumerate(hx):
                want[c[0] + a - len(hz) // 2, c[1] + b - len(hy) // 2, c[2] + e - len(hx) // 2] = \
                    np.float32(ha) * np.float32(hb) * np.float32(he)
    assert np.allclose(r, want, rtol=1e-6, atol=0)
    assert np.count_nonzero(r) == len(hx) * len(hy) * len(hz)


def test_oracle_constant_volume_gives_partial_tap_sums_at_the_faces(forc):
    h = forc.gauss_taps(2.)  # 11 taps, sums to 1
    d = np.full((16, 16, 16), 3., np.float32)
    r = forc.convolve_sep3(d, h, h, h)
    assert abs(r[8, 8, 8] - 3.) < 1e-5  # the whole kernel is inside
    cs = np.cumsum(h)
    # x = 0: taps reading in[0 + 5 - ht] with ht <= 5 stay inside -> sum(h[:6])
    assert abs(r[8, 8, 0] - 3. * cs[5]) < 1e-5
    assert abs(r[8, 0, 0] - 3. * cs[5] ** 2) < 1e-5
    assert abs(r[0, 0, 0] - 3. * cs[5] ** 3) < 1e-5
    assert abs(r[8, 8, 15] - 3. * (1 - cs[4])) < 1e-5  # ht >= 5 stay inside at the far face


@pytest.mark.parametrize("taps", [(3, 3, 3), (11, 7, 19), (4, 6, 2), (1, 1, 5), (65, 3, 3)])
def test_oracle_equals_scipy_zero_padded_convolution(forc, taps):
    from scipy.ndimage import convolve1d
    rng = np.random.default_rng(sum(taps))
    d = rng.random((14, 17, 70), dtype=np.float32)
    hs = [rng.random(n) - .3 for n in taps]
    r = forc.convolve_sep3(d, *hs)
    w = d.astype(np.float64)
    for axis, h in zip((2, 1, 0), hs):
        w = convolve1d(w, np.asarray(h, np.float32).astype(np.float64), axis=axis, mode="constant")
    assert np.abs(r - w).max() < 2e-5 * max(1., np.abs(w).max())


def test_oracle_fused_and_unfused_accumulation_differ_by_rounding_only(forc):
    d = scenes.vol_g(48, np.uint16)
    h = forc.gauss_taps(4.)
    a, b = forc.convolve_sep3(d, h, h, h, fused=True), forc.convolve_sep3(d, h, h, h, fused=False)
    assert np.abs(a - b).max() <= 2e-6 * a.max()


def test_blur_processor_taps_follow_the_reference_formula():
    from spimagine_b200 import imageprocessor as ip
    # models/imageprocessor.py:52-55: N = 2 sigma + 1; x = arange(-N, N + 1); h = exp(-x^2 / 2 / sigma^2); h /= sum(h)
    h = ip.BlurProcessor(sigma=4.)._taps()[0]
    assert len(h) == 19 and abs(h.sum() - 1) < 1e-12 and np.argmax(h) == 9
    assert np.allclose(h[9] / h[8], np.exp(1 / 32.))
    hx, hy, hz = ip.BlurXYZProcessor(sx=1., sy=2., sz=3.)._taps()
    assert (len(hx), len(hy), len(hz)) == (7, 11, 15)


def test_processor_interface_matches_the_reference():
    from spimagine_b200 import imageprocessor as ip
    p = ip.BlurProcessor(sigma=2.)
    assert p.name == "blur" and p.kwargs == {"sigma": 2.} and p.sigma == 2.
    p.set_params(sigma=3.)
    assert p.sigma == 3.
    with pytest.raises(AttributeError):
        p.nothing
    assert ip.BlurXYZProcessor().kwargs == {"sx": 4., "sy": 4., "sz": 4.}
    d = np.arange(24.).reshape(2, 3, 4)
    assert ip.CopyProcessor().apply(d) is d and ip.CopyProcessor().name == "copy"
    assert ip.LucyRichProcessor().apply(d) is d and ip.LucyRichProcessor().name == "RL-Deconv"
    n = ip.NoiseProcessor(sigma=1).apply(d)
    assert n.shape == d.shape and n.min() >= 0
    f = ip.FuncProcessor(lambda data, para: data * para, "myfunc", para=.5)
    assert f.name == "myfunc" and np.array_equal(f.apply(d), d * .5)
    with pytest.raises(NotImplementedError):
        ip.ImageProcessor("x").apply(d)


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_sep3(data, hx, hy, hz, fuse=1):
    from spimagine_b200 import imageprocessor as ip
    vf = ip._shared_filter(0)
    vf.set_tuning(0, fuse)
    try:
        return ip.convolve_sep3(data, hx, hy, hz)
    finally:
        vf.set_tuning(0, 0)


@pytest.mark.gpu
def test_gpu_per_pass_times_add_up():
    """spv_filter_last_pass_ms: three kernels (x, y, z) or two with the fused x + y pass; their times add up to the
    time of the whole convolution (events on one stream, back to back)"""
    from spimagine_b200 import imageprocessor as ip
    vf = ip.VolumeFilter(0)
    try:
        data = scenes.random_vol((40, 64, 128), np.uint16, seed=5)
        taps = [np.full(n, 1. / n) for n in (19, 19, 19)]
        for fuse, n in ((0, 3), (1, 2)):
            vf.set_tuning(0, fuse)
            vf.load(data)
            vf.convolve_sep3(*taps)
            vf.sync()
            ms = vf.last_pass_ms()
            assert len(ms) == n and all(m > 0 for m in ms)
            assert abs(sum(ms) - vf.last_ms()) < 0.02 * vf.last_ms() + 1e-3
    finally:
        vf.set_tuning(0, 0)
        vf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("x_pairs,axis", [(0, 16), (1, 32), (2, 1602), (2, 1604), (0, 1604)])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
def test_gpu_kernel_variants_equal_the_oracle_bitwise(forc, dtype, x_pairs, axis):
    """every x / axis kernel variant (single rows or row pairs, one / two / four columns per thread, packed
    fma.rn.f32x2) gives the oracle's bits: volumes with several 64-row tiles per CTA, cut tiles on every face, rows of
    whole words (132, 8) and not (the kernels fall back: 131 as uint8 / uint16)"""
    from spimagine_b200 import imageprocessor as ip
    vf = ip._shared_filter(0)
    try:
        for shape, taps in (((9, 150, 132), (19, 7, 31)), ((70, 3, 8), (5, 19, 3)), ((6, 40, 131), (11, 19, 19)),
                            ((2, 700, 260), (47, 3, 1))):
            rng = np.random.default_rng(sum(shape) + sum(taps))
            data = scenes.random_vol(shape, dtype, seed=sum(shape))
            hs = [rng.random(n) - .2 for n in taps]
            want = forc.convolve_sep3(data, *hs)
            vf.set_tuning(1, axis)
            vf.set_tuning(2, x_pairs)
            got = ip.convolve_sep3(data, *hs)
            assert np.array_equal(got, want), (shape, taps)
    finally:
        vf.set_tuning(1, 1)
        vf.set_tuning(2, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [1, 0])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("shape,taps", [((40, 50, 300), (19, 19, 19)),   # several x tiles, interior + face chunks
                                        ((3, 100, 130), (17, 19, 5)),    # fused x + y, padded, several y tiles
                                        ((5, 70, 260), (31, 29, 3)),     # fused, the largest class
                                        ((4, 64, 128), (3, 2, 7)),       # fused, the smallest class
                                        ((33, 17, 129), (7, 11, 3)),     # ragged extents, exact instantiations
                                        ((20, 35, 64), (5, 13, 17)),     # padded instantiations (5 -> 7, 13 -> 15, 17 -> 19)
                                        ((9, 6, 5), (19, 19, 19)),       # volume smaller than the kernel on every axis
                                        ((18, 20, 40), (4, 6, 2)),       # even tap counts
                                        ((24, 8, 33), (1, 1, 63)),
                                        ((70, 5, 6), (3, 3, 101))])      # beyond the unrolled sizes: generic kernel
def test_gpu_convolve_sep3_equals_the_oracle_bitwise(forc, dtype, shape, taps, fuse):
    rng = np.random.default_rng(len(shape) + sum(taps))
    data = scenes.random_vol(shape, dtype, seed=sum(shape))
    hs = [rng.random(n) - .2 for n in taps]
    got = _gpu_sep3(data, *hs, fuse=fuse)
    want = forc.convolve_sep3(data, *hs)
    assert got.dtype == np.float32 and got.shape == shape
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [16, 32, 2, 4])
def test_gpu_axis_pass_variants_are_identical(forc, variant):
    from spimagine_b200 import imageprocessor as ip
    rng = np.random.default_rng(variant)
    data = scenes.random_vol((45, 70, 136), np.float32, seed=9)
    hs = [rng.random(n) - .2 for n in (3, 19, 13)]
    vf = ip._shared_filter(0)
    vf.set_tuning(1, variant)
    try:
        got = ip.convolve_sep3(data, *hs)
    finally:
        vf.set_tuning(1, 1)
    assert np.array_equal(got, forc.convolve_sep3(data, *hs))


@pytest.mark.gpu
def test_gpu_generic_x_pass_and_other_element_types(forc):
    rng = np.random.default_rng(5)
    hs = [rng.random(n) for n in (77, 3, 5)]  # x pass beyond the unrolled sizes
    for dt in (np.int16, np.float64, np.int32):
        data = (rng.random((12, 30, 90)) * 1000 - 300).astype(dt)
        assert np.array_equal(_gpu_sep3(data, *hs), forc.convolve_sep3(data.astype(np.float32), *hs))


@pytest.mark.gpu
def test_gpu_non_finite_voxels_spread_exactly_as_in_the_reference(forc):
    data = scenes.random_vol((30, 40, 200), np.float32, seed=3)
    data[15, 20, 100] = np.nan
    data[3, 2, 7] = np.inf
    for taps in ((5, 13, 17), (16, 17, 18)):  # padded instantiations: three passes / fused x + y
        hs = [np.abs(np.random.default_rng(1).random(n)) + .1 for n in taps]
        got, want = _gpu_sep3(data, *hs), forc.convolve_sep3(data, *hs)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        ok = np.isfinite(want)
        assert np.array_equal(got[ok], want[ok])


@pytest.mark.gpu
def test_gpu_blur_processors_and_chains(forc):
    from spimagine_b200 import imageprocessor as ip
    data = scenes.vol_g(64, np.uint16, seed=3)
    h4, h2 = forc.gauss_taps(4.), forc.gauss_taps(2.)
    assert np.array_equal(ip.BlurProcessor().apply(data), forc.convolve_sep3(data, h4, h4, h4))
    hx, hy, hz = forc.gauss_taps(1.), forc.gauss_taps(2.), forc.gauss_taps(3.)
    assert np.array_equal(ip.BlurXYZProcessor(1., 2., 3.).apply(data), forc.convolve_sep3(data, hx, hy, hz))
    # a chain on the device (second convolution reads the first result) == the reference's host chain
    vf = ip.VolumeFilter()
    vf.load(data)
    ip.BlurProcessor(2.).apply_device(vf)
    ip.BlurXYZProcessor(1., 2., 3.).apply_device(vf)
    want = forc.convolve_sep3(forc.convolve_sep3(data, h2, h2, h2), hx, hy, hz)
    assert np.array_equal(vf.result(), want)
    assert vf.last_ms() > 0 and vf.launch_count() == 6  # two convolutions of three passes each
    vf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_gpu_filtered_volume_reaches_the_renderer_without_the_host(forc, dtype):
    """apply_chain == the reference's flow (gui/mainwidget.py:455-465): data = proc.apply(data) for every processor,
    then renderer.update_data(data) (which casts to the renderer's element type)."""
    from spimagine_b200 import VolumeRenderer, imageprocessor as ip
    data = scenes.vol_g(64, dtype, seed=2)
    peak = float(data.max())
    M, P = scenes.gui_camera(0.5, 3.5)
    procs = [ip.CopyProcessor(), ip.BlurProcessor(2.), ip.LucyRichProcessor()]
    a = VolumeRenderer((160, 120))
    a.set_data(data)
    a.set_modelView(M); a.set_projection(P)
    ms = ip.apply_chain(a, data, procs)
    assert ms > 0
    a.render(maxVal=peak)
    h = forc.gauss_taps(2.)
    blurred = forc.convolve_sep3(data, h, h, h)
    b = VolumeRenderer((160, 120))
    b.set_data(data)
    b.set_modelView(M); b.set_projection(P)
    b.update_data(blurred)  # host path: astype(self.dtype) then upload
    b.render(maxVal=peak)
    assert np.array_equal(a.output, b.output) and np.array_equal(a.output_alpha, b.output_alpha)
    assert a.data_min_max == b.data_min_max
    # a host-only processor in the middle of the chain
    procs = [ip.BlurProcessor(1.), ip.FuncProcessor(lambda d: d * 0.5), ip.BlurProcessor(1.)]
    ip.apply_chain(a, data, procs)
    a.render(maxVal=peak)
    h1 = forc.gauss_taps(1.)
    want = forc.convolve_sep3(forc.convolve_sep3(data, h1, h1, h1) * 0.5, h1, h1, h1)
    b.update_data(want)
    b.render(maxVal=peak)
    assert np.array_equal(a.output, b.output)
    a.close(); b.close()


@pytest.mark.gpu
def test_gpu_filter_errors():
    from spimagine_b200 import imageprocessor as ip, _lib
    vf = ip.VolumeFilter()
    with pytest.raises(_lib.SpvError):
        vf._check(vf._lib.spv_filter_convolve_sep3(vf._f, None, 1, None, 1, None, 1))  # nothing loaded
    vf.load(np.zeros((4, 4, 4), np.float32))
    with pytest.raises(_lib.SpvError):
        vf.convolve_sep3(np.ones(2000), [1.], [1.])
    with pytest.raises(ValueError):
        vf.load(np.zeros((4, 4), np.float32))
    vf.close()


def test_processors_against_the_references_own_classes(forc):
    """tests/golden/processors_ref.json: the reference's imageprocessor.py executed with gputools replaced by a
    recorder (tests/golden/make_processor_golden.py).  Taps are float64-exact; the spectrum expression agrees with the
    restatement the GPU tests use; names, kwargs and identity behaviour are the reference's."""
    import json
    import os
    from spimagine_b200 import imageprocessor as ip
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "processors_ref.json")) as f:
        ref = json.load(f)
    for key, want in ref["taps"].items():
        if key.startswith("blur_xyz"):
            got = ip.BlurXYZProcessor(sx=1., sy=2., sz=3.)._taps()
        else:
            got = ip.BlurProcessor(sigma=float(key.split("_")[1]))._taps()
        assert len(got) == 3
        for g, w in zip(got, want):
            assert np.array_equal(np.asarray(g, np.float64), np.array(w)), key
    data = np.array(ref["fft"]["data"], dtype=ref["fft"]["dtype"])
    for log in (False, True):
        want = np.array(ref["fft"]["log_%s" % log]["value"])
        got = forc.fft_spectrum(data, log=log)
        assert got.shape == want.shape
        # the reference's product with the float64 scalar 1/sqrt(size) is float64 under NEP 50, float32 before
        assert np.allclose(got, want, rtol=3e-7, atol=1e-6 * np.abs(want).max())
    made = {"copy": ip.CopyProcessor(), "blur": ip.BlurProcessor(), "blur_xyz": ip.BlurXYZProcessor(),
            "noise": ip.NoiseProcessor(), "fft": ip.FFTProcessor(), "lucy": ip.LucyRichProcessor(),
            "func": ip.FuncProcessor(lambda d, k=2: d * k, "times", k=3)}
    for key, want in ref["interface"].items():
        assert made[key].name == want["name"] and made[key].kwargs == want["kwargs"], key
        for k, v in want["kwargs"].items():
            assert getattr(made[key], k) == v
    vol = np.zeros((2, 2, 2), np.float32)
    assert (ip.CopyProcessor().apply(vol) is vol) == ref["identity"]["copy"]
    assert (ip.LucyRichProcessor().apply(vol) is vol) == ref["identity"]["lucy"]
    assert float(made["func"].apply(np.ones(1))[0]) == ref["identity"]["func"]
