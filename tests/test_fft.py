"""FFTProcessor: libspimfft.so against the numpy restatement of the reference expression
(models/imageprocessor.py:82-98 with gputools' pad helpers restated, oracle/filters.py).  CPU tests pin the
restatement with analytic answers and check the ABI; GPU tests are the parity tests."""
import ctypes
import os
import re

import numpy as np
import pytest

import scenes
from oracle import filters as forc
from spimagine_b200 import imageprocessor as ip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ restatement, ABI (CPU)
def test_pad_helpers_known_answers():
    a6, a8, a5 = np.arange(6), np.arange(8), np.arange(5)
    assert forc.pad_to_shape(a6, (8,), "wrap").tolist() == [5, 0, 1, 2, 3, 4, 5, 0]
    assert forc.pad_to_shape(a5, (8,), "wrap").tolist() == [3, 4, 0, 1, 2, 3, 4, 0]      # 2 in front, 1 behind
    assert forc.pad_to_shape(a8, (6,)).tolist() == [1, 2, 3, 4, 5, 6]
    assert forc.pad_to_shape(a8, (5,)).tolist() == [1, 2, 3, 4, 5]                       # 1 dropped in front, 2 behind
    assert forc.pad_to_shape(a8, (8,)) is a8
    assert forc.pad_to_power2(np.zeros((3, 8, 5))).shape == (4, 8, 8)
    assert forc.pad_to_power2(np.zeros((1, 2, 4))).shape == (1, 2, 4)


def test_spectrum_restatement_known_answers():
    n = 16
    const = np.full((n, n, n), 3., np.float32)
    s = forc.fft_spectrum(const)
    assert s.dtype == np.float32 and s.shape == const.shape
    want = np.zeros_like(s)
    want[n // 2, n // 2, n // 2] = 3. * np.sqrt(n ** 3)
    assert np.allclose(s, want, atol=1e-3)
    z, y, x = np.meshgrid(*(np.arange(n),) * 3, indexing="ij")
    wave = np.cos(2 * np.pi * (2 * x + 3 * y + 1 * z) / n).astype(np.float32)
    s = forc.fft_spectrum(wave)
    peaks = np.argwhere(s > 1.)
    assert sorted(map(tuple, peaks)) == [(n // 2 - 1, n // 2 - 3, n // 2 - 2), (n // 2 + 1, n // 2 + 3, n // 2 + 2)]
    assert np.allclose(s[tuple(peaks[0])], np.sqrt(n ** 3) / 2, rtol=1e-5)
    assert np.allclose(forc.fft_spectrum(wave, log=True), np.log2(0.001 + s))
    odd = scenes.random_vol((5, 6, 7), np.uint16, seed=1)
    s = forc.fft_spectrum(odd)
    assert s.shape == odd.shape
    # Parseval on the padded volume, restricted to the crop: the DC coefficient survives the crop at P/2 - floor(d/2)
    padded = forc.pad_to_power2(odd.astype(np.float64), "wrap")
    assert np.isclose(s[4 - 1, 4 - 1, 4 - 0], padded.sum() / np.sqrt(padded.size), rtol=1e-5)


def test_fft_library_exports_what_the_header_declares():
    text = open(os.path.join(ROOT, "include", "spimfft.h")).read()
    names = sorted(set(re.findall(r"SPF_API[^;(]*?\b(spf_[a-z0-9_]+)\s*\(", text)))
    assert len(names) == 9
    assert os.path.exists(ip.FFT_LIB_PATH), "build with python -m spimagine_b200.build"
    lib = ctypes.CDLL(ip.FFT_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libspimfft.so does not export %s" % n
    assert set(ip.load_fft()._signatures) == set(names)


def test_fft_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from spimagine_b200 import _lib
    with pytest.raises(_lib.SpvError):
        ip.FFTProcessor().apply(np.zeros((4, 4, 4), np.float32))


# ------------------------------------------------------------------ parity (GPU)
def _tol(data, want=0.):
    """What two float32 FFT libraries can differ by, per element:
      4e-6 * rms(data)      eps * log2(N) * rms: the rounding of the transform, spread over the coefficients;
      1e-6 * |coefficient|  a few ulp of the coefficient itself;
      1e-7 * peak           one ulp of the largest coefficient (the DC term, sqrt(N) * mean): along the axes through
                            DC a coefficient is what is left after plane sums of DC size cancel, so its error is an
                            ulp of THOSE, not of the result (measured: 2.3e-8 of the peak on random uint16 data)."""
    want = np.abs(np.asarray(want, np.float64))
    return 4e-6 * float(np.sqrt(np.mean(np.asarray(data, np.float64) ** 2))) + 1e-6 * want + 1e-7 * float(want.max())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((32, 32, 32), np.float32), ((16, 64, 32), np.uint16), ((5, 6, 7), np.uint16),
                                         ((33, 20, 65), np.float32), ((1, 9, 4), np.uint8), ((7, 1, 1), np.int16),
                                         ((48, 100, 130), np.uint16)])
def test_spectrum_matches_the_restatement(shape, dtype):
    data = scenes.random_vol(shape, dtype, seed=3)
    want = forc.fft_spectrum(data, precise=True)
    got = ip.FFTProcessor().apply(data)
    assert got.dtype == np.float32 and got.shape == data.shape
    excess = np.abs(got - want) - _tol(data, want)
    assert excess.max() <= 0, (float(excess.max()), np.unravel_index(excess.argmax(), excess.shape))
    # numpy's own float32 transform (the reference's arithmetic) sits inside the same bound
    assert (np.abs(forc.fft_spectrum(data) - want) <= _tol(data, want)).all()
    plan = ip._shared_plan()
    assert plan.padded_shape() == tuple(forc._next_power_of_2(n) for n in shape)
    assert plan.last_ms() > 0 and plan.launch_count() >= 2


@pytest.mark.gpu
def test_spectrum_log_types_and_analytic_answers():
    n = 32
    z, y, x = np.meshgrid(*(np.arange(n),) * 3, indexing="ij")
    wave = (.5 + .5 * np.cos(2 * np.pi * (3 * x - 2 * y + 5 * z) / n)).astype(np.float32)
    s = ip.FFTProcessor().apply(wave)
    c = n // 2
    assert np.isclose(s[c, c, c], .5 * np.sqrt(n ** 3), rtol=1e-5)
    assert np.isclose(s[c + 5, c - 2, c + 3], .25 * np.sqrt(n ** 3), rtol=1e-5)
    assert np.isclose(s[c - 5, c + 2, c - 3], .25 * np.sqrt(n ** 3), rtol=1e-5)
    s[c, c, c] = s[c + 5, c - 2, c + 3] = s[c - 5, c + 2, c - 3] = 0
    assert s.max() < 1e-3
    lg = ip.FFTProcessor(log=True).apply(wave)
    want = forc.fft_spectrum(wave, log=True, precise=True)
    assert (np.abs(2. ** lg - 2. ** want) <= _tol(wave, 2. ** want) + 1e-6).all()
    assert ip.FFTProcessor(log=True).log is True and ip.FFTProcessor().name == "fft"
    # element types the device converts itself, and one it does not (float64 goes through float32 like astype)
    base = scenes.random_vol((12, 10, 20), np.uint8, seed=5)
    ref = forc.fft_spectrum(base, precise=True)
    for dt in (np.uint8, np.int16, np.uint16, np.float32, np.float64, np.int32):
        got = ip.FFTProcessor().apply(base.astype(dt))
        assert (np.abs(got - ref) <= _tol(base, ref)).all(), dt


@pytest.mark.gpu
def test_spectrum_in_a_chain_reaches_the_renderer_on_the_device():
    """blur -> spectrum -> renderer without leaving the device equals the reference-shaped host chain."""
    from spimagine_b200 import VolumeRenderer
    vol = scenes.vol_g(40, np.uint16, seed=2)
    chain = [ip.BlurProcessor(sigma=1.), ip.FFTProcessor(log=True)]
    M, P = scenes.gui_camera(0.3, 3.5)
    a, b = VolumeRenderer((96, 80)), VolumeRenderer((96, 80))
    try:
        for r in (a, b):
            r.set_data(vol.astype(np.float32))
            r.set_modelView(M)
            r.set_projection(P)
        ms = ip.apply_chain(a, vol, chain)
        assert ms > 0
        host = vol
        for p in chain:
            host = p.apply(host)
        b.update_data(host)
        a.render(maxVal=10., minVal=-10.)
        b.render(maxVal=10., minVal=-10.)
        assert np.array_equal(a.output, b.output) and a.output.max() > 0
        # a host-only processor behind the spectrum gets the spectrum as a host array
        seen = []
        ip.apply_chain(a, vol, [ip.FFTProcessor(), ip.FuncProcessor(lambda d: (seen.append(d.copy()), d)[1])])
        assert np.array_equal(seen[0], ip.FFTProcessor().apply(vol))
    finally:
        a.close()
        b.close()
