"""Input recipes shared by the parity tests.  The volumes and cameras are the ones the reference's own
(assert-free) tests use, plus the synthetic Vol-G volume of SURVEY.md 8d."""
import numpy as np

from spimagine_b200.utils.transform_matrices import (mat4_identity, mat4_perspective, mat4_rotation, mat4_scale,
                                                     mat4_translate)


def grid(N):
    x = np.linspace(-1, 1, N)
    return np.meshgrid(x, x, x, indexing="ij")  # (z, y, x)


def two_blobs(N=128):
    """tests/test_rendering/test_simple_rendering.py:17-36"""
    Z, Y, X = grid(N)
    R1 = np.sqrt((X - .2) ** 2 + (Y + .2) ** 2 + Z ** 2)
    R2 = np.sqrt((X + .2) ** 2 + (Y + .2) ** 2 + Z ** 2)
    return (255 * (np.exp(-30 * R1 ** 2) + np.exp(-30 * R2 ** 2))).astype(np.float32)


def gaussian(N=128, amp=200., k=10.):
    """tests/test_volumerender/test_volumerender.py:65-93, 173-188: amp * exp(-k R^2)"""
    Z, Y, X = grid(N)
    return (amp * np.exp(-k * (X ** 2 + Y ** 2 + Z ** 2))).astype(np.float32)


def iso_sphere(N=64):
    """tests/test_rendering/test_simple_rendering.py:55-68: uint16 900*exp(-10 R), iso at maxVal/2 = 10"""
    Z, Y, X = grid(N)
    return (900. * np.exp(-10. * np.sqrt(X ** 2 + Y ** 2 + Z ** 2))).astype(np.uint16)


def linspace_vol(N=64):
    """tests/test_volumerender/test_volumerender.py:31-40"""
    return np.linspace(0, 1, N ** 3).reshape((N,) * 3).astype(np.float32)


def random_vol(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype) == np.float32:
        return rng.random(shape, dtype=np.float32)
    hi = 65535 if np.dtype(dtype) == np.uint16 else 255
    return rng.integers(0, hi + 1, size=shape).astype(dtype)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def vol_g(N, dtype=np.uint16, seed=0, noise=0.01, shape=None, t=0):
    """SURVEY.md 8d Vol-G: 8 Gaussian blobs + 1 % hashed noise, peak 1.0 (f32) / 60000 (u16) / 250 (u8)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-.6, .6, (8, 3)) + 0.01 * t
    s = rng.uniform(.08, .25, 8)
    a = rng.uniform(.3, 1., 8)
    nz, ny, nx = shape if shape is not None else (N, N, N)
    z = np.linspace(-1, 1, nz, dtype=np.float32)[:, None, None]
    y = np.linspace(-1, 1, ny, dtype=np.float32)[None, :, None]
    x = np.linspace(-1, 1, nx, dtype=np.float32)[None, None, :]
    v = np.zeros((nz, ny, nx), np.float32)
    for i in range(8):
        v += np.float32(a[i]) * np.exp(-((x - np.float32(c[i, 2])) ** 2 + (y - np.float32(c[i, 1])) ** 2 +
                                          (z - np.float32(c[i, 0])) ** 2) / np.float32(2 * s[i] ** 2))
    with np.errstate(over="ignore"):
        idx = np.arange(v.size, dtype=np.uint64) ^ np.uint64(seed)
        h = _splitmix64(idx)
    v += np.float32(noise) * ((h >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)).reshape(v.shape)
    v /= v.max()
    if np.dtype(dtype) == np.float32:
        return v
    peak = 60000. if np.dtype(dtype) == np.uint16 else 250.
    return np.rint(v * peak).astype(dtype)


def gui_camera(theta=0.0, dist=4.0, fovy=60.):
    """P = perspective(60,1,.1,10), M = translate(0,0,-4) . rotation(theta, 0,1,0)"""
    P = mat4_perspective(fovy, 1., .1, 10)
    M = np.dot(mat4_translate(0, 0, -dist), mat4_rotation(theta + 1e-3, 0, 1, 0))
    return M, P


def tilted_camera(dist=3.2):
    M = np.dot(mat4_translate(0.05, -0.03, -dist),
               np.dot(mat4_rotation(0.6, 0.3, 1, 0.2), mat4_scale(.9, .9, .9)))
    return M, mat4_perspective(45, 1., .1, 10)
