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


def _hash32(x):
    """lowbias32 integer hash on uint32 arrays (wrap-around arithmetic)."""
    x = x.copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def vol_g(N, dtype=np.uint16, seed=0, noise=0.01, shape=None, t=0):
    """SURVEY.md 8d Vol-G: 8 Gaussian blobs (centres U(-.6,.6)^3 drifting 0.01 t, sigma U(.08,.25), amplitude
    U(.3,1)) on linspace(-1,1) per axis, plus 1 % hashed per-voxel noise, scaled to peak 1.0 (f32), 60000 (u16)
    or 250 (u8).  The blobs are evaluated as separable products so that 512^3 takes seconds."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-.6, .6, (8, 3)) + 0.01 * t
    s = rng.uniform(.08, .25, 8)
    a = rng.uniform(.3, 1., 8)
    nz, ny, nx = shape if shape is not None else (N, N, N)
    z = np.linspace(-1, 1, nz, dtype=np.float32)
    y = np.linspace(-1, 1, ny, dtype=np.float32)
    x = np.linspace(-1, 1, nx, dtype=np.float32)
    v = np.zeros((nz, ny, nx), np.float32)
    for i in range(8):
        k = np.float32(1. / (2 * s[i] ** 2))
        gz = (np.float32(a[i]) * np.exp(-k * (z - np.float32(c[i, 0])) ** 2)).astype(np.float32)
        gy = np.exp(-k * (y - np.float32(c[i, 1])) ** 2).astype(np.float32)
        gx = np.exp(-k * (x - np.float32(c[i, 2])) ** 2).astype(np.float32)
        v += gz[:, None, None] * (gy[:, None] * gx[None, :])[None, :, :]
    if noise:
        for k0 in range(0, nz, 64):  # in slabs: bounds the temporaries for 512^3 and up
            k1 = min(k0 + 64, nz)
            idx = np.arange(k0 * ny * nx, k1 * ny * nx, dtype=np.uint64).astype(np.uint32)
            h = _hash32(idx ^ np.uint32((seed * 2654435761) & 0xffffffff))
            v[k0:k1] += (np.float32(noise) * (h >> np.uint32(8)).astype(np.float32) /
                         np.float32(1 << 24)).reshape(k1 - k0, ny, nx)
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
