"""Unit quaternions for camera rotations and their interpolation between keyframes.

Host-side mirror of spimagine/utils/quaternion.py:5-83 (class Quaternion, quaternion_slerp): same names, the
(w, x, y, z) component order in `.data`, the same operator set and the same slerp rule (normalise both ends, plain
lerp when |<q1,q2>| > 0.9998, shorter great circle otherwise), because keyframe files store `.data` and the
renderer's modelView is built from `toRotation4()`.  Pure numpy float64.

One deliberate difference: `conj()` negates x, y, z.  The reference's (quaternion.py:23) returns -y in the z slot;
nothing on the render path calls it.
"""
import numpy as np

__all__ = ["Quaternion", "quaternion_slerp"]


class Quaternion(object):
    def __init__(self, w=1., x=0., y=0., z=0.):
        self.data = np.array([w, x, y, z])

    @classmethod
    def _of(cls, data):
        """The quaternion whose components ARE the 4-vector `data` (a fresh array): what Quaternion(*data) builds,
        without unpacking the vector and packing it again (the record loop interpolates one per frame)."""
        q = cls.__new__(cls)
        q.data = data
        return q

    @classmethod
    def copy(cls, rhs):
        return cls(*rhs.data)

    def __getitem__(self, i):
        return self.data[i]

    def __setitem__(self, i, val):
        self.data[i] = val

    def __repr__(self):
        return "Quaternion(%s,%s,%s,%s)" % tuple(self.data)

    def __add__(self, q):
        return Quaternion._of(self.data + q.data)

    def __sub__(self, q):
        return Quaternion._of(self.data - q.data)

    def __mul__(self, q):
        """Hamilton product with a Quaternion, component scaling with a number."""
        if not isinstance(q, Quaternion):
            return Quaternion._of(np.asarray(q * self.data))
        w1, v1 = self.data[0], self.data[1:]
        w2, v2 = q.data[0], q.data[1:]
        # written out per component so that the rounding matches the reference's expression order
        # (quaternion.py:40-43): w1*w2 - x1*x2 - y1*y2 - z1*z2, then w1*v2 + v1*w2 + (v1 x v2)
        return Quaternion(w1 * w2 - v1[0] * v2[0] - v1[1] * v2[1] - v1[2] * v2[2],
                          w1 * v2[0] + v1[0] * w2 + v1[1] * v2[2] - v1[2] * v2[1],
                          w1 * v2[1] - v1[0] * v2[2] + v1[1] * w2 + v1[2] * v2[0],
                          w1 * v2[2] + v1[0] * v2[1] - v1[1] * v2[0] + v1[2] * w2)

    def conj(self):
        return Quaternion(self.data[0], -self.data[1], -self.data[2], -self.data[3])

    def norm(self):
        # np.linalg.norm of a real vector: sqrt(x . x) after a conversion of integer components to float
        d = self.data if self.data.dtype.kind == "f" else self.data.astype(float)
        return np.sqrt(d.dot(d))

    def dot(self, q):
        return np.inner(self.data, q.data)

    def normalize(self):
        return Quaternion._of(self.data * 1. / self.norm())

    def toRotation3(self):
        a, b, c, d = self.data
        return np.array([
            [a ** 2 + b ** 2 - c ** 2 - d ** 2, 2 * (b * c - a * d), 2 * (b * d + a * c)],
            [2 * (b * c + a * d), a ** 2 - b ** 2 + c ** 2 - d ** 2, 2 * (c * d - a * b)],
            [2 * (b * d - a * c), 2 * (c * d + a * b), a ** 2 - b ** 2 - c ** 2 + d ** 2]])

    def toRotation4(self):
        m = np.zeros((4, 4))
        m[:3, :3] = self.toRotation3()
        m[3, 3] = 1
        return m


def quaternion_slerp(q1, q2, t):
    """Spherical interpolation from q1 (t = 0) to q2 (t = 1); quaternion.py:73-83."""
    q1, q2 = q1.normalize(), q2.normalize()
    prod = q1.dot(q2)
    if abs(prod) > .9998:  # nearly parallel: the arc is a chord (not renormalised, as in the reference)
        return q1 + (q2 - q1) * t
    if prod < 0:  # the other representative of q2 is closer
        q2 = q2 * (-1.)
        prod *= -1.
    w = np.arccos(prod)
    return (q1 * (np.sin((1. - t) * w) / np.sin(w))) + q2 * (np.sin(t * w) / np.sin(w))
