"""OpenGL-convention 4x4 camera matrices used to drive the renderer.

Host-side mirror of the helpers a spimagine caller builds `modelView` / `projection` from
(reference: spimagine/utils/transform_matrices.py:21-118 and the quaternion -> rotation formula of
spimagine/utils/quaternion.py:56-63).  Same names, argument meaning and dtypes; pure numpy.
"""
import numpy as np

__all__ = ["mat4_scale", "mat4_rotation", "mat4_rotation_euler", "mat4_perspective", "mat4_stereo_perspective",
           "mat4_frustrum", "mat4_ortho", "mat4_identity", "mat4_translate", "mat4_lookat"]


def mat4_identity():
    return np.identity(4)


def mat4_scale(x=1., y=1., z=1.):
    return np.diag(np.array([x, y, z, 1.], np.float32))


def mat4_translate(x=0, y=0, z=0):
    m = np.identity(4)
    m[:3, 3] = x, y, z
    return m


def mat4_rotation(w=0, x=1, y=0, z=0):
    """Rotation by angle `w` (radians) about the axis (x, y, z), via the unit quaternion
    (cos w/2, sin w/2 * axis)."""
    axis = np.array([x, y, z], np.float32)
    axis *= 1. / np.sqrt(1. * np.sum(axis ** 2))
    a = np.cos(.5 * w)
    b, c, d = np.sin(.5 * w) * axis
    return np.array([
        [a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c), 0],
        [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b), 0],
        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d, 0],
        [0, 0, 0, 1]])


def mat4_rotation_euler(yaw=0, pitch=0, roll=0):
    """z-y'-x' convention."""
    return np.dot(mat4_rotation(yaw, 0, 0, 1), np.dot(mat4_rotation(pitch, 0, 1, 0), mat4_rotation(roll, 1, 0, 0)))


def mat4_perspective(fovy=45, aspect=1., z1=0.1, z2=10):
    """gluPerspective(fovy [degrees], aspect, zNear, zFar)."""
    f = 1. / np.tan(fovy / 180. * np.pi / 2.)
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 1. * f / aspect
    m[1, 1] = f
    m[2, 2] = -1. * (z2 + z1) / (z2 - z1)
    m[2, 3] = -2. * z1 * z2 / (z2 - z1)
    m[3, 2] = -1
    return m


def mat4_frustrum(left, right, bottom, top, zNear, zFar):
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 2. * zNear / (right - left)
    m[0, 2] = 1. * (right + left) / (right - left)
    m[1, 1] = 2. * zNear / (top - bottom)
    m[1, 2] = (top + bottom) / (top - bottom)
    m[2, 2] = -1. * (zFar + zNear) / (zFar - zNear)
    m[2, 3] = -2. * zFar * zNear / (zFar - zNear)
    m[3, 2] = -1.
    return m


def mat4_stereo_perspective(fovy=45, aspect=1., z1=0.1, z2=10, eye_shift=0):
    h = z1 * np.tan(fovy / 180. * np.pi / 2.)
    w = h * aspect
    return mat4_frustrum(-w - eye_shift, w - eye_shift, -h, h, z1, z2)


def mat4_ortho(x1=-1, x2=1, y1=-1, y2=1, z1=-1, z2=1):
    """glOrtho."""
    m = np.zeros((4, 4), np.float32)
    m[0, 0], m[0, 3] = 2. / (x2 - x1), -1. * (x2 + x1) / (x2 - x1)
    m[1, 1], m[1, 3] = 2. / (y2 - y1), -1. * (y2 + y1) / (y2 - y1)
    m[2, 2], m[2, 3] = -2. / (z2 - z1), -1. * (z2 + z1) / (z2 - z1)
    m[3, 3] = 1.
    return m


def mat4_lookat(eye, center, up):
    """gluLookAt."""
    eye = np.array(eye, np.float32)
    fwd = np.array(center, np.float32) - eye
    fwd *= 1. / np.sqrt(np.sum(fwd ** 2))
    side = np.cross(fwd, np.array(up, np.float32))
    side *= 1. / np.sqrt(np.sum(side ** 2))
    up = np.cross(side, fwd)
    m = np.identity(4)
    m[0, :3], m[1, :3], m[2, :3] = side, up, -fwd
    return np.dot(m, mat4_translate(*(-eye)))
