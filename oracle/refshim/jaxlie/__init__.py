"""Stand-in for `jaxlie` (test infrastructure only, see ../README.md): SO3 / SE3 restated
from jaxlie's published definitions (quaternion wxyz; SE3 parameters = wxyz_xyz; adjoint =
[[R, S(t) R], [0, R]]), with leading batch axes like jaxlie >= 1.4."""
import numpy as np

from jax._core import asarray as _arr


def _f(x):
    return np.asarray(x, dtype=float)


def _skew(v):
    v = _f(v)
    z = np.zeros_like(v[..., 0])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


class SO3:
    def __init__(self, wxyz):
        self.wxyz = _arr(wxyz)

    def get_batch_axes(self):
        return self.wxyz.shape[:-1]

    # ---- factories
    @staticmethod
    def identity(batch_axes=()):
        return SO3(np.broadcast_to(np.array([1.0, 0.0, 0.0, 0.0]), (*batch_axes, 4)).copy())

    @staticmethod
    def from_x_radians(theta):
        t = _f(theta)
        return SO3.exp(np.stack([t, np.zeros_like(t), np.zeros_like(t)], -1))

    @staticmethod
    def from_y_radians(theta):
        t = _f(theta)
        return SO3.exp(np.stack([np.zeros_like(t), t, np.zeros_like(t)], -1))

    @staticmethod
    def from_z_radians(theta):
        t = _f(theta)
        return SO3.exp(np.stack([np.zeros_like(t), np.zeros_like(t), t], -1))

    @staticmethod
    def from_rpy_radians(roll, pitch, yaw):
        return SO3.from_z_radians(yaw) @ SO3.from_y_radians(pitch) @ SO3.from_x_radians(roll)

    @staticmethod
    def from_quaternion_xyzw(xyzw):
        x = _f(xyzw)
        return SO3(np.concatenate([x[..., 3:4], x[..., 0:3]], -1))

    @staticmethod
    def from_matrix(matrix):
        m = _f(matrix)
        assert m.shape[-2:] == (3, 3)
        m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
        m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
        m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
        t0 = 1 + m00 - m11 - m22
        q0 = np.stack([m21 - m12, t0, m10 + m01, m02 + m20], -1)
        t1 = 1 - m00 + m11 - m22
        q1 = np.stack([m02 - m20, m10 + m01, t1, m21 + m12], -1)
        t2 = 1 - m00 - m11 + m22
        q2 = np.stack([m10 - m01, m02 + m20, m21 + m12, t2], -1)
        t3 = 1 + m00 + m11 + m22
        q3 = np.stack([t3, m21 - m12, m02 - m20, m10 - m01], -1)
        c0, c1, c2 = m22 < 0, m00 > m11, m00 < -m11
        t = np.where(c0, np.where(c1, t0, t1), np.where(c2, t2, t3))
        q = np.where(c0[..., None], np.where(c1[..., None], q0, q1), np.where(c2[..., None], q2, q3))
        return SO3(q * 0.5 / np.sqrt(t)[..., None])

    @staticmethod
    def exp(tangent):
        t = _f(tangent)
        theta_sq = np.sum(t * t, -1)
        small = theta_sq < 1e-16  # jaxlie switches to a Taylor expansion near zero
        safe = np.where(small, 1.0, theta_sq)
        theta = np.sqrt(safe)
        real = np.where(small, 1.0 - theta_sq / 8.0, np.cos(0.5 * theta))
        imag = np.where(small, 0.5 - theta_sq / 48.0, np.sin(0.5 * theta) / theta)
        return SO3(np.concatenate([real[..., None], imag[..., None] * t], -1))

    # ---- accessors
    def as_matrix(self):
        q = _f(self.wxyz)
        q = q * np.sqrt(2.0 / np.sum(q * q, -1, keepdims=True))
        o = q[..., :, None] * q[..., None, :]
        R = np.stack([
            np.stack([1.0 - o[..., 2, 2] - o[..., 3, 3], o[..., 1, 2] - o[..., 3, 0], o[..., 1, 3] + o[..., 2, 0]], -1),
            np.stack([o[..., 1, 2] + o[..., 3, 0], 1.0 - o[..., 1, 1] - o[..., 3, 3], o[..., 2, 3] - o[..., 1, 0]], -1),
            np.stack([o[..., 1, 3] - o[..., 2, 0], o[..., 2, 3] + o[..., 1, 0], 1.0 - o[..., 1, 1] - o[..., 2, 2]], -1)], -2)
        return _arr(R)

    def parameters(self):
        return self.wxyz

    def log(self):
        q = _f(self.wxyz)
        w = q[..., 0]
        n_sq = np.sum(q[..., 1:] ** 2, -1)
        small = n_sq < 1e-16
        n = np.sqrt(np.where(small, 1.0, n_sq))
        w_safe = np.where(w == 0, 1.0, w)
        half = np.arctan2(np.where(w < 0, -n, n), np.abs(w))
        f = np.where(small, 2.0 / w_safe - 2.0 / 3.0 * n_sq / w_safe**3, 2.0 * half / n)
        return _arr(f[..., None] * q[..., 1:])

    def inverse(self):
        return SO3(_f(self.wxyz) * np.array([1.0, -1.0, -1.0, -1.0]))

    def normalize(self):
        q = _f(self.wxyz)
        return SO3(q / np.linalg.norm(q, axis=-1, keepdims=True))

    def adjoint(self):
        return self.as_matrix()

    def apply(self, target):
        return _arr(np.einsum("...ij,...j->...i", np.asarray(self.as_matrix()), _f(target)))

    def multiply(self, other):
        a, b = _f(self.wxyz), _f(other.wxyz)
        w0, x0, y0, z0 = (a[..., k] for k in range(4))
        w1, x1, y1, z1 = (b[..., k] for k in range(4))
        return SO3(np.stack([
            -x0 * x1 - y0 * y1 - z0 * z1 + w0 * w1,
            x0 * w1 + y0 * z1 - z0 * y1 + w0 * x1,
            -x0 * z1 + y0 * w1 + z0 * x1 + w0 * y1,
            x0 * y1 - y0 * x1 + z0 * w1 + w0 * z1], -1))

    def __matmul__(self, other):
        if isinstance(other, SO3):
            return self.multiply(other)
        return self.apply(other)


class SE3:
    def __init__(self, wxyz_xyz):
        self.wxyz_xyz = _arr(wxyz_xyz)

    def get_batch_axes(self):
        return self.wxyz_xyz.shape[:-1]

    @staticmethod
    def identity(batch_axes=()):
        return SE3(np.broadcast_to(np.array([1.0, 0, 0, 0, 0, 0, 0]), (*batch_axes, 7)).copy())

    @staticmethod
    def from_rotation_and_translation(rotation, translation):
        q, t = _f(rotation.wxyz), _f(translation)
        assert t.shape[-1] == 3
        sh = np.broadcast_shapes(q.shape[:-1], t.shape[:-1])
        return SE3(np.concatenate([np.broadcast_to(q, (*sh, 4)), np.broadcast_to(t, (*sh, 3))], -1))

    @staticmethod
    def from_rotation(rotation):
        return SE3.from_rotation_and_translation(rotation, np.zeros(3))

    @staticmethod
    def from_translation(translation):
        return SE3.from_rotation_and_translation(SO3.identity(), translation)

    @staticmethod
    def from_matrix(matrix):
        m = _f(matrix)
        assert m.shape[-2:] in ((4, 4), (3, 4))
        return SE3.from_rotation_and_translation(SO3.from_matrix(m[..., :3, :3]), m[..., :3, 3])

    def rotation(self):
        return SO3(np.asarray(self.wxyz_xyz)[..., :4])

    def translation(self):
        return _arr(np.asarray(self.wxyz_xyz)[..., 4:])

    def parameters(self):
        return self.wxyz_xyz

    def as_matrix(self):
        R = np.asarray(self.rotation().as_matrix())
        t = np.asarray(self.translation())
        H = np.zeros((*R.shape[:-2], 4, 4))
        H[..., :3, :3] = R
        H[..., :3, 3] = t
        H[..., 3, 3] = 1.0
        return _arr(H)

    def inverse(self):
        Rinv = self.rotation().inverse()
        return SE3.from_rotation_and_translation(Rinv, -np.asarray(Rinv.apply(self.translation())))

    def adjoint(self):
        R = np.asarray(self.rotation().as_matrix())
        X = np.zeros((*R.shape[:-2], 6, 6))
        X[..., :3, :3] = R
        X[..., :3, 3:] = _skew(self.translation()) @ R
        X[..., 3:, 3:] = R
        return _arr(X)

    def apply(self, target):
        return _arr(np.asarray(self.rotation().apply(target)) + np.asarray(self.translation()))

    def multiply(self, other):
        return SE3.from_rotation_and_translation(
            self.rotation() @ other.rotation(),
            np.asarray(self.rotation().apply(other.translation())) + np.asarray(self.translation()))

    def __matmul__(self, other):
        if isinstance(other, SE3):
            return self.multiply(other)
        return self.apply(other)
