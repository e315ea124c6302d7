"""Minimal stand-in for mi.ScalarTransform4f (look_at / rotate / scale / translate / inverse / @), enough to
express the camera transforms of the reference (cameras.py:174-197, examples/optimize_volume.py:70-76)."""
from __future__ import annotations

import math

import numpy as np


class Transform4f:
    def __init__(self, m=None):
        if isinstance(m, Transform4f):
            m = m.matrix
        self.matrix = np.eye(4) if m is None else np.array(m, dtype=np.float64).reshape(4, 4)

    # Mitsuba's chained builders post-multiply: T().a().b() == A @ B
    def _chain(self, other):
        return Transform4f(self.matrix @ other)

    def translate(self, v):
        m = np.eye(4)
        m[:3, 3] = np.broadcast_to(np.asarray(v, np.float64), (3,))
        return self._chain(m)

    def scale(self, v):
        m = np.eye(4)
        m[0, 0], m[1, 1], m[2, 2] = np.broadcast_to(np.asarray(v, np.float64), (3,))
        return self._chain(m)

    def rotate(self, axis, angle):
        a = np.asarray(axis, np.float64)
        a = a / np.linalg.norm(a)
        t = math.radians(angle)
        c, s = math.cos(t), math.sin(t)
        x, y, z = a
        R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                      [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                      [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
        m = np.eye(4)
        m[:3, :3] = R
        return self._chain(m)

    def look_at(self, origin, target, up):
        origin = np.asarray(origin, np.float64)
        d = np.asarray(target, np.float64) - origin
        d = d / np.linalg.norm(d)
        left = np.cross(np.asarray(up, np.float64), d)
        left = left / np.linalg.norm(left)
        new_up = np.cross(d, left)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
        return self._chain(m)

    def inverse(self):
        return Transform4f(np.linalg.inv(self.matrix))

    def __matmul__(self, other):
        if isinstance(other, Transform4f):
            return Transform4f(self.matrix @ other.matrix)
        v = np.asarray(other, np.float64)
        if v.shape == (3,):  # point
            r = self.matrix @ np.append(v, 1.0)
            return r[:3] / r[3]
        return self.matrix @ v

    def transform_vector(self, v):
        return self.matrix[:3, :3] @ np.asarray(v, np.float64)

    def __repr__(self):
        return f"Transform4f({self.matrix.tolist()})"


ScalarTransform4f = Transform4f
T = Transform4f
