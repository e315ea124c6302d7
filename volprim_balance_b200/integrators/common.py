"""Host-side mirror of volprim/integrators/common.py for the hot path.

What lives here is only what the `volprim_rf` / `volprim_tomography` plugins need on the host: the
10-float ellipsoid record, the kernel selector (names, flags, errors identical to the reference) and the
test-fixture builder.  All per-ray arithmetic (kernel eval, density integrals, ray/ellipsoid quadratic;
reference common.py:153-159, 193-243, 251-259, 287-333, 346-367) runs inside libvolprim_cuda.so.
`primitive_tracing`, `pdf`, `inv_cdf` and the Dr.Jit stack are `volprim_prb` machinery and out of scope.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass

import numpy as np
import torch

from .. import _cabi


class ADMode(enum.Enum):
    """dr.ADMode"""
    Primal = 0
    Forward = 1
    Backward = 2


class Properties(dict):
    """Stand-in for mi.Properties: a dict with `.get(name, default)` (reference volprim_rf.py:26-46)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)


@dataclass
class Ray3f:
    """Batch of rays: o, d [R,3] float32 CUDA tensors, maxt [R] or None (= infinity)."""
    o: torch.Tensor
    d: torch.Tensor
    maxt: torch.Tensor | None = None


@dataclass
class Ellipsoid:
    """Reference common.py:47-91.  `quat` is stored imaginary-first (i, j, k, r)."""
    center: torch.Tensor
    scale: torch.Tensor
    quat: torch.Tensor
    rot: torch.Tensor | None = None
    extent: float | None = 3.0

    @staticmethod
    def ravel(center, scale, quat) -> torch.Tensor:
        """common.py:55-65: AoS pack to the flat [N*10] `primitives.data` layout."""
        return torch.cat([center.reshape(-1, 3), scale.reshape(-1, 3), quat.reshape(-1, 4)], dim=1).reshape(-1)

    @staticmethod
    def unravel(data) -> "Ellipsoid":
        """common.py:67-74."""
        d = data.reshape(-1, 10)
        quat = d[:, 6:10]
        return Ellipsoid(d[:, 0:3], d[:, 3:6], quat, quat_to_matrix(quat), extent=None)


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """dr.quat_to_matrix(q, size=3) for q = (x, y, z, w); NOT normalised (reference quirk Q6)."""
    x, y, z, w = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


class Kernel:
    """common.py:94-147.  Only the selector and the flags are host-side; evaluation is in the CUDA kernels."""

    @staticmethod
    def factory(props):
        name = props.get('kernel_type', 'gaussian')
        if name == 'gaussian':
            return GaussianKernel(props)
        elif name == 'epanechnikov':
            return EpanechnikovKernel(props)
        else:
            raise Exception('Unknown kernel type! Should be one of "gaussian", "triangle" or "epanechnikov".')

    def __init__(self, props):
        self.type = props.get('kernel_type', 'gaussian')
        self.normalized = props.get('kernel_normalized', False)
        self.full_range = props.get('kernel_full_range', False)

    @property
    def cabi_id(self) -> int:
        raise NotImplementedError


class GaussianKernel(Kernel):
    @property
    def cabi_id(self) -> int:
        return _cabi.KERNEL_GAUSSIAN


class EpanechnikovKernel(Kernel):
    @property
    def cabi_id(self) -> int:
        return _cabi.KERNEL_EPANECHNIKOV


def euler_to_quat(euler_deg) -> np.ndarray:
    """dr.euler_to_quat(deg2rad(euler)) -> (x, y, z, w); ZYX convention as in Dr.Jit."""
    rx, ry, rz = (math.radians(float(a)) for a in euler_deg)
    cx, sx = math.cos(rx / 2), math.sin(rx / 2)
    cy, sy = math.cos(ry / 2), math.sin(ry / 2)
    cz, sz = math.cos(rz / 2), math.sin(rz / 2)
    w = cx * cy * cz + sx * sy * sz
    x = sx * cy * cz - cx * sy * sz
    y = cx * sy * cz + sx * cy * sz
    z = cx * cy * sz - sx * sy * cz
    return np.array([x, y, z, w], np.float32)


class EllipsoidsFactory:
    """Helper class to build ellipsoid datasets for testing purposes (reference common.py:566-596)."""

    def __init__(self):
        self.centers, self.scales, self.quaternions, self.sigmats, self.albedos = [], [], [], [], []

    def add(self, mean, scale, sigmat=1.0, albedo=1.0, euler=(0.0, 0.0, 0.0)):
        self.centers.append(np.broadcast_to(np.asarray(mean, np.float32), (3,)).copy())
        self.scales.append(np.broadcast_to(np.asarray(scale, np.float32), (3,)).copy())
        self.quaternions.append(euler_to_quat(euler))
        self.sigmats.append(float(sigmat))
        if isinstance(albedo, (float, int)):
            albedo = [float(albedo)] * 3
        self.albedos.append(np.asarray(albedo, np.float32))

    def build(self):
        n = len(self.centers)
        centers = torch.from_numpy(np.array(self.centers, np.float32).reshape(n, 3))
        scales = torch.from_numpy(np.array(self.scales, np.float32).reshape(n, 3))
        quats = torch.from_numpy(np.array(self.quaternions, np.float32).reshape(n, 4))
        sigmats = torch.tensor(self.sigmats, dtype=torch.float32).reshape(n, 1)
        albedos = torch.from_numpy(np.array(self.albedos, np.float32).reshape(n, -1))
        return centers, scales, quats, sigmats, albedos


def get_ellipsoids_shape(scene, requested=True):
    """common.py:22-33."""
    for shape in scene.shapes():
        if getattr(shape, "is_ellipsoids", False):
            return shape
    if requested:
        raise Exception("Couldn't find ellipsoids shape in the scene!")
    return None
