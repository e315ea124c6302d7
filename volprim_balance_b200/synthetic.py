"""Synthetic 3DGS-style primitive clouds and camera rings.

All arrays are numpy float32 in the reference layouts (`primitives.data` flat N*10 with the quaternion
stored imaginary-first (i, j, k, r) -- /root/reference/volprim/integrators/common.py:55-74).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Cloud:
    data: np.ndarray        # [N, 10] center3, scale3, quat(i,j,k,r)
    opacities: np.ndarray   # [N]
    sh_coeffs: np.ndarray   # [N, 3 (D+1)^2] coefficient-major, channel-minor
    extent: float = 3.0

    @property
    def n(self) -> int:
        return self.data.shape[0]


def make_cloud(n: int, sigma0: float, seed: int, sh_degree: int = 3, centers: str = "uniform",
               mu_opacity: float = -1.0, extent: float = 3.0) -> Cloud:
    """`centers`: "uniform" U[-1,1]^3 (cfg 2/4), "truck" N(0,0.6^2) clipped to [-2,2] (cfg 3),
    "dense" N(0,0.4^2) (cfg 5)."""
    rng = np.random.default_rng(seed)
    if centers == "uniform":
        c = rng.uniform(-1.0, 1.0, size=(n, 3))
    elif centers == "truck":
        c = np.clip(rng.normal(0.0, 0.6, size=(n, 3)), -2.0, 2.0)
    elif centers == "dense":
        c = rng.normal(0.0, 0.4, size=(n, 3))
    else:
        raise ValueError(f"unknown centre distribution {centers!r}")
    s = np.exp(rng.normal(math.log(sigma0), 0.5, size=(n, 3)))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    op = 1.0 / (1.0 + np.exp(-rng.normal(mu_opacity, 1.5, size=n)))
    nb = (sh_degree + 1) ** 2
    sh = np.empty((n, nb, 3))
    sh[:, 0, :] = rng.normal(0.0, 0.5, size=(n, 3))
    if nb > 1:
        sh[:, 1:, :] = rng.normal(0.0, 0.1, size=(n, nb - 1, 3))
    data = np.concatenate([c, s, q], axis=1).astype(np.float32)
    return Cloud(data, op.astype(np.float32), sh.reshape(n, nb * 3).astype(np.float32), extent)


def sigma0_for_hits(n: int, target_crossings: float, half_extent: float = 1.0, extent: float = 3.0) -> float:
    """First guess for sigma0 from the mean number of bounding ellipsoids a ray through a uniform cube
    crosses: H = n * pi (extent sigma)^2 e^{0.5} * L / V with L = 2 h, V = (2h)^3 (log-normal axes)."""
    area_per = n * math.pi * extent * extent * math.exp(0.5)
    return math.sqrt(target_crossings * (2 * half_extent) ** 2 / area_per)


@dataclass
class Camera:
    """Pin-hole camera in Mitsuba's `perspective` convention (cameras.py:114-137): fov along x."""
    to_world: np.ndarray  # [4,4], columns = (left, up, forward, origin) in Mitsuba's look_at convention
    fov_x_deg: float
    width: int
    height: int
    near_clip: float = 0.01
    far_clip: float = 10000.0
    cx: float = 0.0
    cy: float = 0.0


def look_at(origin, target, up) -> np.ndarray:
    """mi.ScalarTransform4f.look_at: dir = normalize(target-origin); left = normalize(cross(up, dir));
    new_up = cross(dir, left); columns (left, new_up, dir, origin)."""
    origin = np.asarray(origin, np.float64)
    d = np.asarray(target, np.float64) - origin
    d /= np.linalg.norm(d)
    left = np.cross(np.asarray(up, np.float64), d)
    left /= np.linalg.norm(left)
    nup = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, nup, d, origin
    return m


def ring_camera(i: int, n_views: int, width: int, height: int, radius: float = 4.0, fov_x_deg: float = 40.0,
                elevation_deg: float = 15.0) -> Camera:
    """Ring of radius 4 around the origin, elevations alternating +-15 degrees, looking at the origin."""
    az = 2.0 * math.pi * i / max(n_views, 1)
    el = math.radians(elevation_deg if i % 2 == 0 else -elevation_deg)
    o = radius * np.array([math.cos(el) * math.sin(az), math.sin(el), math.cos(el) * math.cos(az)])
    return Camera(look_at(o, [0, 0, 0], [0, 1, 0]), fov_x_deg, width, height)


def camera_rays(cam: Camera, pixel_offsets=((0.5, 0.5),), tile=None):
    """Pixel rays in Mitsuba's perspective-sensor convention (rays start ON the near plane, maxt = far-near
    along the ray): returns o [H*W*S,3], d [H*W*S,3], maxt [H*W*S] float32, pixel-major then sample."""
    W, H = cam.width, cam.height
    offs = np.asarray(pixel_offsets, np.float64).reshape(-1, 2)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    px = (xs[..., None] + offs[:, 0]).reshape(-1)
    py = (ys[..., None] + offs[:, 1]).reshape(-1)
    return rays_from_samples(cam, px / W, py / H)


def rays_from_samples(cam: Camera, u, v):
    """u, v in [0,1]^2 film coordinates (u to the right in the image, v down)."""
    W, H = cam.width, cam.height
    aspect = W / H
    tan_half = math.tan(math.radians(cam.fov_x_deg) * 0.5)
    # Mitsuba: camera looks along +z, +x is LEFT in the image, +y up.
    x = (1.0 - 2.0 * np.asarray(u, np.float64) + cam.cx * 2.0) * tan_half
    y = (1.0 - 2.0 * np.asarray(v, np.float64) + cam.cy * 2.0) * tan_half / aspect
    dloc = np.stack([x, y, np.ones_like(x)], axis=-1)
    dloc /= np.linalg.norm(dloc, axis=-1, keepdims=True)
    inv_z = 1.0 / dloc[:, 2]
    Rm = cam.to_world[:3, :3]
    d = dloc @ Rm.T
    near_t = cam.near_clip * inv_z
    far_t = cam.far_clip * inv_z
    o = cam.to_world[:3, 3][None, :] + d * near_t[:, None]
    return o.astype(np.float32), d.astype(np.float32), (far_t - near_t).astype(np.float32)
