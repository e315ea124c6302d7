"""Camera specifications for the ray-generation step in front of the hot path.

Mirrors the interface of the reference's volprim/cameras.py that the integrator callers use:
`CameraSpecs` (same constructor arguments and attributes) -> `to_dict()` = the `perspective` sensor dictionary with
the reference's keys (cameras.py:114-137), and `JSONCameraSpecsIO` for 3DGS-style `cameras.json` files
(cameras.py:169-217).  KRT / COLMAP ingestion (cameras.py:221-375, colmap_loader.py) is dataset tooling outside the
hot path and is not provided.  Transforms are numpy 4x4 matrices wrapped in `transforms.Transform4f`.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .transforms import Transform4f

_FLIP_XY = (-1.0, -1.0, 1.0)      # handedness flip between the 3DGS / GSplat camera frame and Mitsuba's


def fov2focal(fov: float, width: int) -> float:
    """Focal length in pixels of a sensor `width` pixels wide with a horizontal field of view `fov` (degrees)."""
    return 0.5 * width / math.tan(0.5 * math.radians(fov))


def focal2fov(focal_length: float, width: int) -> float:
    """Horizontal field of view (degrees) of a sensor `width` pixels wide with the given focal length (pixels)."""
    return math.degrees(2.0 * math.atan2(0.5 * width, focal_length))


@dataclass
class CameraSpecs:
    """Pin-hole camera description.  Exactly one of `fov` (degrees, along x) and `focal_length` (pixels) is given;
    the other is derived.  `cx`, `cy` are principal-point offsets in [-1, 1]; the distortion coefficients are carried
    for completeness only (the perspective sensor ignores them, as in the reference)."""
    name: str
    width: int
    height: int
    to_world: Transform4f
    fov: Optional[float] = None
    focal_length: Optional[float] = None
    near_clip: float = 0.1
    far_clip: float = 10000.0
    cx: float = 0.0
    cy: float = 0.0
    k1: float = 0.0
    k2: float = 0.0
    k3: float = 0.0
    k4: float = 0.0
    k5: float = 0.0
    k6: float = 0.0
    p1: float = 0.0
    p2: float = 0.0

    def __post_init__(self):
        self.to_world = Transform4f(self.to_world)
        if (self.fov is None) == (self.focal_length is None):
            raise Exception('CameraSpecs: either FOV or focal length should be set!')
        if self.fov is None:
            self.fov = focal2fov(self.focal_length, self.width)
        else:
            self.focal_length = fov2focal(self.fov, self.width)

    # -- derived matrices ---------------------------------------------------------------------------------------
    def viewmat(self) -> np.ndarray:
        """World-to-camera matrix in the GSplat convention."""
        return np.array(self.to_world.scale(_FLIP_XY).inverse().matrix)

    def K(self) -> np.ndarray:
        """3x3 intrinsics with the principal point at the image centre."""
        f = self.focal_length
        return np.array([[f, 0.0, 0.5 * self.width], [0.0, f, 0.5 * self.height], [0.0, 0.0, 1.0]])

    # -- sensor dictionary --------------------------------------------------------------------------------------
    def to_dict(self, resolution_factor: float = 1.0, pixel_format: str = 'rgb', pixel_filter: str = 'tent') -> dict:
        """The `perspective` sensor dictionary (keys as the reference emits them for mi.load_dict)."""
        film = dict(type='hdrfilm', rfilter=dict(type=pixel_filter), pixel_format=pixel_format,
                    width=int(self.width * resolution_factor), height=int(self.height * resolution_factor))
        return dict(type='perspective', principal_point_offset_x=self.cx, principal_point_offset_y=self.cy,
                    fov_axis='x', fov=self.fov, to_world=self.to_world, near_clip=self.near_clip,
                    far_clip=self.far_clip, film=film)

    @staticmethod
    def from_dict(d: dict, name: str = '') -> "CameraSpecs":
        film = d['film']
        return CameraSpecs(name, film['width'], film['height'], d['to_world'], fov=d['fov'],
                           near_clip=d.get('near_clip', 0.1), far_clip=d.get('far_clip', 10000.0),
                           cx=d.get('principal_point_offset_x', 0.0), cy=d.get('principal_point_offset_y', 0.0))

    def __repr__(self):
        body = '\n'.join(f'  {k}: {v}' for k, v in vars(self).items())
        return f'CameraSpecs[\n{body}]'


class CameraSpecsIO:
    """Reader / writer interface."""

    @staticmethod
    def load(filename: str) -> List[CameraSpecs]:
        raise Exception('Loader not implemented')

    @staticmethod
    def write(specs: List[CameraSpecs], filename: str):
        raise Exception('Loader not implemented')


class JSONCameraSpecsIO(CameraSpecsIO):
    """3DG-dataset `cameras.json`: a list of {id, img_name, width, height, position, rotation, fx, fy}."""

    @staticmethod
    def load(filename: str) -> List[CameraSpecs]:
        with open(filename) as fh:
            entries = json.load(fh)
        out = []
        for e in entries:
            pose = np.eye(4)
            pose[:3, :3] = np.asarray(e['rotation'], dtype=np.float64)     # stored as the camera-to-world rotation
            pose[:3, 3] = np.asarray(e['position'], dtype=np.float64)
            # near / far as the reference hard-codes them for these datasets (cameras.py:193-194)
            out.append(CameraSpecs(e['img_name'], e['width'], e['height'], Transform4f(pose).scale(_FLIP_XY),
                                   focal_length=e['fx'], near_clip=0.01 * 10, far_clip=100.0))
        return out

    @staticmethod
    def write(specs: List[CameraSpecs], filename: str):
        entries = []
        for idx, cam in enumerate(specs):
            pose = np.array((cam.to_world @ Transform4f().scale(_FLIP_XY)).matrix)
            entries.append(dict(rotation=pose[:3, :3].tolist(), position=pose[:3, 3].tolist(), fx=cam.focal_length,
                                fy=cam.focal_length, width=cam.width, height=cam.height, id=idx, img_name=cam.name))
        with open(filename, 'w', encoding='utf-8') as fh:
            fh.write(json.dumps(entries, ensure_ascii=False))
