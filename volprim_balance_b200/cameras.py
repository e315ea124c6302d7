"""Camera specifications in the formats of volprim/cameras.py: `CameraSpecs` -> perspective-sensor dictionary
(cameras.py:114-137) and the 3DGS `cameras.json` reader / writer (cameras.py:169-217).  KRT and COLMAP
ingestion (cameras.py:221-375, colmap_loader.py) are dataset tooling outside the hot path and not provided."""
from __future__ import annotations

import json
import math
from typing import List

import numpy as np

from .transforms import Transform4f


def fov2focal(fov: float, width: int):
    '''Focal length (pixels) for a given sensor resolution and FOV (degrees)'''
    return (width / 2.0) / math.tan(math.radians(fov) * 0.5)


def focal2fov(focal_length: float, width: int):
    '''FOV (degrees) for a given sensor resolution and focal length'''
    return 2.0 * math.degrees(math.atan2(0.5 * width, focal_length))


class CameraSpecs:
    '''
    Camera information data structure (reference cameras.py:53-165).
    '''
    def __init__(self, name: str, width: int, height: int, to_world, fov: float = None, focal_length: float = None,
                 near_clip: float = 0.1, far_clip: float = 10000.0, cx: float = 0.0, cy: float = 0.0,
                 k1=0.0, k2=0.0, k3=0.0, k4=0.0, k5=0.0, k6=0.0, p1=0.0, p2=0.0):
        self.name = name
        self.width, self.height = width, height
        self.to_world = Transform4f(to_world)
        self.fov, self.focal_length = fov, focal_length
        self.near_clip, self.far_clip = near_clip, far_clip
        self.cx, self.cy = cx, cy
        self.k1, self.k2, self.k3, self.k4, self.k5, self.k6 = k1, k2, k3, k4, k5, k6
        self.p1, self.p2 = p1, p2
        if self.fov is None:
            self.fov = focal2fov(self.focal_length, self.width)
        elif self.focal_length is None:
            self.focal_length = fov2focal(self.fov, self.width)
        else:
            raise Exception('CameraSpecs: either FOV or focal length should be set!')

    def viewmat(self) -> np.ndarray:
        '''World-to-camera matrix in the GSplat convention.'''
        return np.array(self.to_world.scale([-1, -1, 1]).inverse().matrix)

    def K(self) -> np.ndarray:
        return np.array([[self.focal_length, 0.0, self.width / 2.0],
                         [0.0, self.focal_length, self.height / 2.0],
                         [0.0, 0.0, 1.0]])

    def to_dict(self, resolution_factor: float = 1.0, pixel_format: str = 'rgb', pixel_filter: str = 'tent') -> dict:
        '''Corresponding sensor dictionary (keys identical to the reference's Mitsuba dictionary).'''
        return {
            'type': 'perspective',
            'principal_point_offset_x': self.cx,
            'principal_point_offset_y': self.cy,
            'fov_axis': 'x',
            'fov': self.fov,
            'to_world': self.to_world,
            'near_clip': self.near_clip,
            'far_clip': self.far_clip,
            'film': {
                'type': 'hdrfilm',
                'rfilter': {'type': pixel_filter},
                'pixel_format': pixel_format,
                'width': int(self.width * resolution_factor),
                'height': int(self.height * resolution_factor),
            }
        }

    @staticmethod
    def from_dict(d: dict, name: str = ''):
        return CameraSpecs(name=name, to_world=d['to_world'], fov=d['fov'], width=d['film']['width'],
                           height=d['film']['height'], cx=d.get('principal_point_offset_x', 0.0),
                           cy=d.get('principal_point_offset_y', 0.0), near_clip=d.get('near_clip', 0.1),
                           far_clip=d.get('far_clip', 10000.0))

    def __repr__(self):
        return "CameraSpecs[\n" + '\n'.join([f"  {k}: {v}" for k, v in self.__dict__.items()]) + "]"


class CameraSpecsIO:
    @staticmethod
    def load(filename: str) -> List[CameraSpecs]:
        raise Exception('Loader not implemented')

    @staticmethod
    def write(specs: List[CameraSpecs], filename: str):
        raise Exception('Loader not implemented')


class JSONCameraSpecsIO(CameraSpecsIO):
    '''
    Load / write sensor dictionaries from json file (e.g. 3DG datasets) -- reference cameras.py:169-217.
    '''
    @staticmethod
    def load(filename: str) -> List[CameraSpecs]:
        with open(filename) as f:
            sensors = json.load(f)
        specs = []
        for sensor in sensors:
            to_world = np.eye(4)
            to_world[:3, :3] = np.array(sensor['rotation']).transpose(0, 1)  # (a no-op transpose, as in the reference)
            to_world[:3, 3] = np.array(sensor['position'])
            to_world = Transform4f(to_world).scale([-1, -1, 1])
            specs.append(CameraSpecs(name=sensor['img_name'], width=sensor['width'], height=sensor['height'],
                                     focal_length=sensor['fx'], to_world=to_world, near_clip=0.01 * 10,
                                     far_clip=100.0))
        return specs

    @staticmethod
    def write(specs: List[CameraSpecs], filename: str):
        sensors = []
        for i, cam in enumerate(specs):
            to_world = cam.to_world @ Transform4f().scale([-1, -1, 1])
            m = np.array(to_world.matrix)
            sensors.append({'rotation': m[:3, :3].transpose(0, 1).tolist(), 'position': m[:3, 3].tolist(),
                            'fx': cam.focal_length, 'fy': cam.focal_length, 'width': cam.width,
                            'height': cam.height, 'id': i, 'img_name': cam.name})
        with open(filename, 'w', encoding='utf-8') as f:
            f.write(json.dumps(sensors, ensure_ascii=False))
