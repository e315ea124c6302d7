"""volprim_balance_b200 -- B200-native implementation of volprim's per-ray volumetric-primitive integration
(the `volprim_rf` / `volprim_tomography` integrators of gitmon/volprim-balance).

    import volprim_balance_b200 as volprim
    scene = volprim.load_dict({'type': 'scene', 'integrator': {'type': 'volprim_rf', ...}, 'primitives': {...}, ...})
    image = volprim.render(scene, sensor=0, spp=4)

Python host code with PyTorch tensors at the boundary calls libvolprim_cuda.so (include/volprim_cuda.h)
through ctypes; there is no Dr.Jit / Triton / OptiX and no CPU fallback on the hot path.
"""
from . import _cabi
from ._cabi import VolprimCudaError
from . import transforms
from .transforms import ScalarTransform4f, Transform4f
from . import integrators
from .integrators import ADMode, Ellipsoid, EllipsoidsFactory, Properties, Ray3f
from . import scene as _scene
from .scene import (BatchSensor, EllipsoidsShape, PerspectiveSensor, Scene, SceneParameters, load_dict, render,
                    render_to_host, traverse)
from . import cameras, io, optimizers, parallel, training, utils, synthetic

__version__ = "0.2.0"
