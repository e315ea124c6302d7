from . import common
from . import base
from . import volprim_rf
from . import volprim_tomography
from .base import create_integrator, register_integrator
from .common import ADMode, Ellipsoid, EllipsoidsFactory, Kernel, Properties, Ray3f
