"""`volprim_tomography`: absorption-only integrator for ellipsoid primitives, B200 implementation.

Drop-in for volprim/integrators/volprim_tomography.py (same plugin name, parameters, exceptions); the loop
of reference volprim_tomography.py:37-127 executes in libvolprim_cuda.so.  Only a constant environment
emitter is supported (the reference evaluates `scene.environment()` on escape, :105-111)."""
from __future__ import annotations

from .. import _cabi
from .base import VolprimIntegratorBase, register_integrator
from .common import Kernel, Properties


class VolumetricPrimitiveTomographyIntegrator(VolprimIntegratorBase):
    '''
    Parameters:
        max_depth (int): Maximum path depth. A value of -1 indicates no limit.
        kernel_type (str): one of ['gaussian', 'epanechnikov'].
        hide_emitters (bool): Hide emitters from indirect light sampling.
    '''
    integrator_id = _cabi.INTEGRATOR_TOMO
    attribute_name = 'sigma_t'

    def __init__(self, props=None):
        props = Properties(props or {})
        super().__init__(props)
        props['kernel_full_range'] = True
        props['kernel_normalized'] = False
        self.kernel = Kernel.factory(props)

    def traverse(self, callback):
        callback.put_parameter("max_depth", self.max_depth, 'NonDifferentiable')
        callback.put_parameter('kernel_type', self.kernel.type, 'NonDifferentiable')
        callback.put_parameter('hide_emitters', self.hide_emitters, 'NonDifferentiable')

    def parameters_changed(self, keys):
        if 'kernel_type' in keys:
            self.kernel = Kernel.factory({
                'kernel_type': self.kernel.type,
                'kernel_full_range': True,
                'kernel_normalized': False
            })

    def to_string(self):
        return "VolumetricPrimitiveTomographyIntegrator[]"


register_integrator("volprim_tomography", lambda props: VolumetricPrimitiveTomographyIntegrator(props))
