"""`volprim_rf`: radiance-field integrator for ellipsoid primitives, B200 implementation.

Drop-in for volprim/integrators/volprim_rf.py: same plugin name, parameters, defaults and exceptions.
The per-ray loop (closest-hit iteration, SH emission, transmittance, compositing, PRB adjoint;
reference volprim_rf.py:63-192) executes in libvolprim_cuda.so (csrc/vp_trace.cu)."""
from __future__ import annotations

from .. import _cabi
from .base import VolprimIntegratorBase, register_integrator
from .common import Kernel, Properties


class VolumetricPrimitiveRadianceFieldIntegrator(VolprimIntegratorBase):
    '''
    Parameters:
        max_depth (int): Maximum path depth. A value of -1 indicates no limit.
        rr_depth (int): Minimum path depth before enabling the Russian roulette path termination.
        kernel_type (str): one of ['gaussian', 'epanechnikov'].
        srgb_primitives (bool): convert the composited sRGB radiance to linear (default True).
    '''
    integrator_id = _cabi.INTEGRATOR_RF
    attribute_name = 'opacities'

    def __init__(self, props=None):
        props = Properties(props or {})
        super().__init__(props)
        max_depth = int(props.get("max_depth", 64))
        rr_depth = int(props.get('rr_depth', -1))
        if rr_depth < 0 and rr_depth != -1:
            raise Exception("\"rr_depth\" must be set to -1 (infinite) or a value >= 0")
        self.rr_depth = rr_depth if rr_depth != -1 else 0xFFFFFFFF
        # Russian roulette is enabled by the reference iff rr_depth >= 0 and (rr_depth < max_depth or
        # max_depth == -1) (volprim_rf.py:39); every shipped configuration keeps it off.  It runs in the primal pass
        # only (volprim_rf.py:178) and draws `sampler.next_1d()` once per loop iteration: the kernels restate
        # Mitsuba's `independent` sampler (PCG32 stream per ray seeded by TEA(seed, ray index); third-party,
        # unpinned).  `rr_seed` / `rr_skip` (samples a ray drew before sample(): 2 under render(), the film
        # position) select the stream position.
        self.use_rr = rr_depth >= 0 and (rr_depth < max_depth or max_depth == -1)
        self.rr_seed = int(props.get('rr_seed', 0))
        self.rr_skip = int(props.get('rr_skip', 0))
        self.srgb_primitives = props.get('srgb_primitives', True)
        props['kernel_full_range'] = True
        props['kernel_normalized'] = True
        self.kernel = Kernel.factory(props)

    def traverse(self, callback):
        callback.put_parameter("max_depth", self.max_depth, 'NonDifferentiable')
        callback.put_parameter("rr_depth", self.rr_depth, 'NonDifferentiable')
        callback.put_parameter('srgb_primitives', self.srgb_primitives, 'NonDifferentiable')
        callback.put_parameter('kernel_type', self.kernel.type, 'NonDifferentiable')
        callback.put_parameter('hide_emitters', self.hide_emitters, 'NonDifferentiable')

    def parameters_changed(self, keys):
        if 'kernel_type' in keys:
            self.kernel = Kernel.factory({
                'kernel_type': self.kernel.type,
                'kernel_full_range': True,
                'kernel_normalized': True
            })

    def to_string(self):
        return "VolumetricPrimitiveRadianceFieldIntegrator[]"


register_integrator("volprim_rf", lambda props: VolumetricPrimitiveRadianceFieldIntegrator(props))
