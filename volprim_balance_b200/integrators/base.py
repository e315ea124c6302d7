"""Common host logic of the two integrator plugins: parameter marshalling, the sample() contract of
mitsuba.ad.integrators.common.RBIntegrator, and the render / render_backward drivers around it."""
from __future__ import annotations

import torch

from .. import _cabi
from .common import ADMode, Kernel, Properties, Ray3f, get_ellipsoids_shape

_REGISTRY = {}


def register_integrator(name, factory):
    """mi.register_integrator(name, lambda props: Cls(props))"""
    _REGISTRY[name] = factory


def create_integrator(props) -> "VolprimIntegratorBase":
    props = Properties(props)
    name = props.get('type')
    if name not in _REGISTRY:
        raise Exception(f"Unknown integrator plugin '{name}' (this package provides {sorted(_REGISTRY)})")
    return _REGISTRY[name](props)


class VolprimIntegratorBase:
    """Shared parts of volprim_rf.py:23-61 and volprim_tomography.py:24-35."""

    integrator_id = None
    attribute_name = None  # scalar per-primitive attribute the transmission model reads

    def __init__(self, props=None):
        props = Properties(props or {})
        self._props = props
        max_depth = int(props.get("max_depth", 64))
        if max_depth < 0 and max_depth != -1:
            raise Exception('"max_depth" must be set to -1 (infinite) or a value >= 0')
        # Map -1 (infinity) to 2^32-1 bounces
        self.max_depth = max_depth if max_depth != -1 else 0xFFFFFFFF
        self.hide_emitters = bool(props.get('hide_emitters', False))
        self.last = None  # TraceResult of the most recent primal sample() (beta, nhits, hit lists)
        self.record_hits = False
        self.record_cap = 0

    # ---- C-ABI parameter block -------------------------------------------------------------------
    def _vp_params(self, scene, image=None) -> _cabi.vp_params:
        p = _cabi.vp_params()
        p.integrator = self.integrator_id
        p.kernel = self.kernel.cabi_id
        p.max_depth = self.max_depth
        p.srgb_primitives = int(getattr(self, 'srgb_primitives', False))
        p.hide_emitters = int(self.hide_emitters)
        p.t_cutoff, p.eps_advance = 0.01, 1e-4
        env = scene.environment_radiance() if scene is not None else (0.0, 0.0, 0.0)
        p.env[0], p.env[1], p.env[2] = env
        p.image_width, p.image_height = image if image else (0, 0)
        # Russian roulette (volprim_rf.py:39,177-183): primal pass only, PCG32 stream per ray
        p.use_rr = int(bool(getattr(self, 'use_rr', False)))
        p.rr_depth = int(getattr(self, 'rr_depth', 0xFFFFFFFF)) & 0xFFFFFFFF
        p.rr_seed = int(getattr(self, 'rr_seed', 0)) & 0xFFFFFFFF
        p.rr_skip = int(getattr(self, 'rr_skip', 0)) & 0xFFFFFFFF
        return p

    def _cap(self) -> int:
        if self.record_cap:
            return self.record_cap
        return int(min(self.max_depth, 256))

    # ---- RBIntegrator.sample ---------------------------------------------------------------------
    def sample(self, mode, scene, sampler, ray, δL=None, state_in=None, active=True, **kwargs):
        """Same contract as the reference plugins (volprim_rf.py:103-192, volprim_tomography.py:47-127):
        returns (spectrum, valid, aovs, state_out).  `ray` carries CUDA tensors; `sampler` is unused: the only
        consumer of random numbers, Russian roulette, draws from the kernel's own PCG32 restatement (rr_seed).
        Backward mode scatters the parameter gradients into the shape's `.grad` buffers."""
        shape = get_ellipsoids_shape(scene)
        accel = shape.accel()
        image = kwargs.get('image')
        params = self._vp_params(scene, image)
        o, d, maxt = ray.o, ray.d, ray.maxt
        idx = None
        if isinstance(active, torch.Tensor):
            if not bool(active.all()):
                idx = torch.nonzero(active.reshape(-1), as_tuple=False).reshape(-1)
                o, d = o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]
                maxt = None if maxt is None else maxt.reshape(-1)[idx]
                params.image_width = params.image_height = 0
        elif not active:
            R = ray.o.reshape(-1, 3).shape[0]
            z = torch.zeros((R, 3), dtype=torch.float32, device=ray.o.device)
            return z, True, [], z
        attr = shape.attribute(self.attribute_name)
        shape.bind(attr_name=self.attribute_name, with_sh=self.integrator_id == _cabi.INTEGRATOR_RF)

        def expand(x):
            if idx is None:
                return x
            full = torch.zeros((ray.o.reshape(-1, 3).shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            full[idx] = x
            return full

        if mode == ADMode.Primal:
            rec = self._cap() if (self.record_hits or kwargs.get('record', False)) else 0
            res = accel.trace_forward(params, o, d, maxt, record_cap=rec)
            self.last = res
            L = expand(res.rgb)
            return L, True, [], L
        if mode == ADMode.Backward:
            if δL is None or state_in is None:
                raise Exception("Backward mode needs δL and state_in (the primal call's state_out)")
            dL = δL.reshape(-1, 3) if idx is None else δL.reshape(-1, 3)[idx]
            st = state_in.reshape(-1, 3) if idx is None else state_in.reshape(-1, 3)[idx]
            hit_ids = kwargs.get('hit_ids')
            hit_counts = kwargs.get('hit_counts')
            if idx is not None:
                hit_ids = hit_counts = None
            accel.trace_adjoint(params, o, d, maxt, dL, st, hit_ids, hit_counts, out=shape.grad_buffers(self.attribute_name))
            return δL, True, [], state_in
        raise NotImplementedError("forward-mode AD (dr.forward_to) is not provided by the CUDA adjoint")

    def to_string(self):
        return f"{type(self).__name__}[]"

    __repr__ = to_string
