"""Bounds-aware Adam and the L1 / L2 / PSNR losses of volprim/optimizers.py, on torch CUDA tensors.

`BoundedAdam` consumes the gradients the adjoint kernel scatters (param.grad after loss.backward()).
Update rule, NaN handling, masking and the half-step-to-bound + moment reset are those of
reference optimizers.py:72-146."""
from __future__ import annotations

import math
from collections import defaultdict

import torch


class BoundedAdam:
    '''
    If a gradient step reaches one of the bounds, the value is moved by half of the distance towards the bound
    instead and the optimizer state for that entry is reset (reference optimizers.py:18-27).
    '''
    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8, mask_updates=False, uniform=False,
                 params: dict = None):
        assert 0 <= beta_1 < 1 and 0 <= beta_2 < 1 and lr > 0 and epsilon > 0
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.mask_updates, self.uniform = mask_updates, uniform
        self.t = defaultdict(lambda: 0)
        self.bounds = {}
        self.lr_default = lr
        self.lr = {}
        self.variables = {}
        self.state = {}
        if params:
            for k, v in params.items():
                self[k] = v

    # -- mi.ad.Optimizer dictionary protocol -------------------------------------------------------
    def __setitem__(self, key, value):
        v = value.detach().clone().requires_grad_(True)
        self.variables[key] = v
        if key not in self.state or self.state[key][0].shape != v.shape:
            self.reset(key)

    def __getitem__(self, key):
        return self.variables[key]

    def __contains__(self, key):
        return key in self.variables

    def keys(self):
        return self.variables.keys()

    def items(self):
        return self.variables.items()

    def set_learning_rate(self, lr):
        if isinstance(lr, dict):
            self.lr.update(lr)
        else:
            self.lr_default = lr

    def set_bounds(self, key, upper=None, lower=None):
        assert lower is None or upper is None or lower < upper, \
            'Upper bound should be higher than lower bound! Did you mix the argument order?'
        self.bounds[key] = (upper, lower)

    def reset(self, key):
        p = self.variables[key]
        self.state[key] = (torch.zeros_like(p), torch.zeros_like(p))
        self.t[key] = 0

    def zero_grad(self):
        for p in self.variables.values():
            p.grad = None

    @torch.no_grad()
    def step(self, active=None):
        active = active or {}
        for k, p in self.variables.items():
            has_mask = k in active
            mask = active.get(k, None)
            self.t[k] += 1
            lr_scale = math.sqrt(1 - self.beta_2 ** self.t[k]) / (1 - self.beta_1 ** self.t[k])
            lr_t = self.lr.get(k, self.lr_default) * lr_scale
            if p.grad is None:
                continue
            if self._fused_step(k, p, lr_t, has_mask):
                continue
            g_p = torch.nan_to_num(p.grad, nan=0.0, posinf=float('inf'), neginf=float('-inf'))  # isnan -> 0 (:88)
            m_tp, v_tp = self.state[k]
            m_t = self.beta_1 * m_tp + (1 - self.beta_1) * g_p
            v_t = self.beta_2 * v_tp + (1 - self.beta_2) * g_p * g_p
            if self.mask_updates:
                nz = g_p != 0.0
                mask = nz if mask is None else (mask & nz)
            if self.mask_updates or has_mask:
                m_t = torch.where(mask, m_t, m_tp)
                v_t = torch.where(mask, v_t, v_tp)
            if self.uniform:
                step = lr_t * m_t / (torch.sqrt(v_t.max()) + self.epsilon)
            else:
                step = lr_t * m_t / (torch.sqrt(v_t) + self.epsilon)
            if self.mask_updates or has_mask:
                step = torch.where(mask, step, torch.zeros_like(step))
            v = p.detach()
            u = v - step
            if k in self.bounds:
                upper, lower = self.bounds[k]
                over = torch.zeros_like(u, dtype=torch.bool)
                if upper is not None:
                    over = u >= upper
                    v = torch.where(over & (v >= upper), torch.full_like(v, upper), v)
                    u = torch.where(over, v + 0.5 * (upper - v), u)
                if lower is not None:
                    over = u <= lower   # NB: as in the reference, this overwrites the upper-bound mask (:129)
                    v = torch.where(over & (v <= lower), torch.full_like(v, lower), v)
                    u = torch.where(over, v - 0.5 * (v - lower), u)
                m_t = torch.where(over, torch.zeros_like(m_t), m_t)
                v_t = torch.where(over, torch.zeros_like(v_t), v_t)
            self.state[k] = (m_t, v_t)
            self.variables[k] = u.detach().requires_grad_(True)

    def _fused_step(self, k, p, lr_t, has_mask) -> bool:
        """One fused CUDA pass (csrc/vp_optim.cu) for the plain variant on CUDA tensors; the masked / uniform variants
        and CPU tensors use the torch ops below (same arithmetic)."""
        if not p.is_cuda or self.mask_updates or self.uniform or has_mask or p.dtype != torch.float32:
            return False
        import ctypes as C
        from . import _cabi
        lib = _cabi.load_library()
        val = p.detach().contiguous().clone()
        g = p.grad.detach().contiguous()
        if g.data_ptr() % 16:      # e.g. a slice of a packed all-reduce buffer: the kernel wants 16-byte aligned rows
            g = g.clone()
        m, v = self.state[k]
        m, v = m.contiguous(), v.contiguous()
        if m.data_ptr() % 16 or v.data_ptr() % 16:
            m, v = m.clone(), v.clone()
        upper, lower = self.bounds.get(k, (None, None))
        ptr = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(p.device):
            rc = lib.vp_bounded_adam_step(val.numel(), ptr(val), ptr(g), ptr(m), ptr(v), float(lr_t), self.beta_1,
                                          self.beta_2, self.epsilon, int(lower is not None), float(lower or 0.0),
                                          int(upper is not None), float(upper or 0.0),
                                          C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream))
        if rc != 0:
            raise _cabi.VolprimCudaError(f"vp_bounded_adam_step failed ({rc})")
        self.state[k] = (m, v)
        self.variables[k] = val.requires_grad_(True)
        return True

    def __repr__(self):
        return ('BoundedAdam[\n  variables = %s,\n  lr = %s,\n  betas = (%g, %g),\n  eps = %g\n  bounds = %s\n]'
                % (list(self.keys()), dict(self.lr, default=self.lr_default), self.beta_1, self.beta_2,
                   self.epsilon, self.bounds))


def l1_loss_grad(reference, image, n_total=None, sums=None):
    """Fused l1(reference, image) + its gradient w.r.t. `image` + the squared error psnr() needs (vp_l1_loss_grad, one
    pass).  Returns (d_image, sums) with sums[0] = sum |diff| / n_total, sums[1] = sum diff^2 / n_total accumulated into
    `sums` (so that the views of a batch film can be handled one at a time with n_total = the whole film's size)."""
    import ctypes as C
    from . import _cabi
    image, reference = image.detach().contiguous(), reference.detach().contiguous()
    if not image.is_cuda or image.dtype != torch.float32 or reference.shape != image.shape:
        raise ValueError("l1_loss_grad: float32 CUDA tensors of equal shape")
    n_total = float(n_total or image.numel())
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float32, device=image.device)
    d_image = torch.empty_like(image)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(image.device):
        rc = _cabi.load_library().vp_l1_loss_grad(image.numel(), ptr(image), ptr(reference), n_total, ptr(d_image), ptr(sums),
                                                  C.c_void_p(torch.cuda.current_stream(image.device).cuda_stream))
    if rc != 0:
        raise _cabi.VolprimCudaError(f"vp_l1_loss_grad failed ({rc})")
    return d_image, sums


def l1(reference, image):
    '''L1 loss function'''
    return (reference - image).abs().mean()


def l2(reference, image):
    '''L2 loss function'''
    return ((reference - image) ** 2).mean()


def psnr(reference, image):
    '''PSNR loss function'''
    return 20 * torch.log(1.0 * torch.rsqrt(l2(reference, image))) / math.log(10)
