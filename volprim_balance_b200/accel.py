"""Torch-tensor front end of the per-GPU acceleration context (`vp_ctx` in include/volprim_cuda.h).

This is the piece that stands where Mitsuba's scene-side acceleration structure and the Dr.Jit megakernel
launch stood in the reference: it owns the LBVH over the bounding ellipsoids and
runs the traversal / evaluation / adjoint kernels.  Tensors are float32 CUDA tensors in the reference
layouts; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _cabi
from ._cabi import VolprimCudaError, vp_camera, vp_params, vp_stats


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


@dataclass
class TraceResult:
    rgb: torch.Tensor        # [R, 3]
    beta: torch.Tensor       # [R] final throughput
    nhits: torch.Tensor      # [R] int32 (bit pattern of uint32)
    hit_ids: torch.Tensor | None = None  # [cap, R] int32, -1 padded (hit-major: coalesced per hit index)


class EllipsoidAccel:
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise VolprimCudaError("volprim_balance_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _cabi.load_library()
        h = C.c_void_p()
        _cabi.check(self._lib.vp_create(self.device.index, C.byref(h)), None)
        self._h = h
        self.n = 0
        self.sh_floats = 0
        self.built = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- primitives / BVH ------------------------------------------------------------------------
    def set_primitives(self, data, attr=None, sh=None, extent: float = 3.0):
        """data: flat [N*10] or [N,10] (center, scale, quat i-j-k-r); attr: [N] opacities | sigma_t;
        sh: [N, C] or flat."""
        data = _f32(data, self.device).reshape(-1)
        if data.numel() % 10:
            raise ValueError("primitives.data must hold 10 floats per ellipsoid")
        n = data.numel() // 10
        attr_t = None if attr is None else _f32(attr, self.device).reshape(-1)
        if attr_t is not None and attr_t.numel() != n:
            raise ValueError(f"attribute has {attr_t.numel()} entries for {n} ellipsoids")
        sh_t, shf = None, 0
        if sh is not None:
            sh_t = _f32(sh, self.device).reshape(-1)
            if n and sh_t.numel() % n:
                raise ValueError("sh_coeffs size is not a multiple of the primitive count")
            shf = sh_t.numel() // n if n else 0
            if shf == 0:
                sh_t = None
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_set_primitives(self._h, n, _ptr(data), _ptr(attr_t), _ptr(sh_t), shf,
                                                    float(extent), self._stream()), self._h)
        if n != self.n or shf != self.sh_floats:
            self.built = False
        self.n, self.sh_floats = n, shf

    def build(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_build(self._h, self._stream()), self._h)
        self.built = True

    def refit(self):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_refit(self._h, self._stream()), self._h)

    # -- tracing ---------------------------------------------------------------------------------
    def trace_forward(self, params: vp_params, o, d, maxt=None, record_cap: int = 0) -> TraceResult:
        o, d = _f32(o, self.device).reshape(-1, 3), _f32(d, self.device).reshape(-1, 3)
        R = o.shape[0]
        if d.shape[0] != R:
            raise ValueError("ray origins and directions differ in count")
        maxt_t = None if maxt is None else _f32(maxt, self.device).reshape(-1)
        rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        beta = torch.empty((R,), dtype=torch.float32, device=self.device)
        nhits = torch.empty((R,), dtype=torch.int32, device=self.device)
        ids = torch.empty((record_cap, R), dtype=torch.int32, device=self.device) if record_cap > 0 else None
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_trace_forward(self._h, C.byref(params), R, _ptr(o), _ptr(d), _ptr(maxt_t),
                                                   _ptr(rgb), _ptr(beta), _ptr(nhits), _ptr(ids), record_cap, 1, R,
                                                   self._stream()), self._h)
        return TraceResult(rgb, beta, nhits, ids)

    def trace_adjoint(self, params: vp_params, o, d, maxt, dL, state_in, hit_ids=None, hit_counts=None,
                      out=None):
        """Returns (g_data [N*10], g_attr [N], g_sh [N*C] | None); accumulates into `out` if given."""
        o, d = _f32(o, self.device).reshape(-1, 3), _f32(d, self.device).reshape(-1, 3)
        R = o.shape[0]
        maxt_t = None if maxt is None else _f32(maxt, self.device).reshape(-1)
        dL = _f32(dL, self.device).reshape(-1, 3)
        st = _f32(state_in, self.device).reshape(-1, 3)
        if dL.shape[0] != R or st.shape[0] != R:
            raise ValueError("d_L / state_in must have one RGB triple per ray")
        if out is None:
            g_data = torch.zeros(self.n * 10, dtype=torch.float32, device=self.device)
            g_attr = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            g_sh = torch.zeros(self.n * self.sh_floats, dtype=torch.float32, device=self.device) if self.sh_floats else None
        else:
            g_data, g_attr, g_sh = out
        cap = 0
        if hit_ids is not None:
            if hit_counts is None:
                raise ValueError("hit_ids needs hit_counts")
            cap = hit_ids.shape[0]
            assert hit_ids.is_contiguous() and hit_ids.shape[1] == R
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_trace_adjoint(self._h, C.byref(params), R, _ptr(o), _ptr(d), _ptr(maxt_t),
                                                   _ptr(dL), _ptr(st), _ptr(hit_ids), _ptr(hit_counts), cap, 1, R,
                                                   _ptr(g_data), _ptr(g_attr), _ptr(g_sh), self._stream()), self._h)
        return g_data, g_attr, g_sh

    def raygen_perspective(self, cam: vp_camera, spp: int = 1, jitter=None):
        total = cam.width * cam.height * spp
        o = torch.empty((total, 3), dtype=torch.float32, device=self.device)
        d = torch.empty((total, 3), dtype=torch.float32, device=self.device)
        maxt = torch.empty((total,), dtype=torch.float32, device=self.device)
        jit = None if jitter is None else _f32(jitter, self.device).reshape(-1)
        if jit is not None and jit.numel() != 2 * total:
            raise ValueError("jitter must hold 2 floats per sample")
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_raygen_perspective(self._h, C.byref(cam), spp, _ptr(jit), _ptr(o), _ptr(d),
                                                        _ptr(maxt), self._stream()), self._h)
        return o, d, maxt

    def stats(self) -> dict:
        s = vp_stats()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_get_stats(self._h, C.byref(s), self._stream()), self._h)
        return {k: int(getattr(s, k)) for k, _ in vp_stats._fields_}

    def debug_bvh(self):
        """(nodes [N-1,16] float32, perm [N] int32) copies of the BVH -- test introspection."""
        n_int = max(self.n - 1, 0)
        nodes = torch.empty((n_int, 16), dtype=torch.float32, device=self.device)
        perm = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        ni = C.c_int64()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_debug_bvh(self._h, _ptr(nodes), _ptr(perm), C.byref(ni), self._stream()), self._h)
        assert ni.value == n_int
        return nodes, perm
