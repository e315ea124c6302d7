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
from ._cabi import VolprimCudaError, vp_camera, vp_hit_record, vp_params, vp_ray_source, vp_stats


def _ptr(t):
    """Device address of a tensor -- or a raw address (int), for gradient rows that live inside a larger buffer."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t) if t else None
    return C.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


@dataclass
class RaySource:
    """Rays of a render call: explicit (o, d, maxt) tensors, or a perspective sensor evaluated inside the trace
    kernels (`camera` = vp_camera, pixel-major then sample; `rows` = (first row, row count) of a band)."""
    o: torch.Tensor | None = None
    d: torch.Tensor | None = None
    maxt: torch.Tensor | None = None
    camera: vp_camera | None = None
    spp: int = 1
    jitter: torch.Tensor | None = None
    rows: tuple | None = None

    @property
    def n_rays(self) -> int:
        if self.camera is not None:
            rows = self.rows[1] if self.rows else self.camera.height
            return self.camera.width * rows * self.spp
        return self.o.shape[0]

    def to_c(self) -> vp_ray_source:
        rs = vp_ray_source()
        if self.camera is not None:
            rs.camera = C.pointer(self.camera)
            rs.spp = int(self.spp)
            rs.jitter = self.jitter.data_ptr() if self.jitter is not None else None
            if self.rows:
                rs.row_begin, rs.row_count = int(self.rows[0]), int(self.rows[1])
        else:
            rs.ray_o, rs.ray_d = self.o.data_ptr(), self.d.data_ptr()
            rs.ray_maxt = self.maxt.data_ptr() if self.maxt is not None else None
            rs.spp = 1
        return rs


class HitRecord:
    """Ordered hit lists of a primal pass in compressed-row form (vp_hit_record): 4 bytes per recorded hit (20 with
    `with_state`: the primal also leaves every hit's colour and transmittance, and the adjoint's ray-major pass becomes
    the PRB recurrence alone), 8 per ray.  `usable()` reads two device counters (synchronises the stream once)."""

    def __init__(self, n_rays: int, n_prims: int, capacity: int, id_cap: int, device, with_state: bool = False,
                 dense: bool = False):
        self.n_rays, self.n_prims = n_rays, n_prims
        self.capacity, self.id_cap, self.dense = int(max(capacity, 1)), int(id_cap), bool(dense)
        self.total = torch.zeros(2, dtype=torch.int64, device=device)
        self._usable = None
        if dense:
            # hit-major block [id_cap, n_rays] in this record's own buffers: no compaction pass, coalesced replay
            self.ray_offsets = None
            self.ids = torch.empty((id_cap, n_rays), dtype=torch.int32, device=device)
            self.state = torch.empty((id_cap, n_rays, 4), dtype=torch.float32, device=device)
            self.counts = torch.empty(n_rays, dtype=torch.int32, device=device)
            return
        self.counts = None
        self.ray_offsets = torch.empty(n_rays + 1, dtype=torch.int64, device=device)
        self.ids = torch.empty(self.capacity, dtype=torch.int32, device=device)
        self.state = torch.empty((self.capacity, 4), dtype=torch.float32, device=device) if with_state else None

    def to_c(self) -> vp_hit_record:
        r = vp_hit_record()
        r.ray_offsets = self.ray_offsets.data_ptr() if self.ray_offsets is not None else None
        r.ids = self.ids.data_ptr()
        r.state = self.state.data_ptr() if self.state is not None else None
        r.counts = self.counts.data_ptr() if self.counts is not None else None
        r.total = self.total.data_ptr()
        r.capacity, r.id_cap, r.dense = self.capacity, self.id_cap, int(self.dense)
        return r

    def totals(self):
        t = self.total.tolist()
        return int(t[0]), int(t[1])

    def usable(self) -> bool:
        if self._usable is None:
            entries, cut = self.totals()
            self._usable = entries <= self.capacity and cut == 0
        return self._usable

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.ids, self.ray_offsets, self.state, self.counts) if t is not None)

    def lists(self):
        """Python view for tests: list of per-ray id arrays (host)."""
        ids = self.ids.cpu().numpy()
        if self.dense:
            cnt = self.counts.cpu().numpy()
            return [ids[:cnt[r], r] for r in range(self.n_rays)]
        off = self.ray_offsets.cpu().numpy()
        return [ids[off[r]:off[r + 1]] for r in range(self.n_rays)]


@dataclass
class TraceResult:
    rgb: torch.Tensor        # [R, 3]
    beta: torch.Tensor       # [R] final throughput
    nhits: torch.Tensor      # [R] int32 (bit pattern of uint32)
    hit_ids: torch.Tensor | None = None  # [cap, R] int32, -1 padded (hit-major: coalesced per hit index)
    record: HitRecord | None = None      # compressed-row lists (render_forward(record=True))


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
        self.hits_per_ray_estimate = 48.0   # sizes the next hit record; follows the records actually produced
        self.state_budget_bytes = 8 << 30   # records above it keep ids only (4 B / hit) and the adjoint re-shades
        self.dense_budget_bytes = 16 << 30  # a 1080p view at a cap of 128 hits takes 6.4 GB as a dense state record

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
        # the kernel writes only the entries a ray has; the dense debug / parity view is pre-filled with -1
        ids = torch.full((record_cap, R), -1, dtype=torch.int32, device=self.device) if record_cap > 0 else None
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

    # -- sensor-fused / record-replay path (what render() uses) ------------------------------------------
    def set_option(self, name: str, value: int):
        _cabi.check(self._lib.vp_set_option(self._h, name.encode(), int(value)), self._h)

    def new_record(self, n_rays: int, id_cap: int, capacity: int | None = None, with_state: bool | None = None,
                   dense: bool | None = None) -> HitRecord:
        """Kind of record (None = decide from the memory budgets, fastest first):
          dense       hit-major [id_cap, n_rays] block with per-hit (colour, transmittance): no compaction pass, coalesced
                      replay; 24 B * id_cap * n_rays incl. the adjoint's slot array (`dense_budget_bytes`, default 16 GiB)
          rows+state  compressed rows, 20 B per hit (`state_budget_bytes`, default 8 GiB)
          rows        compressed rows, 4 B per hit; the adjoint shades every hit again."""
        if capacity is None:
            capacity = int(n_rays * min(float(id_cap), self.hits_per_ray_estimate * 1.3 + 4.0)) + 4096
        capacity = min(capacity, (1 << 32) - 1)
        if dense is None:
            dense = with_state is not False and self.sh_floats > 0 and id_cap * n_rays * 24 <= self.dense_budget_bytes
        if dense:
            return HitRecord(n_rays, self.n, capacity, id_cap, self.device, dense=True)
        if with_state is None:      # (only volprim_rf hits carry a colour; tomography records are the id lists)
            with_state = self.sh_floats > 0 and capacity * 16 <= self.state_budget_bytes
        return HitRecord(n_rays, self.n, capacity, id_cap, self.device, with_state=with_state)

    def record_bytes(self, n_rays: int, id_cap: int) -> int:
        """Memory the record new_record() would choose for these rays takes (budgeting of multi-view renders)."""
        entries = int(n_rays * min(float(id_cap), self.hits_per_ray_estimate * 1.3 + 4.0)) + 4096
        if self.sh_floats > 0 and id_cap * n_rays * 24 <= self.dense_budget_bytes:
            return id_cap * n_rays * 20 + n_rays * 4
        return entries * (20 if entries * 16 <= self.state_budget_bytes else 4) + n_rays * 8

    def render_forward(self, params: vp_params, rays: RaySource, record=None, id_cap: int = 0,
                       want_beta: bool = True, want_nhits: bool = True) -> TraceResult:
        """vp_render_forward.  `record`: None / False, True (a HitRecord sized from the running estimate is
        allocated) or a HitRecord to fill."""
        R = rays.n_rays
        rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        beta = torch.empty((R,), dtype=torch.float32, device=self.device) if want_beta else None
        nhits = torch.empty((R,), dtype=torch.int32, device=self.device) if want_nhits else None
        rec = None
        if record is True:
            rec = self.new_record(R, id_cap)
        elif isinstance(record, HitRecord):
            rec = record
            rec._usable = None
        crec = rec.to_c() if rec is not None else None
        crays = rays.to_c()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_render_forward(self._h, C.byref(params), C.byref(crays), R, _ptr(rgb), _ptr(beta),
                                                    _ptr(nhits), C.byref(crec) if crec is not None else None,
                                                    self._stream()), self._h)
        return TraceResult(rgb, beta, nhits, None, rec)

    def _grad_out(self, out):
        if out is not None:
            return out
        g_data = torch.zeros(self.n * 10, dtype=torch.float32, device=self.device)
        g_attr = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        g_sh = torch.zeros(self.n * self.sh_floats, dtype=torch.float32, device=self.device) if self.sh_floats else None
        return g_data, g_attr, g_sh

    def render_adjoint(self, params: vp_params, rays: RaySource, dL, state_in, record: HitRecord, out=None):
        """vp_render_adjoint: replay `record`; volprim_rf gathers per primitive (no global reductions).  The caller
        checks record.usable() first (an unusable record makes the kernels return without touching the gradients)."""
        g_data, g_attr, g_sh = self._grad_out(out)
        R = rays.n_rays
        dL, st = _f32(dL, self.device).reshape(-1, 3), _f32(state_in, self.device).reshape(-1, 3)
        if dL.shape[0] != R or st.shape[0] != R:
            raise ValueError("d_L / state_in must have one RGB triple per ray")
        crec, crays = record.to_c(), rays.to_c()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_render_adjoint(self._h, C.byref(params), C.byref(crays), R, _ptr(dL), _ptr(st),
                                                    C.byref(crec), _ptr(g_data), _ptr(g_attr), _ptr(g_sh),
                                                    self._stream()), self._h)
        return g_data, g_attr, g_sh

    def adjoint_begin(self, params: vp_params, rays: RaySource, dL, state_in, record: HitRecord, out):
        g_data, g_attr, g_sh = out
        R = rays.n_rays
        dL, st = _f32(dL, self.device).reshape(-1, 3), _f32(state_in, self.device).reshape(-1, 3)
        crec, crays = record.to_c(), rays.to_c()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_adjoint_begin(self._h, C.byref(params), C.byref(crays), R, _ptr(dL), _ptr(st),
                                                   C.byref(crec), _ptr(g_data), _ptr(g_attr), _ptr(g_sh),
                                                   self._stream()), self._h)

    def adjoint_finish(self, params: vp_params, rays: RaySource, record: HitRecord, prim_begin: int, prim_end: int, out):
        g_data, g_attr, g_sh = out
        crec, crays = record.to_c(), rays.to_c()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.vp_adjoint_finish(self._h, C.byref(params), C.byref(crays), rays.n_rays, C.byref(crec),
                                                    int(prim_begin), int(prim_end), _ptr(g_data), _ptr(g_attr),
                                                    _ptr(g_sh), self._stream()), self._h)

    def note_record(self, record: HitRecord):
        """Feed the size of a finished record back into the estimate that sizes the next one."""
        entries, _ = record.totals()
        if record.n_rays:
            self.hits_per_ray_estimate = max(4.0, entries / record.n_rays)

    # -- film ------------------------------------------------------------------------------------
    def film_splat(self, width, height, spp, rfilter: int, jitter, radiance, accum):
        with torch.cuda.device(self.device):
            rc = self._lib.vp_film_splat(width, height, spp, rfilter, _ptr(jitter), _ptr(radiance), _ptr(accum), self._stream())
        if rc:
            raise VolprimCudaError(f"vp_film_splat failed ({rc})")

    def film_develop(self, width, height, accum, image_view):
        """image_view: [H, W, 3] float32 view with unit element stride in the last two dims (a column block of a
        batch film is fine: the row stride is passed on)."""
        assert image_view.stride(2) == 1 and image_view.stride(1) == 3
        with torch.cuda.device(self.device):
            rc = self._lib.vp_film_develop(width, height, _ptr(accum), _ptr(image_view), image_view.stride(0), self._stream())
        if rc:
            raise VolprimCudaError(f"vp_film_develop failed ({rc})")

    def film_adjoint(self, width, height, spp, rfilter: int, jitter, accum, d_image_view):
        assert d_image_view.stride(2) == 1 and d_image_view.stride(1) == 3
        dL = torch.empty((width * height * spp, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.vp_film_adjoint(width, height, spp, rfilter, _ptr(jitter), _ptr(accum), _ptr(d_image_view),
                                           d_image_view.stride(0), _ptr(dL), self._stream())
        if rc:
            raise VolprimCudaError(f"vp_film_adjoint failed ({rc})")
        return dL

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
