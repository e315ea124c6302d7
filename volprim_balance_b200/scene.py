"""Scene-side host objects: the ellipsoids shape, perspective / batch sensors, `load_dict`, `traverse`,
`render`.  These mirror the slice of Mitsuba's Python API the reference's callers use around the two
integrators (examples/render_3dg_asset.py:53-77, examples/refine_3dg_dataset.py:66-189,
examples/optimize_volume.py:132-249), with PyTorch CUDA tensors in place of Dr.Jit arrays.  No arithmetic
of the per-ray path happens here: `render()` hands the sensor to `vp_render_forward` (rays are generated inside the
trace kernel), reconstructs the film with `vp_film_*`, and gets parameter gradients from `vp_render_adjoint`
(replay of the primal's compressed hit records) or `vp_trace_adjoint` (re-trace).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi
from .accel import EllipsoidAccel
from .integrators.base import VolprimIntegratorBase, create_integrator
from .integrators.common import ADMode, Ray3f
from .transforms import Transform4f

SENSOR_TYPES = ('perspective',)
EMITTER_TYPES = ('constant',)
ELLIPSOID_TYPES = ('ellipsoids', 'ellipsoidsmesh')


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _cabi.VolprimCudaError("volprim_balance_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device('cuda', torch.cuda.current_device())


class EllipsoidsShape:
    """Shape of ShapeType.Ellipsoids (reference common.py:22-33): `data` flat [N*10] + named attributes.

    Built from {'type': 'ellipsoidsmesh', 'filename': ply} (examples/render_3dg_asset.py:59-62) or from
    tensors centers [N,3], scales [N,3], quaternions [N,4] + extra attributes + 'extent'
    (examples/optimize_volume.py:148-156)."""
    is_ellipsoids = True

    def __init__(self, props: dict, shape_id: str = 'primitives', device=None):
        self.id = shape_id
        self.plugin = props.get('type', 'ellipsoidsmesh')
        self.device = _device(device)
        self.extent = float(props.get('extent', 3.0))
        if 'filename' in props:
            from . import io as vio
            d = vio.load_ellipsoids_ply(props['filename'])
        else:
            d = {k: v for k, v in props.items() if k not in ('type', 'extent', 'id')}
        for k in ('centers', 'scales', 'quaternions'):
            if k not in d:
                raise Exception(f"ellipsoids shape '{shape_id}': missing '{k}'")
        t = lambda x: torch.as_tensor(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x),
                                      dtype=torch.float32).to(self.device)
        c, s, q = t(d['centers']).reshape(-1, 3), t(d['scales']).reshape(-1, 3), t(d['quaternions']).reshape(-1, 4)
        if not (c.shape[0] == s.shape[0] == q.shape[0]):
            raise Exception("centers / scales / quaternions differ in primitive count")
        self.data = torch.cat([c, s, q], dim=1).reshape(-1).contiguous()
        self.attributes = OrderedDict()
        for k, v in d.items():
            if k in ('centers', 'scales', 'quaternions'):
                continue
            if isinstance(v, (np.ndarray, torch.Tensor, list, tuple)):
                self.attributes[k] = t(v).reshape(-1).contiguous()
        self._accel = None
        self._bound = None      # (attr_name, with_sh) currently uploaded
        self._dirty = True
        self._topology_n = -1
        self.rebuild_policy = 'rebuild'  # 'rebuild' = reference behaviour on params.update(); 'refit' keeps topology
        self.grad = {}

    @property
    def count(self) -> int:
        return self.data.numel() // 10

    def has_attribute(self, name):
        return name in self.attributes

    def attribute(self, name):
        if name is None:
            return None
        if name not in self.attributes:
            raise Exception(f"Requested ellipsoid attribute '{name}' not found on shape '{self.id}'")
        return self.attributes[name]

    def parameters_changed(self, keys=()):
        self._dirty = True

    def accel(self) -> EllipsoidAccel:
        if self._accel is None:
            self._accel = EllipsoidAccel(self.device)
        return self._accel

    def bind(self, attr_name, with_sh=True):
        """Upload the current parameter values and (re)build the LBVH when something changed."""
        sh = self.attributes.get('sh_coeffs') if with_sh else None
        key = (attr_name, sh is not None)
        if not self._dirty and self._bound == key:
            return
        acc = self.accel()
        acc.set_primitives(self.data, self.attribute(attr_name) if attr_name else None, sh, self.extent)
        if self.rebuild_policy == 'refit' and acc.built and self._topology_n == self.count:
            acc.refit()
        else:
            acc.build()
            self._topology_n = self.count
        self._bound, self._dirty = key, False

    def grad_buffers(self, attr_name):
        n = self.count
        shf = self.attributes['sh_coeffs'].numel() // n if ('sh_coeffs' in self.attributes and n) else 0
        if 'data' not in self.grad or self.grad['data'].numel() != n * 10:
            self.grad['data'] = torch.zeros(n * 10, dtype=torch.float32, device=self.device)
        if attr_name not in self.grad or self.grad[attr_name].numel() != n:
            self.grad[attr_name] = torch.zeros(n, dtype=torch.float32, device=self.device)
        g_sh = None
        if shf and self._bound and self._bound[1]:
            if 'sh_coeffs' not in self.grad or self.grad['sh_coeffs'].numel() != n * shf:
                self.grad['sh_coeffs'] = torch.zeros(n * shf, dtype=torch.float32, device=self.device)
            g_sh = self.grad['sh_coeffs']
        return self.grad['data'], self.grad[attr_name], g_sh

    def zero_grad(self):
        for g in self.grad.values():
            g.zero_()


class PerspectiveSensor:
    """Mitsuba `perspective` sensor as configured by CameraSpecs.to_dict (reference cameras.py:114-137)."""

    def __init__(self, props: dict):
        film = props.get('film', {})
        self.width, self.height = int(film.get('width', 768)), int(film.get('height', 576))
        rf = film.get('rfilter', film.get('filter', {'type': 'gaussian'}))
        self.rfilter = rf.get('type', 'gaussian') if isinstance(rf, dict) else str(rf)
        if props.get('fov_axis', 'x') != 'x':
            raise Exception("perspective sensor: only fov_axis='x' is supported (what CameraSpecs.to_dict emits)")
        self.fov = float(props.get('fov', 40.0))
        self.to_world = Transform4f(props.get('to_world', Transform4f()))
        self.near_clip = float(props.get('near_clip', 1e-2))
        self.far_clip = float(props.get('far_clip', 1e4))
        self.cx = float(props.get('principal_point_offset_x', 0.0))
        self.cy = float(props.get('principal_point_offset_y', 0.0))

    def vp_camera(self) -> _cabi.vp_camera:
        cam = _cabi.vp_camera()
        m = np.asarray(self.to_world.matrix, np.float32)
        for r in range(3):
            for c in range(4):
                cam.to_world[4 * r + c] = float(m[r, c])
        cam.fov_x_deg, cam.near_clip, cam.far_clip = self.fov, self.near_clip, self.far_clip
        cam.cx, cam.cy, cam.width, cam.height = self.cx, self.cy, self.width, self.height
        return cam

    def film_size(self):
        return self.width, self.height


class BatchSensor:
    """Mitsuba `batch` sensor: N perspective sensors laid side by side on one wide film
    (reference examples/refine_3dg_dataset.py:96-107)."""

    def __init__(self, props: dict):
        self.sensors = [v if isinstance(v, PerspectiveSensor) else PerspectiveSensor(v)
                        for k, v in props.items() if isinstance(v, (dict, PerspectiveSensor)) and k != 'film'
                        and (isinstance(v, PerspectiveSensor) or v.get('type') in SENSOR_TYPES)]
        film = props.get('film', {})
        rf = film.get('rfilter', film.get('filter', {'type': 'tent'}))
        self.rfilter = rf.get('type', 'tent') if isinstance(rf, dict) else str(rf)
        if not self.sensors:
            raise Exception("batch sensor without child sensors")
        w = int(film.get('width', sum(s.width for s in self.sensors)))
        self.height = int(film.get('height', self.sensors[0].height))
        if w % len(self.sensors):
            raise Exception("batch film width must be a multiple of the sensor count")
        self.view_width = w // len(self.sensors)
        self.width = w

    def film_size(self):
        return self.width, self.height


class Scene:
    def __init__(self):
        self._shapes, self._sensors, self.integrator = [], [], None
        self._env = None

    def shapes(self):
        return self._shapes

    def sensors(self):
        return self._sensors

    def environment_radiance(self):
        return self._env if self._env is not None else (0.0, 0.0, 0.0)

    def ellipsoids(self) -> EllipsoidsShape:
        from .integrators.common import get_ellipsoids_shape
        return get_ellipsoids_shape(self)


def load_dict(d: dict, device=None):
    """mi.load_dict for the object kinds on the hot path: scene, ellipsoids shape, perspective / batch
    sensor, constant emitter, and the two volprim integrators."""
    kind = d.get('type')
    if kind == 'scene':
        sc = Scene()
        for k, v in d.items():
            if k == 'type':
                continue
            if isinstance(v, (PerspectiveSensor, BatchSensor)):
                sc._sensors.append(v)
                continue
            if not isinstance(v, dict):
                continue
            t = v.get('type')
            if k == 'integrator' or t in ('volprim_rf', 'volprim_tomography'):
                sc.integrator = create_integrator(v)
            elif t in ELLIPSOID_TYPES:
                sc._shapes.append(EllipsoidsShape(v, shape_id=k, device=device))
            elif t in SENSOR_TYPES:
                sc._sensors.append(PerspectiveSensor(v))
            elif t == 'batch':
                sc._sensors.append(BatchSensor(v))
            elif t == 'constant':
                rad = v.get('radiance', 1.0)
                if isinstance(rad, dict):
                    rad = rad.get('value', 1.0)
                rad = np.broadcast_to(np.asarray(rad, np.float32), (3,))
                sc._env = tuple(float(x) for x in rad)
            elif t == 'resources':
                continue
            else:
                raise Exception(f"load_dict: object '{k}' of type '{t}' is outside the volprim hot path "
                                f"(supported: {ELLIPSOID_TYPES + SENSOR_TYPES + EMITTER_TYPES + ('batch',)})")
        return sc
    if kind in SENSOR_TYPES:
        return PerspectiveSensor(d)
    if kind == 'batch':
        return BatchSensor(d)
    if kind in ELLIPSOID_TYPES:
        return EllipsoidsShape(d, device=device)
    if kind in ('volprim_rf', 'volprim_tomography'):
        return create_integrator(d)
    raise Exception(f"load_dict: unsupported object type '{kind}'")


class SceneParameters(OrderedDict):
    """mi.traverse(scene): '<shape id>.data', '<shape id>.<attribute>' -> CUDA tensors; `update()` pushes
    assigned values back and triggers the acceleration-structure rebuild (refine_3dg_dataset.py:155-159)."""

    def __init__(self, scene: Scene):
        super().__init__()
        self._scene = scene
        for sh in scene.shapes():
            super().__setitem__(f'{sh.id}.data', sh.data)
            for k, v in sh.attributes.items():
                super().__setitem__(f'{sh.id}.{k}', v)

    def update(self, values=None):  # noqa: A003 - mirrors mi.SceneParameters.update
        if values:
            for k, v in dict(values).items():
                self[k] = v
        for sh in self._scene.shapes():
            for key, val in self.items():
                sid, _, name = key.partition('.')
                if sid != sh.id:
                    continue
                val = val if isinstance(val, torch.Tensor) else torch.as_tensor(val)
                flat = val.to(device=sh.device, dtype=torch.float32).reshape(-1)
                if name == 'data':
                    sh.data = flat
                else:
                    sh.attributes[name] = flat
            sh.parameters_changed()
        return []


def traverse(scene: Scene) -> SceneParameters:
    return SceneParameters(scene)


# ---------------------------------------------------------------------------------------------------
# render
# ---------------------------------------------------------------------------------------------------
REUSE_PRIMAL_RECORDS = True      # False: always re-trace in the backward pass, as RBIntegrator.render_backward does
RECORD_BUDGET_BYTES = 32 << 30   # hit records kept between the primal and the backward pass, summed over the views of a
                                 # render() call; views beyond the budget are re-traced in the backward pass


def _sample_positions(W, H, spp, seed, jitter, device):
    if not jitter:
        return None
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return torch.rand((W * H * spp, 2), generator=g, device=device, dtype=torch.float32)


def _rfilter_id(name):
    if name not in _cabi.RFILTERS:
        raise Exception(f"unsupported reconstruction filter '{name}' (supported: {sorted(_cabi.RFILTERS)})")
    return _cabi.RFILTERS[name]


class _ViewAux:
    """What the backward pass needs from the primal pass of one view."""
    __slots__ = ('rays', 'W', 'H', 'spp', 'jit', 'rfilter', 'state', 'accum', 'record')


class _RenderOp(torch.autograd.Function):
    """mi.render as a differentiable op: forward = primal render, backward = RBIntegrator.render_backward
    (SURVEY.md section 3.2).  `names` lists the parameter tensors passed in `tensors`, in the same order."""

    @staticmethod
    def forward(ctx, scene, sensor, integrator, seed, seed_grad, spp, spp_grad, jitter, srgb, rows, names, *tensors):
        # When the gradient pass would draw exactly the primal's samples (pixel centres, or the same seed and spp), the
        # primal records its hit lists (compressed rows, 4 B per hit) and the backward pass replays them instead of
        # walking the BVH a second time.  Russian roulette runs in the primal only (reference quirk), so its lists are
        # not what the adjoint visits: no replay then.
        spp_g = spp_grad or spp
        same_samples = spp_g == spp and (not jitter or seed_grad == seed)
        keep = same_samples and REUSE_PRIMAL_RECORDS and not getattr(integrator, 'use_rr', False)
        img, aux = _render_primal(scene, sensor, integrator, seed, spp, jitter, record=keep, srgb=srgb, rows=rows)
        ctx.args = (scene, sensor, integrator, seed_grad, spp_g, jitter, srgb, rows)
        ctx.aux = aux if same_samples else None
        ctx.names = names
        return img

    @staticmethod
    def backward(ctx, grad_img):
        scene, sensor, integrator, seed_grad, spp_grad, jitter, srgb, rows = ctx.args
        shape = scene.ellipsoids()
        shape.zero_grad()
        _render_adjoint(scene, sensor, integrator, seed_grad, spp_grad, jitter, grad_img.contiguous(), srgb, aux=ctx.aux,
                        rows=rows)
        ctx.aux = None
        out = []
        for name in ctx.names:                      # gradients matched to the inputs BY NAME
            g = shape.grad.get(name)
            out.append(None if g is None else g.clone())
        return (None,) * 11 + tuple(out)


def _differentiable_names(shape, integrator):
    names = ['data', integrator.attribute_name]
    if integrator.integrator_id == _cabi.INTEGRATOR_RF and 'sh_coeffs' in shape.attributes:
        names.append('sh_coeffs')
    return names


def _views(sensor):
    if isinstance(sensor, BatchSensor):
        return sensor.sensors, sensor.rfilter
    return [sensor], sensor.rfilter


class _srgb_override:
    def __init__(self, integrator, srgb):
        self.integrator, self.srgb = integrator, srgb

    def __enter__(self):
        self.saved = getattr(self.integrator, 'srgb_primitives', None)
        if self.srgb is not None and self.saved is not None:
            self.integrator.srgb_primitives = self.srgb

    def __exit__(self, *exc):
        if self.saved is not None:
            self.integrator.srgb_primitives = self.saved


def _render_primal(scene, sensor, integrator, seed, spp, jitter, record, srgb=None, rows=None):
    with _srgb_override(integrator, srgb):
        return _render_primal_impl(scene, sensor, integrator, seed, spp, jitter, record, rows)


def _render_primal_impl(scene, sensor, integrator, seed, spp, jitter, record, rows):
    """Per view: rays are generated INSIDE the trace kernel from the sensor (no ray buffers), the film kernels
    reconstruct the image (box / tent / gaussian at their true radii) straight into the view's column block."""
    from .accel import RaySource
    shape = scene.ellipsoids()
    shape.bind(attr_name=integrator.attribute_name, with_sh=integrator.integrator_id == _cabi.INTEGRATOR_RF)
    acc = shape.accel()
    views, rfilter = _views(sensor)
    rf = _rfilter_id(rfilter)
    params = integrator._vp_params(scene, None)
    if rows is not None and len(views) != 1:
        raise Exception("render(rows=...) renders a band of ONE sensor")
    image, aux, x0, budget = None, [], 0, RECORD_BUDGET_BYTES
    Wt = sum(v.width for v in views)
    for vi, s in enumerate(views):
        W, H = s.width, s.height
        band = None
        if rows is not None:
            y0, y1 = int(rows[0]), int(rows[1])
            if not (0 <= y0 < y1 <= H):
                raise Exception(f"render(rows={rows}): band outside the film (height {H})")
            if jitter and rfilter != 'box':
                raise Exception("render(rows=...): a filter wider than a pixel would be cut at the band seams; "
                                "use the box filter or jitter=False")
            band, H = (y0, y1 - y0), y1 - y0
        jit = _sample_positions(W, H, spp, seed * 7919 + vi, jitter, shape.device)
        rays = RaySource(camera=s.vp_camera(), spp=spp, jitter=jit, rows=band)
        want = bool(record)
        if want:   # budget over the views of this call
            need = acc.record_bytes(rays.n_rays, integrator._cap())
            want = need <= budget
            budget -= need if want else 0
        res = acc.render_forward(params, rays, record=want, id_cap=integrator._cap(), want_beta=False, want_nhits=False)
        integrator.last = res
        direct = spp == 1 and (jit is None or rfilter == 'box')   # every sample is exactly its own pixel
        accum = None
        if direct:
            view_img = res.rgb.reshape(H, W, 3)
            if len(views) == 1:
                image = view_img
            else:
                if image is None:
                    image = torch.empty((H, Wt, 3), dtype=torch.float32, device=shape.device)
                image[:, x0:x0 + W] = view_img
        else:
            if image is None:
                image = torch.empty((H, Wt, 3), dtype=torch.float32, device=shape.device)
            accum = torch.zeros(W * H * 4, dtype=torch.float32, device=shape.device)
            acc.film_splat(W, H, spp, rf, jit, res.rgb, accum)
            acc.film_develop(W, H, accum, image[:, x0:x0 + W])
        a = _ViewAux()
        a.rays, a.W, a.H, a.spp, a.jit, a.rfilter, a.state, a.accum, a.record = rays, W, H, spp, jit, rf, res.rgb, accum, res.record
        aux.append(a)
        x0 += W
    return image, aux


def _render_adjoint(scene, sensor, integrator, seed, spp, jitter, grad_img, srgb=None, aux=None, rows=None):
    """RBIntegrator.render_backward: primal with the gradient seed, then sample(Backward) with state_in = the
    primal's state_out (volprim_rf.py:192) -- which, with srgb_primitives, is the LINEAR radiance fed into an
    sRGB-space recursion (reference quirk Q3, reproduced when srgb is left at the plugin's own value)."""
    if aux is None:
        _, aux = _render_primal(scene, sensor, integrator, seed, spp, jitter, record=REUSE_PRIMAL_RECORDS and not
                                getattr(integrator, 'use_rr', False), srgb=srgb, rows=rows)
    shape = scene.ellipsoids()
    acc = shape.accel()
    x0 = 0
    with _srgb_override(integrator, srgb):
        params = integrator._vp_params(scene, None)
        out = shape.grad_buffers(integrator.attribute_name)
        for a in aux:
            g = grad_img[:, x0:x0 + a.W, :]
            x0 += a.W
            if a.accum is None:
                dL = g.reshape(-1, 3) if g.is_contiguous() else g.contiguous().reshape(-1, 3)
            else:
                dL = acc.film_adjoint(a.W, a.H, a.spp, a.rfilter, a.jit, a.accum, g)
            rec = a.record
            if rec is not None:
                acc.note_record(rec)
            if rec is not None and rec.usable():
                acc.render_adjoint(params, a.rays, dL, a.state, rec, out=out)
            else:
                # no (usable) record: re-trace like the primal, on explicit rays of the same sensor samples
                o, d, maxt = _explicit_rays(acc, a)
                p2 = integrator._vp_params(scene, (a.W * a.spp, a.H) if (a.W * a.spp) % 8 == 0 and a.H % 4 == 0 else None)
                acc.trace_adjoint(p2, o, d, maxt, dL, a.state, out=out)


def _explicit_rays(acc, a):
    """Rays of a view (band / jittered) as explicit tensors, for the re-trace fallback of the backward pass."""
    cam = a.rays.camera
    if a.rays.rows:
        y0, n = a.rays.rows
        full_jit = None
        if a.jit is not None:
            full_jit = torch.full((cam.width * cam.height * a.spp, 2), 0.5, dtype=torch.float32, device=acc.device)
            full_jit[y0 * cam.width * a.spp:(y0 + n) * cam.width * a.spp] = a.jit
        o, d, maxt = acc.raygen_perspective(cam, a.spp, full_jit)
        sl = slice(y0 * cam.width * a.spp, (y0 + n) * cam.width * a.spp)
        return o[sl].contiguous(), d[sl].contiguous(), maxt[sl].contiguous()
    return acc.raygen_perspective(cam, a.spp, a.jit)


def render(scene: Scene, params=None, sensor=0, integrator=None, seed=0, seed_grad=0, spp=0, spp_grad=0,
           jitter=True, adjoint_mode='reference_exact', rows=None):
    """mi.render(scene, params, sensor, integrator, seed, seed_grad, spp, spp_grad): returns an [H, W, 3]
    CUDA tensor.  When `params` holds tensors with requires_grad, the result carries a grad_fn whose backward
    is the PRB adjoint (`loss.backward()` plays the role of `dr.backward(loss)`).

    Extensions: `jitter=False` samples pixel centres (deterministic; the reference always jitters with
    Mitsuba's PCG32 stream, which is third-party); `adjoint_mode='corrected'` differentiates through
    srgb_to_linear instead of reproducing the reference's colour-space inconsistency (quirk Q3);
    `rows=(y0, y1)` renders only that band of film rows (image-tile sharding over several GPUs,
    volprim_balance_b200.parallel.render_tiles)."""
    integrator = integrator or scene.integrator
    if not isinstance(integrator, VolprimIntegratorBase):
        raise Exception("render: the scene has no volprim integrator")
    if isinstance(sensor, int):
        sensor = scene.sensors()[sensor]
    spp = int(spp) if spp else 1
    shape = scene.ellipsoids()
    tensors, names = [], []
    if params is not None:
        for name in _differentiable_names(shape, integrator):
            t = params.get(f'{shape.id}.{name}')
            if t is not None:
                tensors.append(t)
                names.append(name)
    need_grad = any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors) and torch.is_grad_enabled()
    if not need_grad:
        return _render_primal(scene, sensor, integrator, seed, spp, jitter, record=False, rows=rows)[0]
    if adjoint_mode not in ('reference_exact', 'corrected'):
        raise Exception("adjoint_mode must be 'reference_exact' or 'corrected'")
    names = tuple(names)
    if adjoint_mode == 'corrected' and getattr(integrator, 'srgb_primitives', False):
        # render in sRGB space and convert with autograd-visible torch ops: the derivative of srgb_to_linear is
        # applied and state_in is the sRGB-space radiance, i.e. the true gradient of the rendered image.
        img_s = _RenderOp.apply(scene, sensor, integrator, seed, seed_grad, spp, spp_grad, jitter, False, rows, names, *tensors)
        return torch.where(img_s <= 0.04045, img_s / 12.92, ((img_s.clamp_min(0.04045) + 0.055) / 1.055) ** 2.4)
    return _RenderOp.apply(scene, sensor, integrator, seed, seed_grad, spp, spp_grad, jitter, None, rows, names, *tensors)


def render_to_host(scene: Scene, sensors=None, out=None, integrator=None, seed=0, spp=0, jitter=True, on_image=None):
    """Render a sequence of sensors and deliver every image in PINNED HOST memory (the loop every caller of the
    reference writes around `mi.render` + `mi.Bitmap(img)`, e.g. examples/render_3dg_asset.py:76-92).

    The device->host copy of view i runs on a copy stream while view i+1 is traced, so the copy is hidden behind
    the kernels instead of serialising with them.  `sensors`: indices or sensor objects (default: all sensors of the
    scene).  `out`: optional list of pinned [H, W, 3] float32 tensors used round-robin (at least two; default: one
    fresh pinned image per view).  `on_image(i, host_image)` is called once image i is complete in host memory
    (while later views are in flight).  Returns the list of host images, all complete on return.  Primal only."""
    integrator = integrator or scene.integrator
    if not isinstance(integrator, VolprimIntegratorBase):
        raise Exception("render_to_host: the scene has no volprim integrator")
    all_sensors = scene.sensors()
    sensors = list(range(len(all_sensors))) if sensors is None else list(sensors)
    sensors = [all_sensors[x] if isinstance(x, int) else x for x in sensors]
    if out is not None and len(out) < 2 and len(sensors) > 1:
        raise Exception("render_to_host: `out` needs at least two pinned images to overlap copies with rendering")
    spp = int(spp) if spp else 1
    main = torch.cuda.current_stream()
    copier = getattr(scene, '_copy_stream', None)
    if copier is None:
        copier = scene._copy_stream = torch.cuda.Stream()
    results, done = [], []
    with torch.no_grad():
        for i, sensor in enumerate(sensors):
            img = _render_primal(scene, sensor, integrator, seed + i, spp, jitter, record=False)[0]
            host = out[i % len(out)] if out is not None else torch.empty(img.shape, dtype=img.dtype).pin_memory()
            if out is not None and i >= len(out):
                done[i - len(out)].synchronize()       # the buffer's previous image is complete (and was handed to on_image)
            ready = torch.cuda.Event()
            ready.record(main)
            copier.wait_event(ready)
            with torch.cuda.stream(copier):
                host.copy_(img, non_blocking=True)
                img.record_stream(copier)
                ev = torch.cuda.Event()
                ev.record(copier)
            results.append(host)
            done.append(ev)
            if on_image is not None and i > 0:
                done[i - 1].synchronize()
                on_image(i - 1, results[i - 1])
    if done:
        done[-1].synchronize()
        if on_image is not None:
            on_image(len(done) - 1, results[-1])
    return results
