"""Torch-backed stand-in for the slice of the drjit / mitsuba Python API that the reference's hot-path source
uses, so that /root/reference/volprim/integrators/{common,volprim_rf,volprim_tomography}.py can be EXECUTED
UNMODIFIED in the authoring container (Mitsuba 3 / Dr.Jit are not installable here).

Used only by make_golden.py to generate the fixtures in this directory.  What it pins: every formula that lives
in the reference's own files (kernel eval, density integrals, ray/ellipsoid quadratic set-up, the two sample()
loops, the PRB adjoint via torch autograd through the reference code).  What it does NOT pin (restated here from
the published definitions, like in the oracle): scene.ray_intersect semantics, dr.sh_eval, dr.quat_to_matrix,
mi.math.srgb_to_linear, mi.math.improved_solve_quadratic.
"""
from __future__ import annotations

import contextlib
import enum
import importlib.util
import math
import sys
import types

import torch

DTYPE = torch.float64


# ------------------------------------------------------------------------------------------------
# array wrappers
# ------------------------------------------------------------------------------------------------
def _raw(x):
    if isinstance(x, Arr):
        return x.t
    if isinstance(x, (bool, int, float)):
        return x
    if isinstance(x, torch.Tensor):
        return x
    raise TypeError(type(x))


class Arr:
    """Dr.Jit-like 1-D array (Float / UInt32 / Bool) over a torch tensor, with masked get / set."""
    __array_priority__ = 100

    def __init__(self, v=0.0, dtype=None):
        if isinstance(v, Arr):
            v = v.t
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(v, dtype=dtype) if dtype is not None else torch.as_tensor(v)
        if dtype is not None and v.dtype != dtype:
            v = v.to(dtype)
        self.t = v

    # arithmetic -----------------------------------------------------------------------------
    def _bin(self, o, f, rev=False):
        if isinstance(o, (Vec, Mat3)):
            return NotImplemented
        a, b = self.t, _raw(o)
        return Arr(f(b, a) if rev else f(a, b))

    def __add__(self, o): return self._bin(o, torch.add if isinstance(_raw(o), torch.Tensor) else lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return Arr(-self.t)
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    def __eq__(self, o): return self._bin(o, lambda a, b: a == b)
    def __ne__(self, o): return self._bin(o, lambda a, b: a != b)
    __hash__ = object.__hash__

    def __and__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        if isinstance(o, bool):
            return Arr(self.t) if o else Arr(torch.zeros_like(self.t))
        return Arr(self.t & _raw(o))
    __rand__ = __and__

    def __or__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        if isinstance(o, bool):
            return Arr(torch.ones_like(self.t)) if o else Arr(self.t)
        return Arr(self.t | _raw(o))
    __ror__ = __or__

    def __invert__(self): return Arr(~self.t)

    # masked access: x[mask] returns the value, x[mask] = v blends
    def __getitem__(self, mask): return Arr(self.t)

    def __setitem__(self, mask, v):
        m = _raw(mask)
        v = _raw(v)
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(v, dtype=self.t.dtype)
        self.t = torch.where(m, v.to(self.t.dtype), self.t)

    def __bool__(self): return bool(self.t.any()) if self.t.dtype == torch.bool else bool(self.t)
    def __int__(self): return int(self.t)
    def __float__(self): return float(self.t)
    def __repr__(self): return f"Arr({self.t})"


def Float(v=0.0): return Arr(v, DTYPE) if not isinstance(v, Arr) else Arr(v.t.to(DTYPE))
def UInt32(v=0): return Arr(v, torch.int64)
def Bool(v=False): return Arr(v, torch.bool) if not isinstance(v, Arr) else Arr(v.t.clone())


class Vec:
    """Point3f / Vector3f / Color3f / Spectrum / Quaternion4f: a fixed-length tuple of Arr."""

    def __init__(self, *a, n=3):
        if len(a) == 1 and isinstance(a[0], Vec):
            self.c = [Arr(x.t) for x in a[0].c]
        elif len(a) == 1 and isinstance(a[0], (list, tuple)):
            self.c = [x if isinstance(x, Arr) else Float(x) for x in a[0]]
        elif len(a) == 1:
            self.c = [Float(a[0]) for _ in range(n)]
        elif len(a) == 0:
            self.c = [Float(0.0) for _ in range(n)]
        else:
            self.c = [x if isinstance(x, Arr) else Float(x) for x in a]
        self.c = [Arr(x.t) for x in self.c]

    x = property(lambda s: s.c[0])
    y = property(lambda s: s.c[1])
    z = property(lambda s: s.c[2])
    w = property(lambda s: s.c[3])

    def _bin(self, o, f):
        if isinstance(o, Vec):
            return Vec([f(a, b) for a, b in zip(self.c, o.c)])
        return Vec([f(a, o) for a in self.c])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    __radd__ = __add__
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    __rmul__ = __mul__
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __neg__(self): return Vec([-a for a in self.c])
    def __ne__(self, o): return self._bin(o, lambda a, b: a != b)
    def __eq__(self, o): return self._bin(o, lambda a, b: a == b)
    def __and__(self, o): return self._bin(o, lambda a, b: a & b)
    __rand__ = __and__
    def __or__(self, o): return self._bin(o, lambda a, b: a | b)
    __ror__ = __or__
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    __hash__ = object.__hash__
    def __invert__(self): return Vec([~a for a in self.c])
    def __iter__(self): return iter(self.c)
    def __len__(self): return len(self.c)

    def __getitem__(self, i):
        if isinstance(i, int):
            return self.c[i]
        return Vec(self)  # masked read

    def __setitem__(self, i, v):
        if isinstance(i, int):
            self.c[i] = v if isinstance(v, Arr) else Float(v)
            return
        for k in range(len(self.c)):
            m = i.c[k] if isinstance(i, Vec) else i
            val = v.c[k] if isinstance(v, Vec) else v
            self.c[k][m] = val

    def __repr__(self): return f"Vec({[a.t for a in self.c]})"


class Mat3:
    def __init__(self, rows=None):
        self.m = rows if rows is not None else [[Float(0.0)] * 3 for _ in range(3)]

    @property
    def T(self):
        return Mat3([[self.m[j][i] for j in range(3)] for i in range(3)])

    def __mul__(self, v):
        assert isinstance(v, Vec)
        return Vec([self.m[i][0] * v.c[0] + self.m[i][1] * v.c[1] + self.m[i][2] * v.c[2] for i in range(3)])


# ------------------------------------------------------------------------------------------------
# drjit
# ------------------------------------------------------------------------------------------------
def _map(f, x):
    if isinstance(x, Vec):
        return Vec([_map(f, a) for a in x.c])
    if isinstance(x, Arr):
        return Arr(f(x.t))
    return f(torch.as_tensor(x, dtype=DTYPE)).item() if not isinstance(x, torch.Tensor) else f(x)


def _map2(f, a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        n = len(a) if isinstance(a, Vec) else len(b)
        return Vec([_map2(f, a.c[i] if isinstance(a, Vec) else a, b.c[i] if isinstance(b, Vec) else b) for i in range(n)])
    ta = _raw(a) if isinstance(a, Arr) else torch.as_tensor(a, dtype=DTYPE)
    tb = _raw(b) if isinstance(b, Arr) else torch.as_tensor(b, dtype=DTYPE)
    return Arr(f(ta, tb))


class ADMode(enum.Enum):
    Primal = 0
    Forward = 1
    Backward = 2


def _make_drjit():
    dr = types.ModuleType('drjit')
    dr.pi = math.pi
    dr.ADMode = ADMode
    dr.exp = lambda x: _map(torch.exp, x)
    dr.log = lambda x: _map(torch.log, x)
    dr.sqrt = lambda x: math.sqrt(x) if isinstance(x, (int, float)) else _map(torch.sqrt, x)
    dr.abs = lambda x: _map(torch.abs, x)
    dr.rcp = lambda x: 1.0 / x
    dr.isfinite = lambda x: _map(torch.isfinite, x)
    dr.maximum = lambda a, b: _map2(torch.maximum, a, b)
    dr.minimum = lambda a, b: _map2(torch.minimum, a, b)
    dr.select = lambda c, a, b: (Vec([dr.select(c.c[i] if isinstance(c, Vec) else c, a.c[i] if isinstance(a, Vec) else a,
                                                b.c[i] if isinstance(b, Vec) else b)
                                      for i in range(len(a) if isinstance(a, Vec) else len(b))])
                                 if isinstance(a, Vec) or isinstance(b, Vec)
                                 else Arr(torch.where(_raw(c), torch.as_tensor(_raw(a), dtype=DTYPE),
                                                      torch.as_tensor(_raw(b), dtype=DTYPE))))
    dr.dot = lambda a, b: a.c[0] * b.c[0] + a.c[1] * b.c[1] + a.c[2] * b.c[2]
    dr.squared_norm = lambda a: dr.dot(a, a)
    dr.norm = lambda a: dr.sqrt(dr.dot(a, a))

    def _any(x):
        if isinstance(x, Vec):
            r = x.c[0]
            for a in x.c[1:]:
                r = r | a
            return r
        return x
    dr.any = _any

    def _max(x):
        assert isinstance(x, Vec)
        r = x.c[0]
        for a in x.c[1:]:
            r = dr.maximum(r, a)
        return r
    dr.max = _max

    def detach(x):
        if isinstance(x, Vec):
            return Vec([detach(a) for a in x.c])
        if isinstance(x, Arr):
            return Arr(x.t.detach())
        if isinstance(x, _Ray):
            return _Ray(detach(x.o), detach(x.d), detach(x.maxt))
        return x
    dr.detach = detach
    dr.syntax = lambda f=None, **kw: f if f is not None else (lambda g: g)
    dr.hint = lambda cond, **kw: bool(cond)

    @contextlib.contextmanager
    def resume_grad(when=True):
        if when:
            with torch.enable_grad():
                yield
        else:
            yield
    dr.resume_grad = resume_grad
    dr.suspend_grad = torch.no_grad

    def backward_from(x):
        tot = None
        for a in (x.c if isinstance(x, Vec) else [x]):
            s = a.t.sum()
            tot = s if tot is None else tot + s
        if tot.requires_grad:
            tot.backward()
    dr.backward_from = backward_from

    def dispatch(target, func, *args):
        return func(target, *args)
    dr.dispatch = dispatch

    def zeros(tp, n=1):
        return tp() if isinstance(tp, type) else tp(0)
    dr.zeros = zeros

    # ---- third-party numerics restated (PARITY UNPINNED) ----
    def quat_to_matrix(q, size=3):
        x, y, z, w = q.c
        return Mat3([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    dr.quat_to_matrix = quat_to_matrix

    def sh_eval(d, degree):
        x, y, z = d.c
        Y = [None] * (degree + 1) ** 2
        Y[0] = Float(torch.full_like(x.t, 0.28209479177387814))
        if degree >= 1:
            Y[2] = 0.48860251190291992 * z
            Y[3] = -0.48860251190291992 * x
            Y[1] = -0.48860251190291992 * y
        if degree >= 2:
            z2 = z * z
            Y[6] = 0.94617469575756008 * z2 - 0.31539156525251999
            tb = -1.0925484305920792 * z
            Y[7], Y[5] = tb * x, tb * y
            c1, s1 = x * x - y * y, x * y + y * x
            Y[8], Y[4] = 0.54627421529603959 * c1, 0.54627421529603959 * s1
        if degree >= 3:
            Y[12] = z * (1.8658816629505769 * z2 - 1.1195289977703462)
            tc = -2.2852289973223288 * z2 + 0.45704579946446572
            Y[13], Y[11] = tc * x, tc * y
            td = 1.4453057213202769 * z
            Y[14], Y[10] = td * c1, td * s1
            c2, s2 = x * c1 - y * s1, x * s1 + y * c1
            Y[15], Y[9] = -0.59004358992664352 * c2, -0.59004358992664352 * s2
        return Y
    dr.sh_eval = sh_eval
    dr.alloc_local = lambda *a, **k: None
    return dr


# ------------------------------------------------------------------------------------------------
# mitsuba
# ------------------------------------------------------------------------------------------------
class _Ray:
    def __init__(self, o, d=None, maxt=None):
        if isinstance(o, _Ray):
            o, d, maxt = Vec(o.o), Vec(o.d), Arr(o.maxt.t)
        self.o, self.d, self.maxt = o, d, maxt

    def __call__(self, t):
        return self.o + self.d * t


class _ShapeType(enum.Enum):
    Mesh = 1
    Ellipsoids = 2

    def __pos__(self):
        return self


class _SI:
    def __init__(self):
        self.t = None
        self.p = None
        self.shape = None
        self.prim_index = None
        self._valid = None

    def is_valid(self):
        return self._valid


class Properties(dict):
    pass


class RefShape:
    """One ellipsoids shape; attribute tensors are torch leaves so that autograd reaches them."""

    def __init__(self, data10, attrs: dict, extent=3.0):
        self.data = data10        # [N, 10]
        self.attrs = attrs        # name -> [N, k]
        self.extent = extent

    def shape_type(self):
        return _ShapeType.Ellipsoids

    def has_attribute(self, name):
        return name in self.attrs or name in ('ellipsoid', 'extent')

    def eval_attribute_x(self, name, si, active):
        # an invalid interaction carries a null shape pointer in Mitsuba: dispatch on it yields zeros
        ok = si.prim_index.t >= 0
        idx = si.prim_index.t.clamp_min(0)
        src = self.data if name == 'ellipsoid' else self.attrs[name]
        g = src[idx]                                   # [R, k]
        if isinstance(active, Arr):
            ok = ok & active.t
        g = torch.where(ok.unsqueeze(-1), g, torch.zeros_like(g))
        return _Rows(g.transpose(0, 1))

    def eval_attribute_1(self, name, si, active):
        if name == 'extent':
            e = torch.full((si.prim_index.t.shape[0],), self.extent, dtype=DTYPE)
            return Float(torch.where(si.prim_index.t >= 0, e, torch.zeros_like(e)))
        ok = si.prim_index.t >= 0
        idx = si.prim_index.t.clamp_min(0)
        g = self.attrs[name].reshape(-1)[idx]
        if isinstance(active, Arr):
            ok = ok & active.t
        g = torch.where(ok, g, torch.zeros_like(g))
        return Float(g)


class _Rows:
    """What eval_attribute_x returns: indexable by component, with .shape[0] = component count."""

    def __init__(self, t):
        self.t = t
        self.shape = t.shape

    def __getitem__(self, i):
        return Arr(self.t[i])


class RefScene:
    """scene.ray_intersect = closest FRONT-FACE entry (t > 0) among the analytic ellipsoids at extent * scale
    (decision 1 of DESIGN.md; third-party semantics, unpinned).  Records the hit sequence."""

    def __init__(self, shape: RefShape, env=(1.0, 1.0, 1.0)):
        self.shape = shape
        self.env = env
        self.hit_log = []

    def shapes(self):
        return [self.shape]

    def environment(self):
        env = self.env

        class _E:
            def eval(self, si, active):
                return Vec(float(env[0]), float(env[1]), float(env[2]))
        return _E()

    def ray_intersect(self, ray, coherent=None, ray_flags=None, active=True):
        with torch.no_grad():
            d10 = self.shape.data.detach()
            o = torch.stack([a.t for a in ray.o.c], -1)       # [R,3]
            d = torch.stack([a.t for a in ray.d.c], -1)
            c, s, q = d10[:, 0:3], d10[:, 3:6] * self.shape.extent, d10[:, 6:10]
            x, y, z, w = q.unbind(-1)
            Rm = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                              2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                              2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
            v = o[:, None, :] - c[None, :, :]                                  # [R,N,3]
            oo = torch.einsum('nji,rnj->rni', Rm, v) / s[None]
            dd = torch.einsum('nji,rj->rni', Rm, d) / s[None]
            a = (dd * dd).sum(-1)
            b = -(oo * dd).sum(-1)
            cc = (oo * oo).sum(-1) - 1
            l = oo + (b / a)[..., None] * dd
            discr = 1 - (l * l).sum(-1)
            sq = torch.sqrt((a * discr).clamp_min(0))
            qq = b + torch.where(b >= 0, sq, -sq)
            x0, x1 = cc / qq, qq / a
            tn = torch.minimum(x0, x1)
            ok = (discr >= 0) & torch.isfinite(x0) & torch.isfinite(x1) & (tn > 0) & (tn <= ray.maxt.t[:, None])
            tn = torch.where(ok, tn, torch.full_like(tn, float('inf')))
            t, idx = tn.min(dim=1)
            valid = torch.isfinite(t)
            if isinstance(active, Arr):
                valid = valid & active.t
        si = _SI()
        si._valid = Arr(valid)
        si.t = Float(torch.where(valid, t, torch.zeros_like(t)))
        si.prim_index = Arr(torch.where(valid, idx, torch.full_like(idx, -1)))
        si.shape = self.shape
        si.p = ray(si.t)
        self.hit_log.append(si.prim_index.t.clone())
        return si


class _Sampler:
    def next_1d(self):
        return Float(0.5)


def _make_mitsuba(dr):
    mi = types.ModuleType('mitsuba')
    mi.Float, mi.UInt32, mi.Bool = Float, UInt32, Bool
    mi.Point3f = mi.Vector3f = mi.Color3f = mi.Spectrum = lambda *a: Vec(*a, n=3)
    mi.Quaternion4f = lambda *a: Vec(*a, n=4)
    mi.Matrix3f = lambda *a: Mat3()
    mi.ShapePtr = lambda *a: None
    mi.Ray3f = _Ray
    mi.Scene = mi.Sampler = mi.SurfaceInteraction3f = _SI
    mi.ShapeType = _ShapeType
    mi.Properties = Properties

    class RayFlags(enum.IntFlag):
        All = 0xfff
        BackfaceCulling = 0x1000
    mi.RayFlags = RayFlags

    class ParamFlags:
        NonDifferentiable = 1
    mi.ParamFlags = ParamFlags
    mi.LogLevel = types.SimpleNamespace(Warn=1)
    mi.Log = lambda *a, **k: None
    mi.registered = {}
    mi.register_integrator = lambda name, f: mi.registered.__setitem__(name, f)
    mi.variant = lambda: 'torch_shim'

    m = types.ModuleType('mitsuba.math')

    def srgb_to_linear(x):
        return dr.select(x <= 0.04045, x / 12.92, _map(lambda t: ((t.clamp_min(0.0) + 0.055) / 1.055) ** 2.4, x))
    m.srgb_to_linear = srgb_to_linear

    def improved_solve_quadratic(a, b, c, discr):
        sq = dr.sqrt(dr.maximum(a * discr, 0.0))
        q = b + dr.select(b >= 0.0, sq, -sq)
        x0, x1 = c / q, q / a
        valid = (discr >= 0.0) & dr.isfinite(x0) & dr.isfinite(x1)
        return valid, dr.minimum(x0, x1), dr.maximum(x0, x1)
    m.improved_solve_quadratic = improved_solve_quadratic
    mi.math = m

    class RBIntegrator:
        def __init__(self, props=None):
            props = props if props is not None else Properties()
            self.hide_emitters = props.get('hide_emitters', False)
    ad = types.ModuleType('mitsuba.ad')
    adi = types.ModuleType('mitsuba.ad.integrators')
    adc = types.ModuleType('mitsuba.ad.integrators.common')
    adc.RBIntegrator = RBIntegrator
    adc.mis_weight = lambda *a: None
    mi.ad, ad.integrators, adi.common = ad, adi, adc
    return mi, {'mitsuba.math': m, 'mitsuba.ad': ad, 'mitsuba.ad.integrators': adi, 'mitsuba.ad.integrators.common': adc}


def load_reference(ref_root='/root/reference'):
    """Imports the reference's integrator modules over the stand-in; returns (dr, mi, common, rf, tomo)."""
    dr = _make_drjit()
    mi, subs = _make_mitsuba(dr)
    sys.modules['drjit'] = dr
    sys.modules['mitsuba'] = mi
    sys.modules.update(subs)
    pkg = types.ModuleType('volprim')
    pkg.__path__ = [f'{ref_root}/volprim']
    sub = types.ModuleType('volprim.integrators')
    sub.__path__ = [f'{ref_root}/volprim/integrators']
    sys.modules['volprim'], sys.modules['volprim.integrators'] = pkg, sub
    out = []
    for name in ('stack', 'common', 'volprim_rf', 'volprim_tomography'):
        spec = importlib.util.spec_from_file_location(f'volprim.integrators.{name}',
                                                      f'{ref_root}/volprim/integrators/{name}.py')
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        out.append(mod)
    return dr, mi, out[1], out[2], out[3]
