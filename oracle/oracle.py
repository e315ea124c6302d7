"""ctypes front end of the CPU oracle (oracle/volprim_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py.  The product package never imports this module.
Parity status: the formulas owned by /root/reference are pinned by tests/golden (reference source
executed over a torch stand-in for drjit/mitsuba); the third-party halves are unpinned restatements
(see the header of volprim_oracle.c and DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

RF, TOMO = 0, 1
GAUSS, EPAN = 0, 1


class _Params(C.Structure):
    _fields_ = [
        ("integrator", C.c_int32),
        ("kernel", C.c_int32),
        ("max_depth", C.c_uint32),
        ("srgb_primitives", C.c_int32),
        ("hide_emitters", C.c_int32),
        ("brute_force", C.c_int32),
        ("diagnostics", C.c_int32),
        ("pad_", C.c_int32),
        ("t_cutoff", C.c_double),
        ("eps_advance", C.c_double),
        ("env", C.c_double * 3),
        ("use_rr", C.c_int32),
        ("rr_depth", C.c_uint32),
        ("rr_seed", C.c_uint32),
        ("rr_skip", C.c_uint32),
    ]


@dataclass
class Params:
    integrator: int = RF
    kernel: int = GAUSS
    max_depth: int = 64           # volprim_rf.py:26 ; -1 => unlimited
    srgb_primitives: bool = True  # volprim_rf.py:41
    hide_emitters: bool = False
    brute_force: bool = False
    diagnostics: bool = False
    t_cutoff: float = 0.01
    eps_advance: float = 1e-4
    env: tuple = (1.0, 1.0, 1.0)
    rr_depth: int = -1            # volprim_rf.py:31; Russian roulette is active iff rr_depth >= 0 and
    rr_seed: int = 0              # (rr_depth < max_depth or max_depth == -1)            volprim_rf.py:39
    rr_skip: int = 0

    @property
    def use_rr(self) -> bool:
        return self.rr_depth >= 0 and (self.rr_depth < self.max_depth or self.max_depth == -1)

    def to_c(self) -> _Params:
        md = 0xFFFFFFFF if self.max_depth == -1 else int(self.max_depth)
        rrd = 0xFFFFFFFF if self.rr_depth == -1 else int(self.rr_depth)
        return _Params(int(self.integrator), int(self.kernel), md, int(self.srgb_primitives),
                       int(self.hide_emitters), int(self.brute_force), int(self.diagnostics), 0,
                       float(self.t_cutoff), float(self.eps_advance), (C.c_double * 3)(*self.env),
                       int(self.use_rr), rrd, int(self.rr_seed) & 0xFFFFFFFF, int(self.rr_skip) & 0xFFFFFFFF)


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (gcc, -ffp-contract=off)."""
    need = force or not all(os.path.exists(os.path.join(_BUILD, f"liboracle_{p}.so")) for p in ("f32", "f64"))
    if not need:
        src = os.path.getmtime(os.path.join(_HERE, "volprim_oracle.c"))
        need = any(os.path.getmtime(os.path.join(_BUILD, f"liboracle_{p}.so")) < src for p in ("f32", "f64"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_LIBS: dict = {}


def _lib(precision: str):
    if precision not in _LIBS:
        build()
        lib = C.CDLL(os.path.join(_BUILD, f"liboracle_{precision}.so"))
        real = C.c_float if precision == "f32" else C.c_double
        rp = C.POINTER(real)
        lib.orc_scene_create.restype = C.c_void_p
        lib.orc_scene_create.argtypes = [C.c_int64, rp, rp, rp, C.c_int32, C.c_double, C.c_int32]
        lib.orc_scene_free.argtypes = [C.c_void_p]
        lib.orc_trace_forward.argtypes = [C.c_void_p, C.POINTER(_Params), C.c_int64, rp, rp, rp, rp, rp,
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                          C.c_int32, C.POINTER(C.c_double)]
        lib.orc_trace_adjoint.argtypes = [C.c_void_p, C.POINTER(_Params), C.c_int64, rp, rp, rp, rp, rp,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.orc_replay_forward.argtypes = [C.c_void_p, C.POINTER(_Params), C.c_int64, rp, rp, rp, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_uint32), C.c_int32, rp, rp, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.orc_set_abs_mode.argtypes = [C.c_int]
        lib.orc_quat_to_matrix.argtypes = [rp, rp]
        lib.orc_sh_eval.argtypes = [rp, C.c_int, rp]
        lib.orc_srgb_to_linear.restype = real
        lib.orc_srgb_to_linear.argtypes = [real]
        lib.orc_srgb_to_linear_deriv.restype = real
        lib.orc_srgb_to_linear_deriv.argtypes = [real]
        lib.orc_ray_ellipsoid.restype = C.c_int
        lib.orc_ray_ellipsoid.argtypes = [rp, rp, rp, real, rp, rp]
        lib.orc_kernel_eval.restype = real
        lib.orc_kernel_eval.argtypes = [C.c_int, rp, rp]
        lib.orc_density_integral.restype = real
        lib.orc_density_integral.argtypes = [C.c_int, rp, rp, rp, real]
        lib.orc_rf_transmission.restype = real
        lib.orc_rf_transmission.argtypes = [C.c_int, rp, rp, rp, real]
        lib.orc_pcg32_float_at.restype = C.c_float
        lib.orc_pcg32_float_at.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        lib.orc_pcg32_uint_at.restype = C.c_uint32
        lib.orc_pcg32_uint_at.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        _LIBS[precision] = (lib, real, np.float32 if precision == "f32" else np.float64)
    return _LIBS[precision]


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


@dataclass
class ForwardResult:
    rgb: np.ndarray
    beta: np.ndarray
    nhits: np.ndarray
    hit_ids: np.ndarray | None = None
    hit_t: np.ndarray | None = None
    fragility: np.ndarray | None = None


class Scene:
    """Primitive cloud in the reference layouts: data [N,10] (center, scale, quat i-j-k-r;
    common.py:47-74), attr [N] (opacities | sigma_t), sh [N,C] (volprim_rf.py:88-95)."""

    def __init__(self, data10, attr=None, sh=None, extent: float = 3.0, precision: str = "f32", bvh: bool = True):
        self.lib, self.real, self.np_real = _lib(precision)
        self.precision = precision
        d = np.ascontiguousarray(np.asarray(data10, dtype=self.np_real).reshape(-1, 10))
        self.n = d.shape[0]
        a = None if attr is None else np.ascontiguousarray(np.asarray(attr, dtype=self.np_real).reshape(-1))
        s = None
        self.sh_floats = 0
        if sh is not None:
            s = np.ascontiguousarray(np.asarray(sh, dtype=self.np_real).reshape(self.n, -1))
            self.sh_floats = s.shape[1]
        self._h = self.lib.orc_scene_create(self.n, _ptr(d, self.real), _ptr(a, self.real), _ptr(s, self.real),
                                            self.sh_floats, float(extent), int(bvh))
        self.extent = float(extent)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.orc_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    def _rays(self, o, d, maxt):
        o = np.ascontiguousarray(np.asarray(o, dtype=self.np_real).reshape(-1, 3))
        d = np.ascontiguousarray(np.asarray(d, dtype=self.np_real).reshape(-1, 3))
        m = None if maxt is None else np.ascontiguousarray(np.asarray(maxt, dtype=self.np_real).reshape(-1))
        return o, d, m

    def forward(self, params: Params, o, d, maxt=None, cap: int = 0, fragility: bool = False) -> ForwardResult:
        o, d, m = self._rays(o, d, maxt)
        R = o.shape[0]
        rgb = np.zeros((R, 3), self.np_real)
        beta = np.zeros(R, self.np_real)
        nh = np.zeros(R, np.uint32)
        ids = np.full((R, cap), -1, np.int32) if cap else None
        ht = np.full((R, cap), np.inf, np.float64) if cap else None
        fr = np.zeros((R, 4), np.float64) if fragility else None
        if fragility:
            params = Params(**{**params.__dict__, "diagnostics": True})
        p = params.to_c()
        self.lib.orc_trace_forward(self._h, C.byref(p), R, _ptr(o, self.real), _ptr(d, self.real), _ptr(m, self.real),
                                   _ptr(rgb, self.real), _ptr(beta, self.real), _ptr(nh, C.c_uint32),
                                   _ptr(ids, C.c_int32), _ptr(ht, C.c_double), cap, _ptr(fr, C.c_double))
        return ForwardResult(rgb, beta, nh, ids, ht, fr)

    def replay(self, params: Params, o, d, maxt, ids, counts):
        """Evaluate GIVEN hit lists (ids [R, cap], counts [R]) with the loop's arithmetic.  Returns a dict:
        rgb, beta, valid (every listed primitive was a legal hit), hit_beta [R, cap], cmax [R]."""
        o, d, m = self._rays(o, d, maxt)
        R = o.shape[0]
        ids = np.ascontiguousarray(np.asarray(ids, np.int32).reshape(R, -1))
        cap = ids.shape[1]
        counts = np.ascontiguousarray(np.asarray(counts, np.uint32).reshape(R))
        rgb = np.zeros((R, 3), self.np_real)
        beta = np.zeros(R, self.np_real)
        valid = np.zeros(R, np.int32)
        hb = np.ones((R, max(cap, 1)), np.float64)
        cm = np.zeros(R, np.float64)
        p = params.to_c()
        self.lib.orc_replay_forward(self._h, C.byref(p), R, _ptr(o, self.real), _ptr(d, self.real), _ptr(m, self.real),
                                    _ptr(ids, C.c_int32), _ptr(counts, C.c_uint32), cap, _ptr(rgb, self.real),
                                    _ptr(beta, self.real), _ptr(valid, C.c_int32), _ptr(hb, C.c_double), _ptr(cm, C.c_double))
        return {"rgb": rgb, "beta": beta, "valid": valid.astype(bool), "hit_beta": hb, "cmax": cm}

    def adjoint(self, params: Params, o, d, dL, state_in, maxt=None, abs_terms: bool = False):
        """Returns (g_data [N,10], g_attr [N], g_sh [N,C]) in float64.  abs_terms=True: the sum of the ABSOLUTE per-hit
        contributions instead (bounds the rounding error of an fp32 accumulation of the same terms)."""
        o, d, m = self._rays(o, d, maxt)
        R = o.shape[0]
        dL = np.ascontiguousarray(np.asarray(dL, dtype=self.np_real).reshape(R, 3))
        st = np.ascontiguousarray(np.asarray(state_in, dtype=self.np_real).reshape(R, 3))
        gd = np.zeros((self.n, 10), np.float64)
        ga = np.zeros(self.n, np.float64)
        gs = np.zeros((self.n, max(self.sh_floats, 1)), np.float64)
        p = params.to_c()
        self.lib.orc_set_abs_mode(int(abs_terms))
        try:
            self.lib.orc_trace_adjoint(self._h, C.byref(p), R, _ptr(o, self.real), _ptr(d, self.real), _ptr(m, self.real),
                                       _ptr(dL, self.real), _ptr(st, self.real), _ptr(gd, C.c_double),
                                       _ptr(ga, C.c_double), _ptr(gs, C.c_double))
        finally:
            self.lib.orc_set_abs_mode(0)
        return gd, ga, (gs if self.sh_floats else None)


# ---- scalar helpers for known-answer tests ------------------------------------------------------
def _vec(x, np_real, n):
    a = np.ascontiguousarray(np.asarray(x, dtype=np_real).reshape(n))
    return a


def quat_to_matrix(q, precision="f32"):
    lib, real, npr = _lib(precision)
    out = np.zeros(9, npr)
    lib.orc_quat_to_matrix(_ptr(_vec(q, npr, 4), real), _ptr(out, real))
    return out.reshape(3, 3)


def sh_eval(d, degree, precision="f32"):
    lib, real, npr = _lib(precision)
    out = np.zeros(16, npr)
    lib.orc_sh_eval(_ptr(_vec(d, npr, 3), real), int(degree), _ptr(out, real))
    return out[: (degree + 1) ** 2]


def srgb_to_linear(x, precision="f32"):
    lib, real, npr = _lib(precision)
    return np.array([lib.orc_srgb_to_linear(real(float(v))) for v in np.ravel(x)], npr).reshape(np.shape(x))


def srgb_to_linear_deriv(x, precision="f32"):
    lib, real, npr = _lib(precision)
    return np.array([lib.orc_srgb_to_linear_deriv(real(float(v))) for v in np.ravel(x)], npr).reshape(np.shape(x))


def ray_ellipsoid(o, d, rec10, extent=3.0, precision="f32"):
    lib, real, npr = _lib(precision)
    tn, tf = real(0), real(0)
    ok = lib.orc_ray_ellipsoid(_ptr(_vec(o, npr, 3), real), _ptr(_vec(d, npr, 3), real), _ptr(_vec(rec10, npr, 10), real),
                               real(extent), C.byref(tn), C.byref(tf))
    return bool(ok), tn.value, tf.value


def kernel_eval(kernel, p, rec10, precision="f32"):
    lib, real, npr = _lib(precision)
    return lib.orc_kernel_eval(int(kernel), _ptr(_vec(p, npr, 3), real), _ptr(_vec(rec10, npr, 10), real))


def density_integral(kernel, o, d, rec10, extent=3.0, precision="f32"):
    lib, real, npr = _lib(precision)
    return lib.orc_density_integral(int(kernel), _ptr(_vec(o, npr, 3), real), _ptr(_vec(d, npr, 3), real),
                                    _ptr(_vec(rec10, npr, 10), real), real(extent))


def rf_transmission(kernel, o, d, rec10, opacity, precision="f32"):
    lib, real, npr = _lib(precision)
    return lib.orc_rf_transmission(int(kernel), _ptr(_vec(o, npr, 3), real), _ptr(_vec(d, npr, 3), real),
                                   _ptr(_vec(rec10, npr, 10), real), real(opacity))


def pcg32_float_at(seed: int, idx: int, n: int) -> float:
    return float(_lib("f32")[0].orc_pcg32_float_at(seed & 0xFFFFFFFF, idx & 0xFFFFFFFF, n & 0xFFFFFFFF))


def pcg32_uint_at(initstate: int, initseq: int, n: int) -> int:
    return int(_lib("f32")[0].orc_pcg32_uint_at(initstate, initseq, n))


def num_threads() -> int:
    return _lib("f32")[0].orc_num_threads()


def set_num_threads(n: int) -> None:
    for p in ("f32", "f64"):
        _lib(p)[0].orc_set_num_threads(int(n))
