import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from volprim_balance_b200 import synthetic
from tests.parity_utils import make_params, gpu_scene, oracle_scene, compare_forward
n = 200000
cloud = synthetic.make_cloud(n, synthetic.sigma0_for_hits(n, 50), seed=3)
W, H = 256, 128
o, d, mt = synthetic.camera_rays(synthetic.ring_camera(1, 8, W, H))
to, td, tm = (torch.from_numpy(x).cuda() for x in (o, d, mt))
acc = gpu_scene(cloud)
pt, op = make_params(0, 0, 128, image=(W, H))
pr, _ = make_params(0, 0, 128)
a = acc.trace_forward(pt, to, td, tm, record_cap=128)
b = acc.trace_forward(pr, to, td, tm, record_cap=128)
same = (a.hit_ids == b.hit_ids).all(0)
print("tile vs per-ray identical rays:", float(same.float().mean()), "n diff", int((~same).sum()))
ref = oracle_scene(cloud).forward(op, o, d, mt, cap=128, fragility=True)
ia, ib = a.hit_ids.t().cpu().numpy(), b.hit_ids.t().cpu().numpy()
sa, sb = (ia == ref.hit_ids).all(1), (ib == ref.hit_ids).all(1)
print("tile==oracle", sa.mean(), "perray==oracle", sb.mean())
bad = np.flatnonzero(~sa & sb)[:5]
for r in bad:
    k = np.flatnonzero(ia[r] != ref.hit_ids[r])[0]
    print("ray", r, "first diff at hit", k, "tile", ia[r][max(0,k-1):k+3], "oracle", ref.hit_ids[r][max(0,k-1):k+3], "t", ref.hit_t[r][max(0,k-1):k+3], "frag", ref.fragility[r])
try:
    print(compare_forward(a, ref, 128))
except AssertionError as e:
    print("ASSERT", str(e)[:300])
