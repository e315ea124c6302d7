import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from volprim_balance_b200 import synthetic, _cabi
from volprim_balance_b200.accel import EllipsoidAccel
n = int(sys.argv[1]); W, H = 240, 136
cloud = synthetic.make_cloud(n, 0.0014, seed=2, sh_degree=int(sys.argv[2]) if len(sys.argv) > 2 else 3)
acc = EllipsoidAccel()
acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
acc.build(); torch.cuda.synchronize(); print("build ok", flush=True)
nodes, perm = acc.debug_bvh()
nodes = nodes.cpu().numpy(); perm = perm.cpu().numpy()
links = nodes[:, 12:14].copy().view(np.int32).reshape(-1)
print("perm is permutation:", np.array_equal(np.sort(perm), np.arange(n)))
leaf = ~links[links < 0]; internal = links[links >= 0]
print("leaf range ok", leaf.min(), leaf.max(), "unique", len(np.unique(leaf)) == n, "internal range", internal.min(), internal.max(), "unique", len(np.unique(internal)) == n - 2)
print("boxes finite", np.isfinite(nodes[:, :12]).all(), "lo<=hi", (nodes[:, 0:3] <= nodes[:, 3:6]).all(), (nodes[:, 6:9] <= nodes[:, 9:12]).all())
p = _cabi.vp_params(); p.integrator = 0; p.kernel = 0; p.max_depth = 128
p.srgb_primitives = 1; p.t_cutoff = 0.01; p.eps_advance = 1e-4
o, d, mt = synthetic.camera_rays(synthetic.ring_camera(0, 8, W, H))
to, td, tm = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda()
r = acc.trace_forward(p, to, td, tm); torch.cuda.synchronize(); print("per-ray ok", acc.stats(), flush=True)
p.image_width = W; p.image_height = H
r = acc.trace_forward(p, to, td, tm); torch.cuda.synchronize(); print("tile ok", acc.stats(), flush=True)
