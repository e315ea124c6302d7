import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from volprim_balance_b200 import synthetic, _cabi
from volprim_balance_b200.accel import EllipsoidAccel
n = int(sys.argv[1]); sigma0 = float(sys.argv[2]); centers = sys.argv[3]; kernel = int(sys.argv[4]); W, H = int(sys.argv[5]), int(sys.argv[6])
cloud = synthetic.make_cloud(n, sigma0, seed=2, sh_degree=3, centers=centers, mu_opacity=-1.0)
acc = EllipsoidAccel()
acc.set_primitives(torch.from_numpy(cloud.data), torch.from_numpy(cloud.opacities), torch.from_numpy(cloud.sh_coeffs), 3.0)
acc.build(); torch.cuda.synchronize(); print("build ok", flush=True)
p = _cabi.vp_params(); p.integrator = 0; p.kernel = kernel; p.max_depth = 128
p.srgb_primitives = 1; p.t_cutoff = 0.01; p.eps_advance = 1e-4; p.image_width = W; p.image_height = H
o, d, mt = synthetic.camera_rays(synthetic.ring_camera(0, 8, W, H))
r = acc.trace_forward(p, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda())
torch.cuda.synchronize(); print("tile trace ok", acc.stats(), flush=True)
p.image_width = 0; p.image_height = 0
r = acc.trace_forward(p, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda())
torch.cuda.synchronize(); print("per-ray trace ok", acc.stats(), flush=True)
