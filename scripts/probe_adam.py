import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volprim_balance_b200 import optimizers
n = 59_000_000
opt = optimizers.BoundedAdam(lr=1e-3); opt["x"] = torch.rand(n, device="cuda"); opt.set_bounds("x", lower=1e-6, upper=1.0)
g = torch.randn(n, device="cuda")
import ctypes as C
from volprim_balance_b200 import _cabi
lib = _cabi.load_library()
p, m, v = opt["x"].detach().clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
ptr = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): lib.vp_bounded_adam_step(n, ptr(p), ptr(g), ptr(m), ptr(v), 1e-3, 0.9, 0.999, 1e-8, 1, 1e-6, 1, 1.0, st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib.vp_bounded_adam_step(n, ptr(p), ptr(g), ptr(m), ptr(v), 1e-3, 0.9, 0.999, 1e-8, 1, 1e-6, 1, 1.0, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"k_bounded_adam: n={n} {ms:.3f} ms/step  {28 * n / ms / 1e6:.1f} GB/s ({28 * n / ms / 1e6 / 6544.3 * 100:.1f} % of measured HBM peak)")
