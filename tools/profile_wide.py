"""ncu workload: the 580 -> 1024 up-path layer on engine 2, forward and weight gradient (random table)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hplflownet_b200 import ops
ops.WEIGHT_CACHE = "always"
dev = torch.device("cuda")
h, c, co = 31162, 580, 1024
torch.manual_seed(0)
x = torch.randn(h, ops.round4(c), device=dev)
nbr = torch.randint(-1, h, (15, h), device=dev, dtype=torch.int32)
wp = torch.nn.Parameter(torch.randn(15, c, co, device=dev) * 0.02)
w = ops.with_owner(wp.detach(), wp, "fwd")
amax = ops.absmax(x)
dz = torch.randn(h, co, device=dev)
dz_amax = ops.absmax(dz)
for _ in range(2):
    ops.blur_gemm(x, c, nbr, h, w, None, ops.ACT_LEAKY, precision=2, x_amax=amax)
    ops.blur_wgrad(x, c, nbr, h, dz, co, 15, want_db=False, precision=2, x_amax=amax, dz_amax=dz_amax)
torch.cuda.synchronize()
print("ok")
