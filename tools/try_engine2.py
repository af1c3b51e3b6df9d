"""GPU experiment: engine 2 (register-staged 3xFP16 gather-GEMM) timings on the cfg2 layer and a wide up-path layer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hplflownet_b200 import ops
ops.WEIGHT_CACHE = "always"
from try_conv5 import timeit

dev = torch.device("cuda")
for name, h, c, co in (("cfg2 x32", 242429, 64, 64), ("bcn1_ 580->1024", 31162, 580, 1024), ("bcn3_ 128->256", 14500, 324, 256)):
    torch.manual_seed(0)
    x = torch.randn(h, ops.round4(c), device=dev)
    nbr = torch.randint(-1, h, (15, h), device=dev, dtype=torch.int32)
    wp = torch.nn.Parameter(torch.randn(15, c, co, device=dev) * 0.02)
    w = ops.with_owner(wp.detach(), wp, "fwd")
    amax = ops.absmax(x)
    dz = torch.randn(h, co, device=dev)
    dz_amax = ops.absmax(dz)
    t_f = timeit(lambda: ops.blur_gemm(x, c, nbr, h, w, None, ops.ACT_LEAKY, precision=2, x_amax=amax), 10)
    t_w = timeit(lambda: ops.blur_wgrad(x, c, nbr, h, dz, co, 15, want_db=False, precision=2, x_amax=amax, dz_amax=dz_amax), 10)
    fl = 2.0 * 15 * c * co * h
    print("%-18s fwd %.4f ms (%.0f TFLOP/s)   wgrad %.4f ms (%.0f TFLOP/s)" % (name, t_f, fl / t_f / 1e9, t_w, fl / t_w / 1e9))
