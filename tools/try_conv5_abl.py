"""GPU experiment: ablations of engine 5 (HPL_CONV5_DBG) -- timing only, results are wrong with any bit set."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import ops, plans
ops.WEIGHT_CACHE = "always"          # kernel-only timings: weight images are built once
from hplflownet_b200.batching import concat_lattices
from try_conv5 import timeit

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
nbr = concat_lattices([bench.cloud_tables(s) for s in range(B)])["blur_neighbors"][0].to(dev)
h = nbr.size(1)
plan = plans.build(nbr)
x = torch.randn(h, 64, device=dev)
wp = torch.nn.Parameter(torch.randn(15, 64, 64, device=dev) * 0.05)
w = ops.with_owner(wp.detach(), wp, "fwd")
amax = ops.absmax(x)
x16 = ops.h16b_split(x, 64, amax)
n = int(os.environ.get("ITERS", "20"))
print("dbg %s: %.4f ms" % (os.environ.get("HPL_CONV5_DBG", "0"), timeit(lambda: ops.conv5(x16, plan, 64, w, None, ops.ACT_NONE, amax), n)))
