"""GPU experiment: ablations of engine 5's weight-gradient kernel (HPL_WGRAD5_DBG) -- timing only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import ops, plans
from hplflownet_b200.batching import concat_lattices
from try_conv5 import timeit

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
nbr = concat_lattices([bench.cloud_tables(s) for s in range(B)])["blur_neighbors"][0].to(dev)
h = nbr.size(1)
plan = plans.build(nbr)
x = torch.randn(h, 64, device=dev)
dz = torch.randn(h, 64, device=dev)
ax, az = ops.absmax(x), ops.absmax(dz)
x16, dz16 = ops.h16b_split(x, 64, ax), ops.h16b_split(dz, 64, az)
out = torch.zeros(15, 64, 64, device=dev)
print("dbg %s: %.4f ms" % (os.environ.get("HPL_WGRAD5_DBG", "0"), timeit(lambda: ops.wgrad5(x16, dz16, plan, 64, 64, ax, az, out=out), 20)))
