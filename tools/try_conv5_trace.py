"""GPU experiment: clock64 trace of engine 5's stage pipeline (CTA 0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import ops, plans
ops.WEIGHT_CACHE = "always"          # kernel-only timings: weight images are built once
from hplflownet_b200.batching import concat_lattices

dev = torch.device("cuda")
nbr = concat_lattices([bench.cloud_tables(s) for s in range(32)])["blur_neighbors"][0].to(dev)
h = nbr.size(1)
plan = plans.build(nbr)
x = torch.randn(h, 64, device=dev)
wp = torch.nn.Parameter(torch.randn(15, 64, 64, device=dev) * 0.05)
w = ops.with_owner(wp.detach(), wp, "fwd")
amax = ops.absmax(x)
x16 = ops.h16b_split(x, 64, amax)
for _ in range(3):
    ops.conv5(x16, plan, 64, w, None, ops.ACT_NONE, amax)
tr = torch.zeros(256 * 8 + 64 * 4, dtype=torch.int64, device=dev)
os.environ["HPL_CONV5_TRACE"] = str(tr.data_ptr())
ops.conv5(x16, plan, 64, w, None, ops.ACT_NONE, amax)
torch.cuda.synchronize()
t = tr[:2048].view(256, 8).cpu()
te = tr[2048:].view(64, 4).cpu()
t0 = int(t[0, 0])
print("stage: copy[start, empty-ok, copied, loads-issued, arrived]  mma[start-wait, full-ok, committed]   (clk since start)")
for i in range(0, 64):
    print(i, [int(v) - t0 for v in t[i]])
d = t[1:200, 4] - t[0:199, 4]
print("mean period between copy arrivals: %.0f clk" % d.float().mean().item())
print("copy: row reads %.0f  empty wait + W request + group barrier %.0f  tcgen05.st / STS %.0f  barrier+arrive %.0f" % tuple(((t[8:200, j + 1] - t[8:200, j]).float().mean().item()) for j in range(4)))
print("mma: wait-full %.0f  issue+commit %.0f" % tuple(((t[8:200, j + 1] - t[8:200, j]).float().mean().item()) for j in (5, 6)))

print("epilogue per tile: [wait start, acc ready, done] relative to kernel start")
for k in range(13):
    print(k, [int(v) - t0 for v in te[k][:3]], "wait %d  hold %d  work %d" % (int(te[k][1] - te[k][0]), int(te[k][3] - te[k][1]), int(te[k][2] - te[k][1])))
