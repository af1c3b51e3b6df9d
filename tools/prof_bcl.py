"""Run a few BCL fwd+bwd steps on the bench workload (for ncu captures). Usage: prof_bcl.py [clouds] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hplflownet_b200.batching import concat_lattices

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
mod = bench.make_state().to(dev)
batch = concat_lattices([bench.cloud_tables(s) for s in range(B)])
n_tot = sum(batch["point_counts"])
feat = torch.randn(1, bench.CHANNELS, n_tot, device=dev, requires_grad=True)
gy = torch.randn(1, bench.CHANNELS, n_tot, device=dev)
bary, off, nbr = [batch[k].to(dev) for k in ("barycentric", "lattice_offset", "blur_neighbors")]
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("steps")
for _ in range(steps):
    y = mod(feat, bary, off, nbr, bary, off)
    y.backward(gy)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", n_tot, nbr.shape)
