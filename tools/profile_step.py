"""Workloads for ncu captures (profiles/): one cfg2 x B step through the module API (tile plan prepared), one correlation
layer step (BASELINE configs[2]) and one 7-scale lattice build.  python tools/profile_step.py [step|corr|lattice] [clouds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import hplflownet_b200 as hpl  # noqa: E402
from hplflownet_b200 import plans  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "step"
dev = torch.device("cuda")
if what == "step":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    mod = bench.make_state().to(dev)
    host, res, gy, n_tot, h_tot = bench.make_batch(dev, list(range(B)))
    plans.prepare(res["blur_neighbors"])
    for it in range(3):
        for p in mod.parameters():
            p.grad = None
        res["features"].grad = None
        y = mod(res["features"], res["barycentric"], res["lattice_offset"], res["blur_neighbors"], res["barycentric"], res["lattice_offset"])
        y.backward(gy)
    torch.cuda.synchronize()
    print("step ok: H", h_tot)
elif what == "corr":
    print(bench.corr_leg(dev))
else:
    print(bench.lattice_leg(dev))
