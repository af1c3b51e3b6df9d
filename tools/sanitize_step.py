"""One cfg2-shaped BilateralConvFlex forward + backward (2 clouds of 2048 points, 64 channels, tile plan built on first
use so the engine-5 kernels run) plus one small lattice build -- the workload of tools/sanitize.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hplflownet_b200 as hpl  # noqa: E402
from hplflownet_b200 import plans  # noqa: E402
from hplflownet_b200.batching import concat_lattices  # noqa: E402
from hplflownet_b200.synthetic import frustum_pair  # noqa: E402
from hplflownet_b200.transforms import GenerateDataUnsymmetric  # noqa: E402

plans.PLAN_ON_FIRST_USE = True
dev = torch.device("cuda")
A = type("A", (), {"dim": 3, "scales_filter_map": [[1.0, 1, 1, 1], [0.5, 1, 1, 1]]})
gen = GenerateDataUnsymmetric(A())
items = []
for s in range(2):
    pc1, pc2 = frustum_pair(2048, s)
    items.append(gen([pc1, pc2, pc1])[3][0])
b = concat_lattices(items)
n = sum(b["point_counts"])
torch.manual_seed(0)
mod = hpl.BilateralConvFlex(3, 1, 64, [64], "cuda", use_bias=True, use_leaky=True, use_norm=True, do_splat=True,
                            do_slice=True, last_relu=False, chunk_size=-1).to(dev)
feat = torch.randn(1, 64, n, device=dev, requires_grad=True)
for _ in range(2):
    y = mod(feat, b["barycentric"], b["lattice_offset"], b["blur_neighbors"], b["barycentric"], b["lattice_offset"])
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("sanitize_step ok: H", sum(b["vertex_counts"]), "out", float(y.abs().max()))
