"""Effect of the POINT ORDER on the BCL step: random order (what the reference's loaders deliver) vs points sorted
along a Morton curve before the lattice build (first-occurrence vertex numbering then follows the curve, so a tile
of 128 consecutive vertices is spatially compact and its 15 x 128 gathered rows overlap).
    python tools/try_sorted.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hplflownet_b200 import ops  # noqa: E402
from hplflownet_b200.batching import concat_lattices  # noqa: E402
from hplflownet_b200.synthetic import frustum_pair  # noqa: E402
from hplflownet_b200.transforms import GenerateDataUnsymmetric  # noqa: E402


def morton_order(pc, cell=0.25):
    q = np.floor((pc - pc.min(0)) / cell).astype(np.uint64)
    code = np.zeros(len(pc), np.uint64)
    for b in range(16):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")


def run(sort, B=32, steps=30):
    dev = torch.device("cuda", 0)

    class A:
        dim = 3
        scales_filter_map = [[1.0, 1, -1, -1]]
    gen = GenerateDataUnsymmetric(A())
    items = []
    for s in range(B):
        pc1, pc2 = frustum_pair(8192, s)
        if sort:
            o = morton_order(pc1)
            pc1, pc2 = pc1[o], pc2[o]
        d = gen([pc1, pc2, pc1])[3][0]
        items.append({k: (v.cpu() if not isinstance(v, int) else v) for k, v in d.items()})
    batch = concat_lattices(items)
    n_tot = sum(batch["point_counts"])
    mod = bench.make_state().to(dev)
    feat = torch.randn(1, 64, n_tot, device=dev, requires_grad=True)
    gy = torch.randn(1, 64, n_tot, device=dev)
    bary, off, nbr = [batch[k].to(dev) for k in ("barycentric", "lattice_offset", "blur_neighbors")]
    params = list(mod.parameters())

    def step():
        for p in params:
            p.grad = None
        feat.grad = None
        y = mod(feat, bary, off, nbr, bary, off)
        y.backward(gy)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ops.PROFILE_GEMM = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    by = {}
    for tag, a, b in ops.PROFILE_GEMM:
        by.setdefault(tag, []).append(a.elapsed_time(b))
    ops.PROFILE_GEMM = None
    print("sorted=%d  H=%d  %.3f ms/step  %.0f clouds/s  gemm ms %s" % (
        sort, sum(batch["vertex_counts"]), ms, B / ms * 1e3, {k: round(sum(v) / len(v), 4) for k, v in by.items()}), flush=True)


if __name__ == "__main__":
    run(False)
    run(True)
    for eng in (4,):
        ops.DEFAULT_PRECISION = eng
        print("engine", eng)
        run(False)
        run(True)
