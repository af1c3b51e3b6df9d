"""GPU experiment: the CSR-gather splat (hpl_h16b_splat_csr) against the RED splat + fused split, cfg2 x 32 size, cold L2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import ops, plans
dev = torch.device("cuda")
_, res, gy, n_tot, h = bench.make_batch(dev, list(range(32)))
feat = res["features"][0].detach().contiguous()
bary, off = res["barycentric"][0].contiguous(), res["lattice_offset"][0].contiguous()
splan = plans.prepare_splat(off, h)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, n=8):
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


slot = ops.amax_slots(dev, 2)
rows = ops.cm_to_rows(feat, amax=slot[0:1])
inv = torch.empty(h, device=dev)
print("contributions per row: mean %.2f max %d" % ((splan.ptr[1:] - splan.ptr[:-1]).float().mean().item(), int((splan.ptr[1:] - splan.ptr[:-1]).max())))
print("cm_to_rows (transpose + max): %.4f ms" % timeit(lambda: ops.cm_to_rows(feat, amax=slot[0:1])))
print("h16b_splat_csr (normalised):  %.4f ms" % timeit(lambda: ops.h16b_splat_csr(rows, 64, bary, splan, slot[0:1], normalize=True, inv_out=inv, norm_amax_out=slot[1:2])))
raw = torch.zeros(h, 64, device=dev)
wsum = torch.zeros(h, device=dev)
print("scatter_rows (RED):           %.4f ms" % timeit(lambda: ops.scatter_rows(feat, bary, off, h, True, in_amax=slot[0:1], rows=raw, wsum=wsum)))
print("h16b_split_ex (dispose 0):    %.4f ms" % timeit(lambda: ops.h16b_split_ex(raw, 64, slot[0:1], norm=wsum, inv_out=inv, norm_amax_out=slot[1:2], dispose=0)))
