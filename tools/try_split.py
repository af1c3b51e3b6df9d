"""GPU experiment: the fused split pass with and without re-zeroing its input accumulator (cfg2 x 32 size, cold L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hplflownet_b200 import ops
dev = torch.device("cuda")
h = 242429
raw = torch.randn(h, 64, device=dev)
wsum = torch.rand(h, device=dev) + 0.5
amax = ops.absmax(raw)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)        # 256 MB > L2


def timeit(fn, n=10):
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for d in (0, 2):
    inv = torch.empty(h, device=dev)
    slot = ops.amax_slots(dev, 1)
    t = timeit(lambda: ops.h16b_split_ex(raw.clone() if False else raw, 64, amax, norm=wsum, inv_out=inv, norm_amax_out=slot, dispose=d))
    print("forward split, dispose=%d: %.4f ms" % (d, t))
    raw = torch.randn(h, 64, device=dev)
print("memset 62 MB: %.4f ms" % timeit(lambda: raw.zero_()))
