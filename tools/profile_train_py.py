"""GPU experiment: training step of the full model (8 pairs, one GPU): wall time, kernel time (torch.profiler) and the
Python-side hot spots (cProfile)."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import train as T
from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.synthetic import frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1


class A:
    dim = 3
    evaluate = False
    use_leaky = bcn_use_bias = bcn_use_norm = True
    last_relu = False
    DEVICE = "cuda"
    scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1], [.125, 1, 1, 1], [.0625, 1, 1, 1]]


dev = torch.device("cuda")
torch.manual_seed(0)
model = HPLFlowNet(A()).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
pairs = []
for i in range(8):
    pc1, pc2 = frustum_pair(bench.N_POINTS, 500 + i)
    pairs.append((pc1, pc2, (pc2 - pc1).astype("float32")))
buckets = T.GradBuckets(model.parameters(), bucket_mb=16.0)
for _ in range(2):
    T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
torch.cuda.synchronize()
print("train step: %.1f ms wall" % (1e3 * (time.perf_counter() - t0) / 3))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
    torch.cuda.synchronize()
tot = sum(e.device_time_total for e in prof.key_averages())
print("kernel time in one step: %.1f ms" % (tot / 1e3))
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    print("  %7.2f ms  %5d x  %s" % (e.device_time_total / 1e3, e.count, e.key[:90]))
pr = cProfile.Profile()
pr.enable()
T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
