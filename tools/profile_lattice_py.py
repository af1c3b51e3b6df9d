"""GPU experiment: where the Python side of the GPU lattice build (7 scales, one pair) spends its time."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200.synthetic import frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric


class A:
    dim = 3
    scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1], [.125, 1, 1, 1], [.0625, 1, 1, 1]]


dev = torch.device("cuda")
gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
pc1, pc2 = frustum_pair(bench.N_POINTS, 7)
a, b = torch.from_numpy(pc1.T.copy()).to(dev), torch.from_numpy(pc2.T.copy()).to(dev)
for _ in range(5):
    gen.build(a, b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    gen.build(a, b)
torch.cuda.synchronize()
print("build: %.3f ms per pair" % (1e3 * (time.perf_counter() - t0) / 50))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    gen.build(a, b)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
