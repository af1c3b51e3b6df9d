"""GPU experiment: where the Python side of the eager HPLFlowNet forward spends its time (cProfile, resident lattice)."""
import cProfile, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import ops
from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.synthetic import frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1


class A:
    dim = 3
    evaluate = True
    use_leaky = bcn_use_bias = bcn_use_norm = True
    last_relu = False
    DEVICE = "cuda"
    scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1], [.125, 1, 1, 1], [.0625, 1, 1, 1]]


dev = torch.device("cuda")
torch.manual_seed(0)
model = HPLFlowNet(A()).to(dev).eval()
gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
pc1, pc2 = frustum_pair(bench.N_POINTS, 7)
a, b = torch.from_numpy(pc1.T.copy()).to(dev), torch.from_numpy(pc2.T.copy()).to(dev)
gd = collate_batch1(gen.build(a, b))
with ops.weight_cache_scope(), torch.no_grad():
    for _ in range(4):
        model(a[None], b[None], gd)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        model(a[None], b[None], gd)
    torch.cuda.synchronize()
    print("eager forward: %.3f ms" % (1e3 * (time.perf_counter() - t0) / 20))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        model(a[None], b[None], gd)
    torch.cuda.synchronize()
    pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.print_callers("current_stream")
st.print_callers("zeros")
