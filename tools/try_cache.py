import sys; sys.path.insert(0, "/root/repo")
import torch, time
from hplflownet_b200 import ops, _lib
from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.synthetic import frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
from tests._util import ModelArgs
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = HPLFlowNet(ModelArgs()).to(dev).eval()
gen = GenerateDataUnsymmetric(ModelArgs(), device=dev, index_dtype=torch.int32)
pc1, pc2 = frustum_pair(8192, 7)
a, b = torch.from_numpy(pc1.T.copy()).to(dev), torch.from_numpy(pc2.T.copy()).to(dev)
gd = collate_batch1(gen.build(a, b))
for cache in (False, True):
    ops.WEIGHT_CACHE = cache
    with torch.no_grad():
        for _ in range(3): model(a[None], b[None], gd)
        torch.cuda.synchronize(); _lib.launch_count = 0; t0 = time.perf_counter()
        for _ in range(8): model(a[None], b[None], gd)
        torch.cuda.synchronize()
    print("cache", cache, "forward ms", (time.perf_counter() - t0) / 8 * 1e3, "hpl launches per fwd (claimed)", _lib.launch_count / 8, "cached images", len(ops._weight_images))
