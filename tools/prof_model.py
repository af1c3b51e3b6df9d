import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
with torch.no_grad():
    for _ in range(2):
        y = model(a[None], b[None], gd)
torch.cuda.synchronize()
print("H:", [d["pc1_hash_cnt"].item() for d in gd])
