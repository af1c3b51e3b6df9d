"""One HPLFlowNet training step (1 pair) for ncu launch lists.  Usage: prof_train.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hplflownet_b200 import train as T
from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.synthetic import frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
from tests._util import ModelArgs


class A(ModelArgs):
    evaluate = False


dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = HPLFlowNet(A()).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pairs = []
for i in range(n):
    pc1, pc2 = frustum_pair(8192, 500 + i)
    pairs.append((pc1, pc2, (pc2 - pc1).astype("float32")))
T.train_step(model, opt, gen, pairs, collate_batch1)          # warm-up (allocator, Adam state)
torch.cuda.synchronize()
torch.cuda.profiler.start()
T.train_step(model, opt, gen, pairs, collate_batch1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
