import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hplflownet_b200 import ops
from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
from tests._util import ModelArgs, golden, name_keyed_init_, rel_err
g = golden("model_frustum256.npz")
model = name_keyed_init_(HPLFlowNet(ModelArgs()), int(g["seed"])).cuda().eval()
gen = GenerateDataUnsymmetric(ModelArgs())
pc1, pc2, sf, gd = gen([g["pc1"], g["pc2"], np.zeros_like(g["pc1"])])
for prec in (1, 0):
    ops.DEFAULT_PRECISION = prec
    with torch.no_grad():
        out = model(pc1[None], pc2[None], collate_batch1(gd))
    print("precision", prec, "rel err vs reference fixture", rel_err(out, g["output"]))
