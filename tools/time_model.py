"""Time the full HPLFlowNet forward (8192+8192 pts) for each contraction engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hplflownet_b200 import ops
dev = torch.device("cuda", 0)
for prec in (2, 3, 1):
    ops.DEFAULT_PRECISION = prec
    r = bench.model_leg(dev)
    print("precision", prec, {k: round(v, 2) for k, v in r.items() if k.endswith("_ms")})
