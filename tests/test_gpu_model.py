"""Full HPLFlowNet forward on the B200 layers (SURVEY §8f-1 / BASELINE configs[3] at reduced size):
GPU lattice builder + CUDA bilateral/correlation layers vs the reference model's own output."""
import numpy as np
import pytest
import torch

from hplflownet_b200.HPLFlowNet import HPLFlowNet
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
from tests._util import ModelArgs, assert_close, golden, name_keyed_init_

pytestmark = pytest.mark.gpu


def test_model_forward_matches_reference_output():
    g = golden("model_frustum256.npz")
    model = name_keyed_init_(HPLFlowNet(ModelArgs()), int(g["seed"])).cuda().eval()
    gen = GenerateDataUnsymmetric(ModelArgs())
    pc1, pc2, sf, gd = gen([g["pc1"], g["pc2"], np.zeros_like(g["pc1"])])
    with torch.no_grad():
        out = model(pc1[None], pc2[None], collate_batch1(gd))
    assert out.shape == (1, 3, 256)
    # 1e-5 relative (north_star) through 19 bilateral/correlation layers and K up to 8700
    assert_close(out, g["output"], "flow")
