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


def test_pointwise_stack_matches_torch_conv1d():
    # kernel-size-1 Conv1d stacks on the CUDA GEMM vs the stock op in float64, forward and backward
    from hplflownet_b200.module_utils import Conv1dReLU
    from hplflownet_b200.pointwise import pointwise_stack
    from tests._util import assert_close_grad
    torch.manual_seed(0)
    mods = torch.nn.ModuleList([Conv1dReLU(3, 32, use_leaky=True), Conv1dReLU(32, 64, use_leaky=True),
                                torch.nn.Conv1d(64, 5, kernel_size=1)]).cuda()
    x = torch.randn(1, 3, 3001, device="cuda", requires_grad=True)
    gy = torch.randn(1, 5, 3001, device="cuda")
    y = pointwise_stack(mods, x)
    y.backward(gy)
    got = [y.detach(), x.grad.clone()] + [p.grad.clone() for p in mods.parameters()]
    ref_mods = torch.nn.ModuleList([Conv1dReLU(3, 32, use_leaky=True), Conv1dReLU(32, 64, use_leaky=True),
                                    torch.nn.Conv1d(64, 5, kernel_size=1)]).double().cuda()
    ref_mods.load_state_dict({k: v.double() for k, v in mods.state_dict().items()})
    xr = x.detach().double().requires_grad_(True)
    h = xr
    for m in ref_mods:
        h = m(h)
    h.backward(gy.double())
    want = [h.detach(), xr.grad] + [p.grad for p in ref_mods.parameters()]
    assert_close(got[0], want[0], "output")
    for a, b in zip(got[1:], want[1:]):
        assert_close_grad(a, b, "grad")
