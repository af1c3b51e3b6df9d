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


def test_model_backward_matches_float64_oracle():
    # whole-network gradients (the cfg5 training step's backward): CUDA path vs the oracle in float64
    from oracle import hplflownet as OM
    from oracle import lattice as OL
    from tests._util import assert_close_grad
    g = golden("model_frustum256.npz")
    model = name_keyed_init_(HPLFlowNet(ModelArgs()), int(g["seed"]))
    state = {k: (v.detach().clone().double().requires_grad_(True) if v.is_floating_point() else v.clone())
             for k, v in model.state_dict().items()}
    gd_np = OL.generate(g["pc1"], g["pc2"], ModelArgs.scales_filter_map)
    gd_ref = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd_np]
    gd_ref = [{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()} for d in gd_ref]
    p1, p2 = [torch.from_numpy(np.ascontiguousarray(g[k].T))[None] for k in ("pc1", "pc2")]
    target = torch.from_numpy(np.ascontiguousarray((g["pc2"] - g["pc1"]).T))[None]
    from tests._util import kink_ledger
    with kink_ledger() as ledger:
        out_ref = OM.forward(state, p1.double(), p2.double(), gd_ref)
    loss_ref = torch.norm(out_ref - target.double(), p=2, dim=1).mean()        # EPE3D loss, models/epe3d_loss.py:9
    loss_ref.backward()

    model = model.cuda().train()
    gen = GenerateDataUnsymmetric(ModelArgs())
    pc1, pc2, sf, gd = gen([g["pc1"], g["pc2"], g["pc2"] - g["pc1"]])
    out = model(pc1[None], pc2[None], collate_batch1(gd))
    loss = torch.norm(out - sf[None], p=2, dim=1).mean()
    loss.backward()
    assert_close(loss.detach(), loss_ref.detach(), "loss")
    checked = 0
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        assert_close_grad(p.grad, state[name].grad, "grad " + name, ledger)
        checked += 1
    assert checked >= 100


def test_shallow_model_forward_and_backward():
    """HPLFlowNetShallow: forward vs the reference model's own output (fixture), gradients vs the
    oracle in float64."""
    from hplflownet_b200.HPLFlowNet_shallow import HPLFlowNetShallow
    from oracle import hplflownet as OM
    from oracle import lattice as OL
    from tests._util import ShallowArgs, assert_close_grad
    g = golden("model_shallow_frustum256.npz")
    model = name_keyed_init_(HPLFlowNetShallow(ShallowArgs()), int(g["seed"]))
    state = {k: (v.detach().clone().double().requires_grad_(True) if v.is_floating_point() else v.clone())
             for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    gen = GenerateDataUnsymmetric(ShallowArgs())
    pc1, pc2, sf, gd = gen([g["pc1"], g["pc2"], g["pc2"] - g["pc1"]])
    with torch.no_grad():
        out = model(pc1[None], pc2[None], collate_batch1(gd))
    assert out.shape == (1, 3, 256)
    assert_close(out, g["output"], "flow")

    gd_np = OL.generate(g["pc1"], g["pc2"], ShallowArgs.scales_filter_map)
    gd_ref = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd_np]
    gd_ref = [{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()} for d in gd_ref]
    p1, p2 = [torch.from_numpy(np.ascontiguousarray(g[k].T))[None] for k in ("pc1", "pc2")]
    target = torch.from_numpy(np.ascontiguousarray((g["pc2"] - g["pc1"]).T))[None]
    from tests._util import kink_ledger
    with kink_ledger() as ledger:
        loss_ref = torch.norm(OM.forward_shallow(state, p1.double(), p2.double(), gd_ref) - target.double(), p=2, dim=1).mean()
    loss_ref.backward()
    model.train()
    loss = torch.norm(model(pc1[None], pc2[None], collate_batch1(gd)) - sf[None], p=2, dim=1).mean()
    loss.backward()
    assert_close(loss.detach(), loss_ref.detach(), "loss")
    checked = 0
    for name, p in model.named_parameters():
        if p.grad is not None:
            assert_close_grad(p.grad, state[name].grad, "grad " + name, ledger)
            checked += 1
    assert checked >= 60


def test_model_forward_baseline_size_matches_float64_oracle():
    """BASELINE configs[3]: full HPLFlowNet forward on an 8192+8192-point FlyingThings3D-shaped pair (all seven scales,
    the 256-wide contraction tiles and the K-split path are exercised here: H = 26 k / 35 k on the two finest levels),
    against the oracle composition evaluated in float64 on the host."""
    from hplflownet_b200.synthetic import frustum_pair
    from oracle import hplflownet as OM
    from oracle import lattice as OL
    pc1, pc2 = frustum_pair(8192, 7)
    model = name_keyed_init_(HPLFlowNet(ModelArgs()), 5)
    state = {k: (v.detach().clone().double() if v.is_floating_point() else v.clone()) for k, v in model.state_dict().items()}
    gd_np = OL.generate(pc1, pc2, ModelArgs.scales_filter_map)
    gd_ref = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd_np]
    gd_ref = [{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()} for d in gd_ref]
    p1, p2 = [torch.from_numpy(np.ascontiguousarray(p.T))[None].double() for p in (pc1, pc2)]
    with torch.no_grad():
        want = OM.forward(state, p1, p2, gd_ref)

    model = model.cuda().eval()
    gen = GenerateDataUnsymmetric(ModelArgs())
    a, b, sf, gd = gen([pc1, pc2, pc2 - pc1])
    for k in range(7):                       # the GPU builder's tables equal the oracle's (bit-exact contract)
        assert int(gd[k]["pc1_hash_cnt"]) == int(gd_np[k]["pc1_hash_cnt"])
    with torch.no_grad():
        out = model(a[None], b[None], collate_batch1(gd))
    assert out.shape == (1, 3, 8192)
    assert_close(out, want, "flow at BASELINE size")


def test_model_backward_baseline_size_matches_float64_oracle():
    """BASELINE configs[4] numerics: loss and every parameter gradient of one 8192+8192-point pair (EPE3D loss) against
    the oracle's autograd in float64 -- the wide weight-gradient tiles and split-K data gradients at their real sizes."""
    from hplflownet_b200.synthetic import frustum_pair
    from oracle import hplflownet as OM
    from oracle import lattice as OL
    from tests._util import assert_close_grad
    pc1, pc2 = frustum_pair(8192, 9)
    model = name_keyed_init_(HPLFlowNet(ModelArgs()), 6)
    state = {k: (v.detach().clone().double().requires_grad_(True) if v.is_floating_point() else v.clone())
             for k, v in model.state_dict().items()}
    gd_np = OL.generate(pc1, pc2, ModelArgs.scales_filter_map)
    gd_ref = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd_np]
    gd_ref = [{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()} for d in gd_ref]
    p1, p2 = [torch.from_numpy(np.ascontiguousarray(p.T))[None].double() for p in (pc1, pc2)]
    target = torch.from_numpy(np.ascontiguousarray((pc2 - pc1).T))[None].double()
    from tests._util import kink_ledger
    with kink_ledger() as ledger:
        loss_ref = torch.norm(OM.forward(state, p1, p2, gd_ref) - target, p=2, dim=1).mean()
    loss_ref.backward()

    model = model.cuda().train()
    gen = GenerateDataUnsymmetric(ModelArgs())
    a, b, sf, gd = gen([pc1, pc2, pc2 - pc1])
    loss = torch.norm(model(a[None], b[None], collate_batch1(gd)) - sf[None], p=2, dim=1).mean()
    loss.backward()
    assert_close(loss.detach(), loss_ref.detach(), "loss")
    checked = 0
    for name, p in model.named_parameters():
        if p.grad is not None:
            assert_close_grad(p.grad, state[name].grad, "grad " + name, ledger)
            checked += 1
    assert checked >= 100
