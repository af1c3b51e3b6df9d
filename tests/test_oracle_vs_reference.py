"""The oracle against the UNMODIFIED reference, live (build container only: skipped when /root/reference is absent,
e.g. on the GPU box -- the committed fixtures of tests/golden pin the same functions there).

  * lattice index path: every entry of ``generated_data`` for an 8192+8192-point pair over the 7 scales of
    configs/test_ours_FlyingThings3D.yaml, bit for bit (transforms/transforms.py:358-485 vs oracle/lattice_oracle.c);
  * value path: BilateralConvFlex at BASELINE configs[1] (cfg2) and BilateralCorrelationFlex at configs[2] (cfg3, reduced
    point count to keep the CPU suite short), outputs and gradients (models/bilateralNN.py, models/bnn_flow.py vs
    oracle/bcl.py).
"""
import os

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout only exists in the build container")


@pytest.fixture(scope="module")
def ref():
    from oracle import make_golden as MG
    return MG, MG.import_reference()


def test_lattice_8192_seven_scales_bit_exact(ref):
    MG, (T, _, _) = ref
    from hplflownet_b200.synthetic import frustum_pair
    from oracle import lattice as OL
    from tests._util import bits_equal
    pc1, pc2 = frustum_pair(8192, 17)
    want = MG.ref_generate(T, pc1, pc2, MG.FULL_SFM)
    got = OL.generate(pc1, pc2, MG.FULL_SFM)
    assert len(got) == len(want) == 7
    for k, (g, w) in enumerate(zip(got, want)):
        assert set(g) == set(w)
        for key, v in w.items():
            if isinstance(v, int):
                assert g[key] == v, (k, key)
            else:
                assert bits_equal(g[key], v.numpy()), (k, key)


def test_bcl_cfg2_values_and_gradients(ref):
    MG, (T, BCL, _) = ref
    from hplflownet_b200.synthetic import frustum_pair
    from oracle import bcl as OB
    from tests._util import assert_close
    pc1, pc2 = frustum_pair(8192, 2)
    d = MG.ref_generate(T, pc1, pc2, [[1.0, 1, -1, -1]])[0]
    torch.manual_seed(0)
    mod = BCL(3, 1, 64, [64], "cpu", use_bias=True, use_leaky=True, use_norm=True, do_splat=True, do_slice=True,
              last_relu=False, chunk_size=-1)
    with torch.no_grad():
        mod.bias.normal_(0, 0.3)
    feat = torch.randn(1, 64, 8192, requires_grad=True)
    gy = torch.randn(1, 64, 8192)
    bary, off, nbr = d["pc1_barycentric"][None], d["pc1_lattice_offset"][None], d["pc1_blur_neighbors"][None]
    y = mod(feat, bary, off, nbr, bary, off)
    y.backward(gy)
    state = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
    f2 = feat.detach().clone().requires_grad_(True)
    y2 = OB.bcl_forward(state, f2, bary, off, nbr, bary, off, do_splat=True, do_slice=True, use_norm=True,
                        use_leaky=True, use_bias=True)
    y2.backward(gy)
    assert_close(y2.detach(), y.detach(), "output")
    assert_close(f2.grad, feat.grad, "grad_features")
    for k, p in mod.named_parameters():
        assert_close(state[k].grad, p.grad, "grad " + k)


def test_corr_cfg3_values_and_gradients(ref):
    MG, (T, _, Corr) = ref
    from hplflownet_b200.synthetic import frustum_pair
    from oracle import bcl as OB
    from tests._util import assert_close
    n = 1024                                   # cfg3's module on fewer points (172.8 KB per vertex in the reference)
    pc1, pc2 = frustum_pair(n, 5)
    d = MG.ref_generate(T, pc1, pc2, [[1.0, 1, 1, 1]])[0]
    h1, h2 = d["pc1_hash_cnt"], d["pc2_hash_cnt"]
    torch.manual_seed(1)
    mod = Corr(3, 1, 1, 64, [32, 32], [64, 64], "cpu", use_bias=True, use_leaky=True, use_norm=True, prev_corr_dim=64,
               last_relu=False, chunk_size=-1)
    f1 = torch.randn(1, 64, h1, requires_grad=True)
    f2 = torch.randn(1, 64, h2, requires_grad=True)
    prev = torch.randn(1, 64, n, requires_grad=True)
    args = (d["pc1_barycentric"][None], d["pc1_lattice_offset"][None], d["pc1_corr_indices"][None], d["pc2_corr_indices"][None])
    y = mod(f1, f2, prev, *args, h1, h2)
    gy = torch.randn_like(y)
    y.backward(gy)
    state = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
    g1, g2, gp = (t.detach().clone().requires_grad_(True) for t in (f1, f2, prev))
    y2 = OB.corr_forward(state, g1, g2, gp, *args, use_norm=True, use_leaky=True)
    y2.backward(gy)
    assert_close(y2.detach(), y.detach(), "output")
    for a, b, what in ((g1, f1, "feat1"), (g2, f2, "feat2"), (gp, prev, "prev")):
        assert_close(a.grad, b.grad, "grad " + what)
    for k, p in mod.named_parameters():
        assert_close(state[k].grad, p.grad, "grad " + k)
