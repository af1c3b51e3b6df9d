"""GPU parity of BilateralCorrelationFlex (through the C ABI): reference fixtures + the oracle
on seeded inputs, including the BASELINE cfg3 shape (8192 + 8192 points, C = C' = 64)."""
import pytest
import torch

import hplflownet_b200 as hpl
from hplflownet_b200.synthetic import frustum_pair
from oracle import bcl as OB
from oracle import lattice as OL
from tests._util import assert_close, assert_close_grad, golden, golden_files, grads_from, kink_ledger, oracle_state, state_from, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", golden_files("corr_"))
def test_corr_matches_reference_fixture(name):
    g = golden(name)
    c, prev_dim, use_leaky, last_relu = [int(x) for x in g["cfg"]]
    mod = hpl.BilateralCorrelationFlex(3, 1, 1, c, [int(x) for x in g["corr_out"]], [int(x) for x in g["out_ch"]],
                                       "cuda", use_bias=True, use_leaky=bool(use_leaky), use_norm=True,
                                       prev_corr_dim=prev_dim, last_relu=bool(last_relu), chunk_size=-1)
    mod.load_state_dict(state_from(g), strict=True)
    mod = mod.to(DEV)
    f1, f2 = t(g["feat1"], DEV).requires_grad_(True), t(g["feat2"], DEV).requires_grad_(True)
    prev = t(g["prev_corr_feat"], DEV).requires_grad_(True) if prev_dim else None
    bary, off = t(g["barycentric1"], DEV), t(g["lattice_offset1"], DEV)
    y = mod(f1, f2, prev, bary if prev_dim else None, off if prev_dim else None, t(g["pc1_corr_indices"], DEV),
            t(g["pc2_corr_indices"], DEV), f1.size(-1), f2.size(-1))
    assert_close(y, g["output"], "output")
    y.backward(t(g["grad_output"], DEV))
    assert_close(f1.grad, g["grad_feat1"], "grad_feat1")
    assert_close(f2.grad, g["grad_feat2"], "grad_feat2")
    if prev_dim:
        assert_close(prev.grad, g["grad_prev"], "grad_prev")
    got = dict(mod.named_parameters())
    for k, ref in grads_from(g).items():
        assert_close(got[k].grad, ref, "grad " + k)


@pytest.mark.parametrize("n,c,prev_dim,idx_dtype", [
    (1024, 16, 16, torch.int64),
    (8192, 64, 64, torch.int32),      # BASELINE configs[2]: corr_conv [32,32], blur_conv [64,64]
])
def test_corr_matches_oracle(n, c, prev_dim, idx_dtype):
    pc1, pc2 = frustum_pair(n, 2)
    d = OL.generate(pc1, pc2, [[1.0, 1, 1, 1]])[0]
    h1, h2 = d["pc1_hash_cnt"], d["pc2_hash_cnt"]
    torch.manual_seed(0)
    mod = hpl.BilateralCorrelationFlex(3, 1, 1, c, [32, 32], [64, 64], "cuda", use_bias=True, use_leaky=True,
                                       use_norm=True, prev_corr_dim=prev_dim, last_relu=False, chunk_size=-1)
    state = oracle_state(mod)
    mod = mod.to(DEV)
    f1, f2, prev = torch.randn(1, c, h1), torch.randn(1, c, h2), torch.randn(1, prev_dim, n)
    gy = torch.randn(1, 64, h1)
    bary, off = torch.from_numpy(d["pc1_barycentric"])[None], torch.from_numpy(d["pc1_lattice_offset"])[None]
    i1, i2 = torch.from_numpy(d["pc1_corr_indices"])[None], torch.from_numpy(d["pc2_corr_indices"])[None]

    # oracle evaluated in float64 (see tests/_util.py:oracle_state)
    r1, r2, rp = [x.double().requires_grad_(True) for x in (f1, f2, prev)]
    with kink_ledger() as ledger:
        y_ref = OB.corr_forward(state, r1, r2, rp, bary.double(), off, i1, i2, use_norm=True, use_leaky=True)
    y_ref.backward(gy.double())

    g1, g2, gp = [x.to(DEV).requires_grad_(True) for x in (f1, f2, prev)]
    y = mod(g1, g2, gp, bary.to(DEV), off.to(DEV).to(idx_dtype), i1.to(DEV).to(idx_dtype),
            i2.to(DEV).to(idx_dtype), h1, h2)
    y.backward(gy.to(DEV))
    assert y.shape == (1, 64, h1)
    assert_close(y, y_ref.detach(), "output")
    assert_close_grad(g1.grad, r1.grad, "grad_feat1", ledger)
    assert_close_grad(g2.grad, r2.grad, "grad_feat2", ledger)
    assert_close_grad(gp.grad, rp.grad, "grad_prev", ledger)
    for k, p in mod.named_parameters():
        assert_close_grad(p.grad, state[k].grad, "grad " + k, ledger)
