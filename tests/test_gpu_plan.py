"""Tile plans (csrc/plan.cu) and contraction engine 5 (csrc/gemm_plan.cu) through the ops layer / C ABI:
plan arrays against the table they were built from, and forward / mirrored-tap data gradient / weight gradient
against a float64 torch evaluation of the same gather-GEMM (models/bilateralNN.py:198-221 and its autograd)."""
import numpy as np
import pytest
import torch

from hplflownet_b200 import ops, plans
from hplflownet_b200.synthetic import box_cloud, frustum_pair
from oracle import lattice as OL
from tests._util import assert_close
from tests.test_gpu_gemm import _reference

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _table(n_points, seed, scale=1.0, clouds=1, kind="frustum", dtype=torch.int64):
    """blur_neighbors of `clouds` concatenated lattices (vertex ids shifted per cloud), built by the oracle."""
    tabs, base = [], 0
    for k in range(clouds):
        if kind == "frustum":
            pc1, pc2 = frustum_pair(n_points, seed + k)
        else:
            pc1, pc2 = box_cloud(n_points, seed + k), box_cloud(n_points, seed + k + 100)
        d = OL.generate(pc1, pc2, [[scale, 1, -1, -1]])[0]
        nb = torch.from_numpy(np.asarray(d["pc1_blur_neighbors"]).astype(np.int64))
        tabs.append(torch.where(nb >= 0, nb + base, nb))
        base += nb.size(1)
    return torch.cat(tabs, 1).to(dtype).to(DEV).contiguous()


def _check_plan(plan, nbr):
    h = nbr.size(1)
    order = plan.order.long()
    assert torch.equal(torch.sort(order).values, torch.arange(h, device=DEV)), "order is not a permutation"
    rows = plan.view(0).long()
    assert torch.equal(rows.reshape(-1)[:h], order) and bool((rows.reshape(-1)[h:] == -1).all())
    uniq, local, nu = plan.view(2).long(), plan.view(3).long() & 0xffff, plan.view(1).long()
    umax = uniq.size(1)
    want = nbr.long()[:, rows.clamp(min=0)]                                  # (F, tiles, 128)
    want = torch.where(rows[None] >= 0, want, torch.full_like(want, -1)).permute(1, 0, 2)
    lt = local[:, :nbr.size(0)]
    got = torch.where(lt == umax, torch.full_like(lt, -1), torch.gather(uniq, 1, lt.clamp(max=umax - 1).reshape(lt.size(0), -1)).view_as(lt))
    assert torch.equal(got, want), "local / uniq do not reproduce the table"
    for t in range(0, plan.n_tiles, max(1, plan.n_tiles // 7)):
        assert int(nu[t]) == len(torch.unique(want[t][want[t] >= 0]))
    assert bool((local[:, nbr.size(0):] == umax).all())                      # padded taps read the zero row


@pytest.mark.parametrize("n,scale,clouds,kind,dtype", [
    (8192, 1.0, 1, "frustum", torch.int64),      # cfg2 lattice
    (8192, 1.0, 3, "frustum", torch.int32),      # a batch: components must not mix
    (2048, 1.0, 1, "box", torch.int64),          # cfg1-like
    (8192, 3.0, 1, "frustum", torch.int64),      # finest level of the net: ~1800 connected components
    (60, 1.0, 1, "frustum", torch.int64),        # fewer vertices than one tile... (several components, tiny)
])
def test_plan_reproduces_table_and_is_local(n, scale, clouds, kind, dtype):
    nbr = _table(n, 3, scale, clouds, kind, dtype)
    plan = plans.build(nbr)
    assert plan.usable and plan.symmetric, (plan.max_uniq, plan.overflow)
    _check_plan(plan, nbr)
    ident = plans.build(nbr, order="identity")
    assert plan.sum_uniq <= ident.sum_uniq                                   # the Morton order never gathers more rows
    if n >= 2048 and scale == 1.0:
        assert plan.sum_uniq / plan.n_tiles < 400 < ident.sum_uniq / ident.n_tiles


def test_plan_rejects_a_table_without_locality_and_an_asymmetric_one():
    h = 5000
    nbr = torch.randint(-1, h, (15, h), device=DEV, dtype=torch.int64)
    plan = plans.build(nbr)
    assert not plan.usable and plan.overflow > 0 and not plan.symmetric
    # a lattice table with one entry redirected: still local, no longer its own mirrored transpose
    nbr = _table(2048, 5)
    v = int((nbr[3] >= 0).nonzero()[0])
    nbr[3, v] = nbr[4, v] if int(nbr[4, v]) >= 0 else 0
    plan = plans.build(nbr)
    assert plan.usable and not plan.symmetric


@pytest.mark.parametrize("c,co,act,clouds", [
    (64, 64, ops.ACT_NONE, 2),       # cfg2
    (68, 64, ops.ACT_LEAKY, 1),      # channels not a multiple of 32 (zero-padded block)
    (20, 32, ops.ACT_RELU, 1),       # Co < 64
    (192, 48, ops.ACT_LEAKY, 1),     # 180 accumulate steps: two main accumulators per tile
    (4, 4, ops.ACT_NONE, 1),
])
def test_engine5_forward_dgrad_wgrad_match_float64(c, co, act, clouds):
    nbr = _table(4096, 11, clouds=clouds)
    h = nbr.size(1)
    plan = plans.build(nbr)
    assert plan.usable and plan.symmetric
    torch.manual_seed(c + co)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV) * 2.5
    w = torch.randn(15, c, co, device=DEV) * (15 * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    amax = ops.absmax(x)
    x16 = ops.h16b_split(x, c, amax)
    slot = ops.amax_slots(DEV, 1)
    y = ops.conv5(x16, plan, c, w, bias, act, amax, out_amax=slot)
    want = _reference(x, nbr, w, bias, act)
    assert_close(y[:, :co], want, "engine 5 forward")
    torch.cuda.synchronize()
    assert abs(slot.view(torch.float32).item() - want.abs().max().item()) <= 1e-5 * want.abs().max().item()
    y2 = ops.conv5(x16, plan, c, w, bias, act, amax)
    assert torch.equal(y, y2), "engine 5 is deterministic"
    # conv weight layout (Co, C, F, 1) viewed as (F, C, Co): strided weight operand
    w_conv = w.permute(2, 1, 0).contiguous()
    assert_close(ops.conv5(x16, plan, c, w_conv.permute(2, 1, 0), bias, act, amax)[:, :co], want, "strided weight")

    dz = ops.alloc_rows(h, co, DEV, zero=True)
    dz[:, :co] = torch.randn(h, co, device=DEV)
    dz_amax = ops.absmax(dz)
    dz16 = ops.h16b_split(dz, co, dz_amax)
    tt = ops.transpose_table(nbr, h)
    wd = w.transpose(1, 2)
    if ops.conv5_supported(15, co, c):                      # (the transposed shape Co -> C must fit the kernel: C <= 64)
        dx = ops.conv5(dz16, plan, co, wd, None, ops.ACT_NONE, dz_amax, mirror=True)
        assert_close(dx[:, :c], _reference(dz, tt, wd, None, ops.ACT_NONE), "engine 5 data gradient (mirrored taps)")

    dw = ops.wgrad5(x16, dz16, plan, c, co, amax, dz_amax)
    xd = torch.cat((x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=DEV)), 0)
    want_w = torch.einsum("fvc,vo->fco", xd[nbr.long()][:, :, :c], dz[:, :co].double())
    assert_close(dw, want_w, "engine 5 weight gradient")


def test_engine5_many_tiles_per_sm_and_fused_normalisation():
    """cfg2 x 8 clouds (more tiles than SMs, several flushes of the weight-gradient accumulators) with the density
    normalisation fused into the operand split (bilateralNN.py:185-186) and a loose power-of-two scale bound."""
    nbr = _table(8192, 21, clouds=8, dtype=torch.int32)
    h = nbr.size(1)
    plan = plans.build(nbr)
    torch.manual_seed(0)
    raw = torch.randn(h, 64, device=DEV)
    wsum = torch.rand(h, device=DEV) * 3 + 0.2
    xn = raw / (wsum[:, None] + 1e-5)
    bound = (ops.absmax(xn).view(torch.float32) * 37.0).view(torch.int32)      # any bound within 2^10 is fine
    x16 = ops.h16b_split(raw, 64, bound, norm=wsum)
    w = torch.randn(15, 64, 64, device=DEV) * 0.03
    y = ops.conv5(x16, plan, 64, w, None, ops.ACT_NONE, bound)
    assert_close(y, _reference(xn, nbr, w, None, ops.ACT_NONE), "fused normalisation")
    dz = torch.randn(h, 64, device=DEV) * 1e-4                                   # small gradients: the scale must follow
    dz_amax = ops.absmax(dz)
    dw = ops.wgrad5(x16, ops.h16b_split(dz, 64, dz_amax), plan, 64, 64, bound, dz_amax)
    xd = torch.cat((xn.double(), torch.zeros(1, 64, dtype=torch.float64, device=DEV)), 0)
    assert_close(dw, torch.einsum("fvc,vo->fco", xd[nbr.long()], dz.double()), "weight gradient, 8 clouds")


_VARIANT_SCRIPT = r"""
import torch
from hplflownet_b200 import ops, plans
from tests.test_gpu_plan import _table
from tests.test_gpu_gemm import _reference
from tests._util import assert_close
nbr = _table(8192, 5, clouds=4)
plan = plans.build(nbr)
torch.manual_seed(1)
x = torch.randn(nbr.size(1), 64, device="cuda") * 1.7
w = torch.randn(15, 64, 64, device="cuda") * 0.03
b = torch.randn(64, device="cuda")
amax = ops.absmax(x)
y = ops.conv5(ops.h16b_split(x, 64, amax), plan, 64, w, b, ops.ACT_LEAKY, amax)
assert_close(y, _reference(x, nbr, w, b, ops.ACT_LEAKY), "engine 5 variant")
print("variant ok")
"""


@pytest.mark.parametrize("knob", ["0", "1", "2"])
def test_engine5_operand_paths(knob):
    """The three builds of the forward kernel -- A operand in shared memory (0), in tensor memory with one-tap (1) or
    two-tap (2, default) stages -- on a table with several tiles per CTA.  The knob is read once per process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HPL_CONV5_TMEM=knob)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_engine5_at_bench_size_against_float64_rows_and_engine2():
    """BASELINE configs[1] at the bench's size (32 concatenated 8192-point clouds, H ~ 242 k: 13 tiles per CTA, every
    ring slot and accumulator buffer reused many times): forward, mirrored-tap data gradient and weight gradient of
    engine 5 against a float64 evaluation on a sample of rows / the full reduction, and against engine 2."""
    import bench
    from hplflownet_b200.batching import concat_lattices
    nbr = concat_lattices([bench.cloud_tables(s) for s in range(32)])["blur_neighbors"][0].to(DEV)
    h = nbr.size(1)
    plan = plans.build(nbr)
    assert plan.usable and plan.symmetric and plan.n_tiles > 12 * 148
    torch.manual_seed(3)
    x = torch.randn(h, 64, device=DEV) * 3.0
    w = torch.randn(15, 64, 64, device=DEV) * 0.04
    b = torch.randn(64, device=DEV)
    amax = ops.absmax(x)
    x16 = ops.h16b_split(x, 64, amax)
    y = ops.conv5(x16, plan, 64, w, b, ops.ACT_LEAKY, amax)
    rows = torch.randint(0, h, (4096,), device=DEV)
    xd = torch.cat((x.double(), torch.zeros(1, 64, dtype=torch.float64, device=DEV)), 0)
    pre = torch.einsum("fvc,fco->vo", xd[nbr[:, rows].long()], w.double()) + b.double()
    want = torch.where(pre > 0, pre, 0.1 * pre)
    scale = y.abs().max().item()
    assert (y[rows].double() - want).abs().max().item() <= 1e-5 * scale, "forward, sampled rows"
    y2 = ops.blur_gemm(x, 64, nbr, h, w, b, ops.ACT_LEAKY, precision=2, x_amax=amax)
    assert (y - y2).abs().max().item() <= 2e-5 * scale, "engine 5 vs engine 2"
    dz = torch.randn(h, 64, device=DEV)
    dz_amax = ops.absmax(dz)
    dz16 = ops.h16b_split(dz, 64, dz_amax)
    dx = ops.conv5(dz16, plan, 64, w.transpose(1, 2), None, ops.ACT_NONE, dz_amax, mirror=True)
    tt = ops.transpose_table(nbr, h)
    zd = torch.cat((dz.double(), torch.zeros(1, 64, dtype=torch.float64, device=DEV)), 0)
    want_dx = torch.einsum("fvo,fco->vc", zd[tt[:, rows].long()], w.double())
    assert (dx[rows].double() - want_dx).abs().max().item() <= 1e-5 * dx.abs().max().item(), "data gradient, sampled rows"
    dw = ops.wgrad5(x16, dz16, plan, 64, 64, amax, dz_amax)
    want_dw = torch.zeros(15, 64, 64, dtype=torch.float64, device=DEV)
    for f in range(15):                                                          # (tap by tap: 242 k x 64 doubles at a time)
        want_dw[f] = xd[nbr[f].long()].t() @ dz.double()
    assert_close(dw, want_dw, "weight gradient, full reduction")
