"""The direct-gather contraction kernels in isolation (through the ops layer / C ABI): the tcgen05 engines (4 = TMA /
cp.async staged pre-split operands, 2 = register-staged 3xFP16) and the fp32 CUDA-core anchor (0) against a float64
torch reference of the same gather-GEMM.  Engine 5 (tile plans) has its own file, tests/test_gpu_plan.py."""
import pytest
import torch

from hplflownet_b200 import ops
from tests._util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference(x, nbr, w, bias, act):
    xd = torch.cat((x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=x.device)), 0)
    f, c, co = w.shape
    if nbr is None:
        g = xd[:x.size(0), :c][None]
    else:
        g = xd[nbr.long()][:, :, :c]               # -1 -> the appended zero row
    y = torch.einsum("fvc,fco->vo", g, w.double())
    if bias is not None:
        y = y + bias.double()
    if act == ops.ACT_LEAKY:
        y = torch.where(y > 0, y, 0.1 * y)
    elif act == ops.ACT_RELU:
        y = y.clamp_min(0)
    return y


@pytest.mark.parametrize("precision", [4, 2, 0])
@pytest.mark.parametrize("h,c,co,f,act,cm", [
    (7599, 64, 64, 15, ops.ACT_NONE, False),     # cfg2 blur layer
    (1000, 68, 64, 15, ops.ACT_LEAKY, False),    # bcn1: K per tap not a multiple of 16
    (333, 20, 32, 15, ops.ACT_RELU, True),       # Co < tile, channel-major output
    (4097, 128, 200, 1, ops.ACT_LEAKY, False),   # 1x1 layer, several N tiles, ragged Co
    (130, 580, 72, 15, ops.ACT_NONE, True),      # bcn1_-like K = 8700
    (5, 4, 4, 15, ops.ACT_NONE, False),          # tiny
])
def test_gather_gemm_matches_float64(precision, h, c, co, f, act, cm):
    torch.manual_seed(h + c)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    nbr = None
    if f > 1:
        nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int32)
        nbr[0] = torch.arange(h, device=DEV)
    y = ops.blur_gemm(x, c, nbr, h, w, bias, act, out_channel_major=cm, precision=precision)
    got = y.t()[:, :co] if cm else y[:, :co]
    assert_close(got, _reference(x, nbr, w, bias, act), "gather-gemm precision=%d" % precision)


@pytest.mark.parametrize("precision", [2, 0])
@pytest.mark.parametrize("h,c,co,f", [
    (7599, 64, 64, 15),       # cfg2
    (242429 // 4, 64, 64, 15),  # a batch of clouds: many vertex ranges per CTA column
    (1000, 68, 64, 15),       # M tiles straddle taps
    (333, 20, 32, 15),
    (4097, 128, 200, 1),      # 1x1 layer, ragged Co
    (130, 580, 72, 15),       # K = 8700 forward, M = 8700 here
    (5, 4, 4, 15),
])
def test_wgrad_matches_float64(precision, h, c, co, f):
    torch.manual_seed(h + co)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    dz = ops.alloc_rows(h, co, DEV, zero=True)
    dz[:, :co] = torch.randn(h, co, device=DEV)
    nbr = None
    if f > 1:
        nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int64)
    dw, db = ops.blur_wgrad(x, c, nbr, h, dz, co, f, precision=precision)
    xd = torch.cat((x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=DEV)), 0)
    g = xd[:h, :c][None] if nbr is None else xd[nbr][:, :, :c]
    want = torch.einsum("fvc,vo->fco", g, dz[:, :co].double())
    assert_close(dw, want, "wgrad precision=%d" % precision)
    assert_close(db, dz[:, :co].double().sum(0), "bias grad")


def test_tma_engine_many_tiles_and_int64_table():
    """Engine 4 over several work items per persistent CTA (double-buffered index blocks and TMEM accumulators
    wrap around), int64 table, run twice: results must be identical between runs and match float64."""
    h, c, co, f = 40000, 64, 64, 15
    torch.manual_seed(7)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int64)
    y1 = ops.blur_gemm(x, c, nbr, h, w, bias, ops.ACT_LEAKY, precision=4).clone()
    y2 = ops.blur_gemm(x, c, nbr, h, w, bias, ops.ACT_LEAKY, precision=4)
    assert torch.equal(y1, y2)
    assert_close(y1[:, :co], _reference(x, nbr, w, bias, ops.ACT_LEAKY), "engine 4, 313 tiles")


@pytest.mark.parametrize("h,c", [(1000, 64), (333, 20), (5, 4), (130, 580)])
def test_h16_split_reconstructs_fp32(h, c):
    """hpl_h16_split: (hi + lo * 2^-11) * s reproduces x to 2^-22 relative of max|x|; pad channels and the extra row are 0."""
    torch.manual_seed(c)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV) * 37.0
    amax = ops.absmax(x)
    img = ops.h16_split(x, c, amax)
    ld16 = (c + 7) // 8 * 8
    planes = img.view(torch.float16).view(h + 1, 2, ld16).float()
    m = x.abs().max().item()
    import math
    s = 2.0 ** (math.floor(math.log2(m)) - 13)
    rec = (planes[:, 0] + planes[:, 1] * 2.0 ** -11) * s
    assert (rec[:h, :c] - x[:, :c]).abs().max().item() <= m * 2.0 ** -21
    assert rec[h].abs().max().item() == 0 and (c == ld16 or rec[:, c:].abs().max().item() == 0)


@pytest.mark.parametrize("h,c,co,f,act,cm", [
    (9000, 64, 256, 15, ops.ACT_LEAKY, False),     # 256-wide tile, one K range (30 K blocks)
    (8200, 580, 300, 15, ops.ACT_LEAKY, False),    # K = 8700 split over 4 CTAs per tile (RED partial sums), ragged Co
    (8300, 324, 512, 15, ops.ACT_NONE, True),      # split K, channel-major output (bcn2_-like)
    (8192, 1024, 512, 1, ops.ACT_RELU, False),     # 1x1 layer (conv3-like)
])
def test_wide_tile_engine2_matches_float64(h, c, co, f, act, cm):
    """Engine 2's 256-wide tile (Co >= 256, >= 8192 rows): 16 producer warps, one CTA per SM, K split over blockIdx.z
    with bias / activation / max|out| applied by the follow-up pass."""
    torch.manual_seed(h + co)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    nbr = None
    if f > 1:
        nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int32)
        nbr[0] = torch.arange(h, device=DEV)
    slot = ops.amax_slots(DEV, 1)
    y = ops.blur_gemm(x, c, nbr, h, w, bias, act, out_channel_major=cm, precision=2, out_amax=slot)
    got = y.t()[:, :co] if cm else y[:, :co]
    want = _reference(x, nbr, w, bias, act)
    assert_close(got, want, "wide tile")
    assert slot.view(torch.float32).item() == got.abs().max().item()           # the fused statistic is exact
    # strided weight operand: the same result from a permuted view of the conv-layout weight
    w_conv = w.permute(2, 1, 0).contiguous()                                    # (Co, C, F) like nn.Conv2d's weight
    y2 = ops.blur_gemm(x, c, nbr, h, w_conv.permute(2, 1, 0), bias, act, out_channel_major=cm, precision=2)
    assert_close(y2.t()[:, :co] if cm else y2[:, :co], want, "strided weight")


@pytest.mark.parametrize("h,c,co,f", [(7599, 64, 64, 15), (4097, 128, 200, 1), (333, 20, 32, 15)])
def test_epilogue_amax_is_exact(h, c, co, f):
    """The GEMM epilogue's fused statistic equals max|out| (it scales the next layer's FP16 split)."""
    torch.manual_seed(h)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV) * 300.0
    w = torch.randn(f, c, co, device=DEV)
    nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int32) if f > 1 else None
    slot = ops.amax_slots(DEV, 1)
    y = ops.blur_gemm(x, c, nbr, h, w, None, ops.ACT_LEAKY, precision=2, out_amax=slot)
    assert slot.view(torch.float32).item() == y[:, :co].abs().max().item() > 1000.0


def test_stack_scales_follow_large_activations():
    """Two chained layers whose intermediate activations are ~1e5 (beyond fp16 range unless the second layer's operand
    scale comes from the first layer's fused statistic)."""
    from hplflownet_b200 import _stack
    h, c = 5000, 64
    torch.manual_seed(3)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV) * 1.0e4
    nbr = torch.randint(-1, h, (15, h), device=DEV, dtype=torch.int32)
    w1, w2 = torch.randn(15, c, 64, device=DEV), torch.randn(1, 64, 32, device=DEV) * 0.1
    b1, b2 = torch.randn(64, device=DEV), torch.randn(32, device=DEV)
    xs, chans, _, _ = _stack.forward(x, c, h, [(w1, b1, ops.ACT_LEAKY), (w2, b2, ops.ACT_NONE)], nbr)
    mid = _reference(x, nbr, w1, b1, ops.ACT_LEAKY)
    assert mid.abs().max().item() > 65504.0
    want = mid @ w2[0].double() + b2.double()
    assert_close(xs[-1][:, :32], want, "chained layers")


@pytest.mark.parametrize("h,c,co,f", [
    (8200, 64, 256, 15),        # 256-wide weight-gradient tile: 16 producer warps, one main accumulator
    (8300, 324, 512, 15),       # M tiles straddle taps, two N tiles
    (8192, 1024, 300, 1),       # 1x1 layer, ragged Co (second N tile 44 wide)
    (20000, 68, 260, 15),       # several vertex ranges per tile column, ragged everything
])
def test_wide_tile_wgrad_matches_float64(h, c, co, f):
    torch.manual_seed(h + co)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    dz = ops.alloc_rows(h, co, DEV, zero=True)
    dz[:, :co] = torch.randn(h, co, device=DEV)
    nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=torch.int32) if f > 1 else None
    dw, db = ops.blur_wgrad(x, c, nbr, h, dz, co, f, precision=2)
    xd = torch.cat((x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=DEV)), 0)
    g = xd[:h, :c][None] if nbr is None else xd[nbr.long()][:, :, :c]
    want = torch.einsum("fvc,vo->fco", g, dz[:, :co].double())
    assert_close(dw, want, "wide wgrad")
    assert_close(db, dz[:, :co].double().sum(0), "bias grad")


def test_wide_tile_cluster_multicast_variant():
    """HPL_WIDE_CLUSTER=1: the 256-wide tile launched as thread-block clusters of two M tiles that share the weight
    stream (cp.async.bulk multicast + multicast tcgen05.commit).  The knob is read once per process -> subprocess; an odd
    number of M tiles (idle partner CTA) and the split-K / ragged-Co case are included."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_gemm.py", "-q", "-m", "gpu", "-x", "-k", "test_wide_tile_engine2_matches_float64"]
    out = subprocess.run(cmd, cwd=root, env=dict(os.environ, HPL_WIDE_CLUSTER="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "4 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
