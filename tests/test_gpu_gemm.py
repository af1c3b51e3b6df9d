"""The contraction kernels in isolation (through the ops layer / C ABI): the tcgen05 engines (4 = TMA / cp.async
staged pre-split operands, 3, 2 = 3xFP16, 1 = 3xTF32) and the fp32 CUDA-core anchor (0) against a float64
torch reference of the same gather-GEMM."""
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


@pytest.mark.parametrize("precision", [4, 3, 2, 1, 0])
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


@pytest.mark.parametrize("precision", [3, 2, 1, 0])
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
