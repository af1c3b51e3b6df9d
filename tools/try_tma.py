"""Quick check + timing of the TMA-gathered contraction (engine 4) against float64 and engine 2.
    timeout 300 python tools/try_tma.py [quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hplflownet_b200 import ops  # noqa: E402

DEV = "cuda"


def reference(x, nbr, w, bias, act):
    xd = torch.cat((x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=x.device)), 0)
    f, c, co = w.shape
    g = xd[:x.size(0), :c][None] if nbr is None else xd[nbr.long()][:, :, :c]
    y = torch.einsum("fvc,fco->vo", g, w.double())
    if bias is not None:
        y = y + bias.double()
    if act == ops.ACT_LEAKY:
        y = torch.where(y > 0, y, 0.1 * y)
    elif act == ops.ACT_RELU:
        y = y.clamp_min(0)
    return y


def case(h, c, co, f, act, cm, idt=torch.int32):
    torch.manual_seed(h + c)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    nbr = None
    if f > 1:
        nbr = torch.randint(-1, h, (f, h), device=DEV, dtype=idt)
        nbr[0] = torch.arange(h, device=DEV)
    want = reference(x, nbr, w, bias, act)
    for prec in (4, 2):
        y = ops.blur_gemm(x, c, nbr, h, w, bias, act, out_channel_major=cm, precision=prec)
        torch.cuda.synchronize()
        got = y.t()[:, :co] if cm else y[:, :co]
        err = ((got.double() - want).abs().max() / want.abs().max()).item()
        print("h=%d c=%d co=%d f=%d act=%d cm=%d engine=%d  rel err %.3e %s" % (h, c, co, f, act, cm, prec, err,
                                                                              "OK" if err < 1e-5 else "FAIL"), flush=True)


def timing(h, c, co, f, reps=20):
    torch.manual_seed(0)
    x = ops.alloc_rows(h, c, DEV, zero=True)
    x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    # lattice-like locality: neighbours within a window of the vertex
    base = torch.arange(h, device=DEV)[None]
    nbr = (base + torch.randint(-400, 400, (f, h), device=DEV)).clamp(0, h - 1).to(torch.int32)
    nbr[torch.rand(f, h, device=DEV) < 0.1] = -1
    amax = ops.absmax(x)
    x16 = ops.h16_split(x, c, amax)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEV)
    for prec in (2, 4):
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.blur_gemm(x, c, nbr, h, w, bias, ops.ACT_LEAKY, precision=prec, x_amax=amax, x16=x16 if prec == 4 else None)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        print("timing h=%d c=%d co=%d f=%d engine=%d: median %.3f ms  min %.3f ms  %.1f TFLOP/s (incl. weight prep launches)" % (
            h, c, co, f, prec, med, ts[0], 2.0 * f * c * co * h / med / 1e9), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.h16_split(x, c, amax)
    e1.record()
    torch.cuda.synchronize()
    print("h16_split: %.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    if len(sys.argv) >= 2 and sys.argv[1] == "one":
        case(300, 64, 64, 15, ops.ACT_NONE, False)
        sys.exit(0)
    if len(sys.argv) >= 2 and sys.argv[1] == "prof":
        timing(242429, 64, 64, 15, reps=3)
        sys.exit(0)
    if len(sys.argv) >= 2 and sys.argv[1] == "time":
        timing(242429, 64, 64, 15)
        timing(31162, 580, 1024, 15, reps=5)
        sys.exit(0)
    case(300, 64, 64, 15, ops.ACT_NONE, False)
    case(7599, 64, 64, 15, ops.ACT_NONE, False, torch.int64)
    case(1000, 68, 64, 15, ops.ACT_LEAKY, False)
    case(333, 20, 32, 15, ops.ACT_RELU, True)
    case(4097, 128, 200, 1, ops.ACT_LEAKY, False)
    case(130, 580, 72, 15, ops.ACT_NONE, True)
    case(5, 4, 4, 15, ops.ACT_NONE, False)
    case(100000, 64, 64, 15, ops.ACT_LEAKY, False)
    case(3000, 580, 1024, 15, ops.ACT_LEAKY, False)
    if len(sys.argv) < 2 or sys.argv[1] == "time":
        timing(242429, 64, 64, 15)
        timing(31162, 580, 1024, 15, reps=5)
        timing(31162, 1024, 1024, 1, reps=5)
    print("done in %.1f s" % (time.time() - t0))
