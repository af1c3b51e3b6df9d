"""GPU experiment: engine 5 (tile plans) against engines 0 / 2 on the cfg2 workload.  python tools/try_conv5.py [clouds]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hplflownet_b200 import ops, plans
ops.WEIGHT_CACHE = "always"          # kernel-only timings: weight images are built once  # noqa: E402
from hplflownet_b200.batching import concat_lattices  # noqa: E402


def ref64(x, nbr, w, bias):
    """float64 gather + matmul (torch, test infrastructure only)."""
    h = nbr.size(1)
    out = torch.zeros(h, w.size(2), dtype=torch.float64, device=x.device)
    xz = torch.cat([x.double(), torch.zeros(1, x.size(1), dtype=torch.float64, device=x.device)])
    for f in range(nbr.size(0)):
        idx = nbr[f].long()
        idx = torch.where(idx < 0, torch.full_like(idx, x.size(0)), idx)
        out += xz[idx] @ w[f].double()
    if bias is not None:
        out += bias.double()
    return out


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    c_in = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    c_out = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    dev = torch.device("cuda")
    items = [bench.cloud_tables(s) for s in range(B)]
    batch = concat_lattices(items)
    nbr = batch["blur_neighbors"][0].to(dev)                          # (15, H) int64
    h = nbr.size(1)
    print("clouds %d  H %d" % (B, h))

    t0 = time.perf_counter()
    plan = plans.build(nbr)
    torch.cuda.synchronize()
    print("plan: %.1f ms  sweeps %d  tiles %d  max_uniq %d  overflow %d  mean_uniq %.1f" %
          (1e3 * (time.perf_counter() - t0), plan.sweeps, plan.n_tiles, plan.max_uniq, plan.overflow,
           plan.sum_uniq / max(plan.n_tiles, 1)))
    ident = plans.build(nbr, order="identity")
    print("identity-order plan: max_uniq %d overflow %d mean %.1f" % (ident.max_uniq, ident.overflow, ident.sum_uniq / ident.n_tiles))

    # ---- plan consistency
    order = plan.order.long()
    assert torch.equal(torch.sort(order).values, torch.arange(h, device=dev)), "order is not a permutation"
    rows = plan.view(0).long()
    flat = rows.reshape(-1)
    assert torch.equal(flat[:h], order) and bool((flat[h:] == -1).all())
    if plan.usable:
        uniq, local, nu = plan.view(2).long(), plan.view(3).long() & 0xffff, plan.view(1).long()
        umax = uniq.size(1)
        for t in (0, plan.n_tiles // 2, plan.n_tiles - 1):
            r = rows[t]
            ok = r >= 0
            want = nbr[:, r.clamp(min=0)]                              # (15, 128)
            want = torch.where(ok[None], want, torch.full_like(want, -1))
            lt = local[t, :15]
            got = torch.where(lt == umax, torch.full_like(lt, -1), uniq[t][lt.clamp(max=umax - 1)])
            assert torch.equal(got, want), "tile %d: local/uniq do not reproduce the table" % t
            assert int(nu[t]) == len(torch.unique(want[want >= 0]))
        print("plan arrays reproduce the table")

    torch.manual_seed(0)
    x = torch.randn(h, c_in, device=dev) * 3.0
    w = (torch.randn(15, c_in, c_out, device=dev) * 0.05)
    bias = torch.randn(c_out, device=dev)
    want = ref64(x, nbr, w, bias)
    scale = want.abs().max().item()

    def err(y):
        return ((y[:, :c_out].double() - want).abs().max().item()) / scale

    y0 = ops.blur_gemm(x, c_in, nbr, h, w, bias, ops.ACT_NONE, precision=0)
    amax = ops.absmax(x)
    y2 = ops.blur_gemm(x, c_in, nbr, h, w, bias, ops.ACT_NONE, precision=2, x_amax=amax)
    print("engine 0 err %.2e   engine 2 err %.2e" % (err(y0), err(y2)))
    if not plan.usable:
        print("plan not usable; stop")
        return
    x16 = ops.h16b_split(x, c_in, amax)
    slot = ops.amax_slots(dev, 1)
    y5 = ops.conv5(x16, plan, c_in, w, bias, ops.ACT_NONE, amax, out_amax=slot)
    torch.cuda.synchronize()
    print("engine 5 err %.2e   out_amax %.6g (want %.6g)" % (err(y5), slot.view(torch.float32).item(), (want.abs().max().item())))
    y5b = ops.conv5(x16, plan, c_in, w, bias, ops.ACT_NONE, amax)
    print("engine 5 run-to-run bitwise identical:", torch.equal(y5, y5b))
    # leaky + loose amax bound (x8) must not matter
    amax8 = (amax.view(torch.float32) * 8).view(torch.int32)
    x16b = ops.h16b_split(x, c_in, amax8)
    y5c = ops.conv5(x16b, plan, c_in, w, bias, ops.ACT_LEAKY, amax8)
    wl = torch.where(want > 0, want, 0.1 * want)
    print("engine 5 leaky / 8x amax bound err %.2e" % (((y5c[:, :c_out].double() - wl).abs().max().item()) / scale))

    # ---- data gradient: dx[u] = sum_f dz[nbrT[f, u]] @ w[f]^T == conv5 over the SAME table with mirrored taps
    if c_in == c_out or True:
        dz = torch.randn(h, c_out, device=dev)
        tt = ops.transpose_table(nbr, h)
        wd = w.transpose(1, 2)                                          # (F, Co, C)
        dz_amax = ops.absmax(dz)
        d2 = ops.blur_gemm(dz, c_out, tt, h, wd, None, ops.ACT_NONE, precision=2, x_amax=dz_amax, tag="dgrad")
        want_d = ref64(dz, tt, wd, None)
        dz16 = ops.h16b_split(dz, c_out, dz_amax)
        d5 = ops.conv5(dz16, plan, c_out, wd, None, ops.ACT_NONE, dz_amax, mirror=True, tag="dgrad")
        sd = want_d.abs().max().item()
        print("dgrad: engine 2 err %.2e   engine 5 (mirrored taps) err %.2e" %
              ((d2[:, :c_in].double() - want_d).abs().max().item() / sd, (d5[:, :c_in].double() - want_d).abs().max().item() / sd))

    # ---- weight gradient
    dz = torch.randn(h, c_out, device=dev) * 0.7
    dz_amax = ops.absmax(dz)
    dz16 = ops.h16b_split(dz, c_out, dz_amax)
    xz = torch.cat([x.double(), torch.zeros(1, c_in, dtype=torch.float64, device=dev)])
    want_w = torch.stack([xz[torch.where(nbr[f] < 0, torch.full_like(nbr[f], h), nbr[f])].t() @ dz.double() for f in range(15)])
    dw2, _ = ops.blur_wgrad(x, c_in, nbr, h, dz, c_out, 15, want_db=False, precision=2, x_amax=amax, dz_amax=dz_amax)
    dw5 = ops.wgrad5(x16, dz16, plan, c_in, c_out, amax, dz_amax)
    sw = want_w.abs().max().item()
    print("wgrad: engine 2 err %.2e   engine 5 err %.2e" % ((dw2.double() - want_w).abs().max().item() / sw, (dw5.double() - want_w).abs().max().item() / sw))
    tw2 = timeit(lambda: ops.blur_wgrad(x, c_in, nbr, h, dz, c_out, 15, want_db=False, precision=2, x_amax=amax, dz_amax=dz_amax))
    tw5 = timeit(lambda: ops.wgrad5(x16, dz16, plan, c_in, c_out, amax, dz_amax))
    print("wgrad: engine 2 %.4f ms   engine 5 %.4f ms" % (tw2, tw5))

    wp = torch.nn.Parameter(w.clone())                                 # cached weight images: kernel-only timings
    w = ops.with_owner(wp.detach(), wp, "fwd")
    t2 = timeit(lambda: ops.blur_gemm(x, c_in, nbr, h, w, bias, ops.ACT_NONE, precision=2, x_amax=amax))
    t5 = timeit(lambda: ops.conv5(x16, plan, c_in, w, bias, ops.ACT_NONE, amax))
    ts = timeit(lambda: ops.h16b_split(x, c_in, amax))
    flops = 2.0 * 15 * c_in * c_out * h
    print("engine 2: %.4f ms (%.1f TFLOP/s)   engine 5: %.4f ms (%.1f TFLOP/s)   h16b split: %.4f ms" %
          (t2, flops / t2 / 1e9, t5, flops / t5 / 1e9, ts))


if __name__ == "__main__":
    main()
