import os, sys, torch
sys.path.insert(0, "/root/repo")
from hplflownet_b200 import ops
DEV="cuda"
def timing(h, c, co, f, reps=5):
    torch.manual_seed(0)
    x = ops.alloc_rows(h, c, DEV, zero=True); x[:, :c] = torch.randn(h, c, device=DEV)
    w = torch.randn(f, c, co, device=DEV) * (f * c) ** -0.5
    bias = torch.randn(co, device=DEV)
    nbr = None
    if f > 1:
        base = torch.arange(h, device=DEV)[None]
        nbr = (base + torch.randint(-400, 400, (f, h), device=DEV)).clamp(0, h - 1).to(torch.int32)
    amax = ops.absmax(x)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.blur_gemm(x, c, nbr, h, w, bias, ops.ACT_LEAKY, precision=2, x_amax=amax)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("TN256=%s h=%d c=%d co=%d f=%d: %.3f ms  %.1f TFLOP/s" % (os.environ.get("HPL_GEMM_TN256", "1"), h, c, co, f, ts[len(ts)//2], 2.0*f*c*co*h/ts[len(ts)//2]/1e9), flush=True)
def timing_w(h, c, co, f, reps=5):
    torch.manual_seed(0)
    x = ops.alloc_rows(h, c, DEV, zero=True); x[:, :c] = torch.randn(h, c, device=DEV)
    dz = ops.alloc_rows(h, co, DEV, zero=True); dz[:, :co] = torch.randn(h, co, device=DEV)
    nbr = None
    if f > 1:
        base = torch.arange(h, device=DEV)[None]
        nbr = (base + torch.randint(-400, 400, (f, h), device=DEV)).clamp(0, h - 1).to(torch.int32)
    ax, az = ops.absmax(x), ops.absmax(dz)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.blur_wgrad(x, c, nbr, h, dz, co, f, want_db=False, precision=2, x_amax=ax, dz_amax=az)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("wgrad TN256=%s h=%d c=%d co=%d f=%d: %.3f ms  %.1f TFLOP/s" % (os.environ.get("HPL_GEMM_TN256", "1"), h, c, co, f, ts[len(ts)//2], 2.0*f*c*co*h/ts[len(ts)//2]/1e9), flush=True)
timing_w(31162, 580, 1024, 15)
timing_w(31162, 1024, 1024, 1)
timing_w(52600, 324, 512, 15)
timing_w(14500, 388, 256, 15)
timing(31162, 580, 1024, 15)
timing(31162, 1024, 1024, 1)
timing(52600, 324, 512, 15)
timing(14500, 388, 256, 15)
timing(31162, 1024, 580, 15)
