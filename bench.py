#!/usr/bin/env python
"""bench.py -- point-clouds/sec of one BilateralConvFlex forward+backward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clouds B]

Workload (BASELINE configs[1]): BilateralConvFlex(3, 1, 64, [64], splat+slice,
norm, bias, LeakyReLU, last_relu=False) on FlyingThings3D-shaped frustum clouds of 8192 points,
scale-1.0 lattice (in-lattice = out-lattice).  One *step* = forward+backward over a batch of B
distinct clouds per GPU, concatenated into one launch sequence (the reference is B=1 only).
Multi-GPU: clouds shard by sample, one process per GPU, no data-path collective (weak scaling).

`value`  : clouds/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the module call with HOST (pinned) inputs: features + the index
           tables the reference's DataLoader ships per sample are copied H2D every step, the
           loss scalar and the parameter gradients are read back D2H.
`--impl reference`: the reference's CPU algorithm (oracle/bcl.py port, torch CPU ops on all host
           threads) on the same config, a bounded sample of clouds per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "point-clouds/sec (BCL fwd+bwd, 8192 pts, d=3, 64ch)"
N_POINTS, CHANNELS, SCALE = 8192, 64, 1.0


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(n, h, c, co, f=15, d1=4):
    """bytes each stage must move once (fp32 values, 4-byte indices)."""
    fwd = 4 * (n * c + 2 * d1 * n + h * c) + 4 * (h * c + f * h + f * c * co + co + h * co) + \
        4 * (h * co + 2 * d1 * n + n * co)
    bwd = 4 * (n * co + 2 * d1 * n + h * co) + 4 * (h * co + h * c + f * h + f * c * co) + \
        4 * (h * c + f * c * co + co) + 4 * (h * c + 2 * d1 * n + h + n * c)
    return fwd, bwd


def blur_fwd_bytes(h, c, co, f=15):
    return 4 * (h * c + f * h + f * c * co + co + h * co)


def make_state(seed=0):
    """Reference-format state_dict of the cfg2 module (default nn.Conv2d init under manual_seed)."""
    import torch
    import hplflownet_b200 as hpl
    torch.manual_seed(seed)
    mod = hpl.BilateralConvFlex(3, 1, CHANNELS, [CHANNELS], "cuda", use_bias=True, use_leaky=True,
                                use_norm=True, do_splat=True, do_slice=True, last_relu=False, chunk_size=-1)
    with torch.no_grad():
        mod.bias.normal_(0, 0.1)
    return mod


_GEN = None


def cloud_tables(seed):
    """Index tables of one synthetic cloud, built by the CUDA lattice builder (product path),
    returned as HOST tensors in the reference's format (int64) -- what its DataLoader delivers."""
    global _GEN
    from hplflownet_b200.synthetic import frustum_pair
    from hplflownet_b200.transforms import GenerateDataUnsymmetric
    if _GEN is None:
        class A:
            dim = 3
            scales_filter_map = [[SCALE, 1, -1, -1]]
        _GEN = GenerateDataUnsymmetric(A())
    pc1, pc2 = frustum_pair(N_POINTS, seed)
    d = _GEN([pc1, pc2, pc1])[3][0]
    return {k: (v.cpu() if not isinstance(v, int) else v) for k, v in d.items()}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import bcl as OB
    from oracle import lattice as OL
    from hplflownet_b200.synthetic import frustum_pair
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    mod = make_state()
    state = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
    clouds_per_step = 1
    samples = []
    for s in range(4):
        pc1, pc2 = frustum_pair(N_POINTS, s)
        d = OL.generate(pc1, pc2, [[SCALE, 1, -1, -1]])[0]
        torch.manual_seed(s)
        samples.append((torch.randn(1, CHANNELS, N_POINTS), torch.from_numpy(d["pc1_barycentric"])[None],
                        torch.from_numpy(d["pc1_lattice_offset"])[None],
                        torch.from_numpy(d["pc1_blur_neighbors"])[None], torch.randn(1, CHANNELS, N_POINTS),
                        d["pc1_hash_cnt"]))

    def step(i):
        feat, bary, off, nbr, gy, _ = samples[i % len(samples)]
        f = feat.clone().requires_grad_(True)
        for v in state.values():
            v.grad = None
        y = OB.bcl_forward(state, f, bary, off, nbr, bary, off, do_splat=True, do_slice=True, use_norm=True,
                           use_leaky=True, use_bias=True)
        y.backward(gy)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = clouds_per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BilateralConvFlex(3,1,64,[64]) fwd+bwd, 8192-pt frustum cloud, scale 1.0",
                   "clouds_per_step": clouds_per_step, "H": [s[5] for s in samples]},
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d cloud (oracle/bcl.py, torch CPU ops, %d threads)"
                                   % (args.steps, clouds_per_step, cores)},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on this workload, from the committed ncu
    summary (profiles/ncu_traffic.json, written from an `ncu --set full` capture; null when the kernel is not in it)."""
    p = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(kernel)
    return (e["dram_bytes"], e.get("source")) if e else (None, None)


def make_batch(dev, cloud_ids, scale=SCALE):
    """Concatenated tables of the given clouds + features / output gradient, resident on `dev`; host copies pinned."""
    import torch
    from hplflownet_b200.batching import concat_lattices
    global _GEN
    if scale != SCALE:
        saved, _GEN = _GEN, None
        A = type("A", (), {"dim": 3, "scales_filter_map": [[scale, 1, -1, -1]]})
        from hplflownet_b200.transforms import GenerateDataUnsymmetric
        _GEN = GenerateDataUnsymmetric(A())
        items = [cloud_tables(s) for s in cloud_ids]
        _GEN = saved
    else:
        items = [cloud_tables(s) for s in cloud_ids]
    batch = concat_lattices(items)
    n_tot, h_tot = sum(batch["point_counts"]), sum(batch["vertex_counts"])
    torch.manual_seed(1000 + cloud_ids[0])
    host = {
        "features": torch.randn(1, CHANNELS, n_tot).pin_memory(),
        "barycentric": batch["barycentric"].pin_memory(),
        "lattice_offset": batch["lattice_offset"].pin_memory(),
        "blur_neighbors": batch["blur_neighbors"].pin_memory(),
    }
    gy = torch.randn(1, CHANNELS, n_tot, device=dev)
    resident = {k: v.to(dev) for k, v in host.items()}
    resident["features"].requires_grad_(True)
    return host, resident, gy, n_tot, h_tot


def time_steps(fn, steps, warmup, barrier):
    import torch
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import hplflownet_b200 as hpl
    from hplflownet_b200 import _lib, ops, plans, sharding
    from hplflownet_b200.graphs import GraphedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = sharding.bind_to_gpu_numa_node(local) if world > 1 else None      # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.clouds

    # ---- synthetic batch: B distinct clouds for this rank (weak scaling: B per GPU)
    mod = make_state().to(dev)
    host, resident, gy, n_tot, h_tot = make_batch(dev, sharding.cloud_ids(rank, world, B))
    params = list(mod.parameters())

    def fwd_bwd(t=resident, g=gy):
        for p in params:
            p.grad = None
        t["features"].grad = None
        y = mod(t["features"], t["barycentric"], t["lattice_offset"], t["blur_neighbors"],
                t["barycentric"], t["lattice_offset"])
        y.backward(g)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM.  The tables of the resident batch are planned once, outside the timed region
    # (a per-lattice precomputation like the tables themselves; csrc/plan.cu).  The headline loops rebuild the weight
    # images on every call (ops.weight_cache_scope would skip two small kernels per GEMM because the weights never change
    # here; the number stays conservative).
    t0 = time.perf_counter()
    plan = plans.prepare(resident["blur_neighbors"])
    torch.cuda.synchronize()
    plan_ms = 1e3 * (time.perf_counter() - t0)
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("HPL_BENCH_NO_SMI") != "1":
        sampler.start()                                  # (nvidia-smi needs ~100 ms for its first sample; warm-up is the same load)
    _lib.launch_count = 0
    fwd_bwd()
    launches_per_step = _lib.launch_count
    eager_ms = time_steps(fwd_bwd, args.steps, args.warmup, barrier)
    # the same step as ONE CUDA-graph launch (every kernel of the forward and the backward on the capturing stream)
    graph_ms, graph_err = None, None
    try:
        graphed = GraphedStep(fwd_bwd)
        profiling = os.environ.get("HPL_BENCH_PROFILE") == "1"  # ncu --profile-from-start off: launch list of the timed loop
        barrier()
        if profiling:
            torch.cuda.profiler.start()
        graph_ms = time_steps(graphed.replay, args.steps, args.warmup, barrier)
        if profiling:
            torch.cuda.profiler.stop()
    except Exception as e:                               # noqa: BLE001 -- fall back to the eager number
        graph_err = "%s: %s" % (type(e).__name__, e)
        torch.cuda.synchronize()
    ms = graph_ms if graph_ms is not None else eager_ms
    # separate, untimed pass: CUDA events around the contraction and the splat / slice kernels
    ops.PROFILE_GEMM, ops.PROFILE_ROWS = [], True
    for _ in range(5):
        fwd_bwd()
    torch.cuda.synchronize()
    gemm_events = ops.PROFILE_GEMM
    ops.PROFILE_GEMM, ops.PROFILE_ROWS = None, False
    clocks = sampler.stop() if rank == 0 else None

    # ---- latency of ONE cloud (the reference's B = 1 usage) and the scale-3.0 point (H ~ 26-31 k per cloud)
    extra = {}
    if rank == 0:
        for name, ids, scale in (("latency_1cloud", [900], SCALE), ("cfg2_scale3", list(range(700, 708)), 3.0)):
            try:
                _, res1, gy1, n1, h1 = make_batch(dev, ids, scale)
                plans.prepare(res1["blur_neighbors"])
                f1 = lambda: fwd_bwd(res1, gy1)          # noqa: E731
                e_ms = time_steps(f1, 20, 5, torch.cuda.synchronize) / 20
                g1 = GraphedStep(f1)
                g_ms = time_steps(g1.replay, 50, 5, torch.cuda.synchronize) / 50
                extra[name] = {"clouds": len(ids), "scale": scale, "H": h1, "ms_graph": g_ms, "ms_eager": e_ms,
                               "clouds_per_s": len(ids) / (g_ms * 1e-3)}
                del g1, res1, gy1
            except Exception as e:                       # noqa: BLE001
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.synchronize()

    # ---- e2e: host (pinned) inputs; every step uploads its own inputs H2D and reads its loss + parameter
    # gradients back D2H, all inside the timed region.  The upload of step i+1 is enqueued on a copy stream
    # before step i computes (what a pin_memory DataLoader does), so PCIe and the SMs overlap.  Tables that arrive
    # fresh every step are not planned (plans.py: a table is planned on its second use), so this path runs engine 2.
    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_run(n_steps, host=host):
        def upload():
            with torch.cuda.stream(copy_stream):
                t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return t, ev
        nxt = upload()
        for i in range(n_steps):
            t, ev = nxt
            if i + 1 < n_steps:
                nxt = upload()
            torch.cuda.current_stream().wait_event(ev)
            for v in t.values():
                v.record_stream(torch.cuda.current_stream())
            t["features"].requires_grad_(True)
            for p in params:
                p.grad = None
            y = mod(t["features"], t["barycentric"], t["lattice_offset"], t["blur_neighbors"],
                    t["barycentric"], t["lattice_offset"])
            loss = (y * gy).sum()
            loss.backward()
            out = [loss.detach().cpu()] + [p.grad.cpu() for p in params]     # D2H, synchronises the step
        return out

    e2e_run(max(3, args.warmup // 2))
    barrier()
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = 4 + sum(p.numel() * 4 for p in params)
    # supplementary: the same loop with the index tables in the native int32 format (the module accepts both; the
    # reference's DataLoader ships int64, which is what the headline e2e uploads)
    host32 = dict(host)
    for k in ("lattice_offset", "blur_neighbors"):
        host32[k] = host[k].to(torch.int32).pin_memory()
    e2e_run(3, host32)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps, host32)
    barrier()
    e2e32_s = time.perf_counter() - t0
    h2d32 = sum(v.numel() * v.element_size() for v in host32.values())
    # host -> device copy ceiling of this rank while every rank uploads (explains the e2e scaling: the ranks share the
    # host's memory / PCIe fabric)
    big = host["features"]
    barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        big.to(dev, non_blocking=True)
    barrier()
    h2d_gbs_rank = 10 * big.numel() * 4 / (time.perf_counter() - t0) / 1e9

    # ---- secondary, all ranks: data-parallel HPLFlowNet training step (BASELINE configs[4])
    train = train_leg(dev, rank, world)

    # ---- max over ranks
    ms, eager_ms, e2e_s, e2e32_s = sharding.max_over_ranks([ms, eager_ms, e2e_s, e2e32_s], device=dev)
    (h2d_gbs_min,) = sharding.max_over_ranks([-h2d_gbs_rank], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = sharding.job_throughput(B * args.steps, world, ms * 1e-3)
    e2e_value = sharding.job_throughput(B * e2e_steps, world, e2e_s)

    # ---- roofline of the dominant kernel: the blur contraction (forward launch), tcgen05 3xFP16.
    # achieved = algorithmic FLOPs (2*F*C*Co per vertex) / CUDA-event duration on the launch
    # stream; peak = measured dense bf16 tensor throughput (MEASURED_PEAKS.json, burst).  A 3xFP16
    # contraction issues 3 MMA-equivalents per algorithmic MAC: ceiling peak/3.
    peaks = measured_peaks()
    by_tag = {}
    for tag, a, b in gemm_events:
        by_tag.setdefault(tag, []).append(a.elapsed_time(b))
    avg = {k: sum(v) / len(v) for k, v in by_tag.items()}
    gemm_ms = avg.get("fwd", float("nan"))
    gemm_flops = 2.0 * 15 * CHANNELS * CHANNELS * h_tot
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    fb, bb = algorithmic_bytes(n_tot, h_tot, CHANNELS, CHANNELS)
    all_gemm_ms = sum(avg.get(t, 0.0) for t in ("fwd", "dgrad", "wgrad"))
    big_five_ms = all_gemm_ms + 2 * avg.get("scatter", 0.0) + 2 * avg.get("gather", 0.0)
    # the bandwidth-bound kernels of the step against the measured HBM copy peak (algorithmic bytes, SURVEY 8d:
    # splat / slice-backward scatter 4(N C + 2 d1 N + H C), slice / splat-backward gather 4(H C + 2 d1 N + N C))
    row_bytes = 4.0 * (n_tot * CHANNELS + 8 * n_tot + h_tot * CHANNELS)
    hbm_kernels = {}
    for tag, kernel in (("scatter", "scatter_rows_kernel"), ("gather", "gather_rows_kernel")):
        if tag in avg:
            gbs = row_bytes / (avg[tag] * 1e-3) / 1e9
            traffic, _ = ncu_traffic(kernel)
            hbm_kernels[tag] = {"kernel": kernel, "kernel_ms": avg[tag], "algorithmic_bytes": row_bytes, "achieved_gbs": gbs,
                                "peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "traffic": traffic}
    engine = {5: "tcgen05 3xFP16 on a per-lattice tile plan: distinct neighbour rows staged once per 128-vertex tile, "
                 "pre-split operands, 2 MMAs per K step",
              4: "tcgen05 3xFP16, persistent, operands pre-split in HBM, staged by cp.async / TMA",
              2: "tcgen05 3xFP16 (scaled hi/lo split)", 0: "fp32 CUDA-core FMA"}[ops.DEFAULT_PRECISION]
    kname = {5: "conv5_kernel", 4: "gather_gemm_tma_kernel", 2: "gather_gemm_f16_kernel",
             0: "gather_gemm_kernel"}[ops.DEFAULT_PRECISION]
    if ops.DEFAULT_PRECISION == 5 and not plan.usable:
        kname, engine = "gather_gemm_f16_kernel", "tcgen05 3xFP16 (scaled hi/lo split); the table has no usable tile plan"
    traffic, traffic_src = ncu_traffic(kname)
    step_ms = ms / args.steps
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peaks["source"],
        "kernel": "%s (blur forward, %s)" % (kname, engine),
        "kernel_ms": gemm_ms, "kernel_ms_by_role": avg,
        "frac_of_3_mma_ceiling": achieved / (peaks["bf16_tflops"] / 3.0),
        "by_role_frac_of_3_mma_ceiling": {t: gemm_flops / (avg[t] * 1e-3) / 1e12 / (peaks["bf16_tflops"] / 3.0)
                                          for t in ("fwd", "dgrad", "wgrad") if t in avg},
        "kernel_algorithmic_bytes": blur_fwd_bytes(h_tot, CHANNELS, CHANNELS),
        "kernel_algorithmic_gbs": blur_fwd_bytes(h_tot, CHANNELS, CHANNELS) / (gemm_ms * 1e-3) / 1e9,
        "contraction_share_of_step": all_gemm_ms / step_ms,
        "step_minus_big_five_ms": step_ms - big_five_ms,
        "whole_step_algorithmic_gbs": (fb + bb) / (step_ms * 1e-3) / 1e9,
        "whole_step_frac_of_hbm_peak": (fb + bb) / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
        "hbm_bound_kernels": hbm_kernels,
        "note": "dense fp32-accurate contraction (AI ~205 FLOP/B) -> tensor-bound by the roofline, not HBM-bound; engine 5 is "
                "paced by the SM's shared-memory port (staged rows -> operand tiles -> tensor core), see DESIGN.md 3.2c",
    }

    cpu = cpu_baseline_leg()
    lattice = lattice_leg(dev)
    corr = corr_leg(dev)
    model_fwd = model_leg(dev)

    line = {
        "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BilateralConvFlex(3,1,64,[64]) fwd+bwd, 8192-pt frustum clouds, scale 1.0, "
                               "%d distinct clouds per GPU per step concatenated (reference is B=1)" % B,
                   "clouds_per_gpu_per_step": B, "points_per_step_per_gpu": n_tot, "vertices_per_step_per_gpu": h_tot,
                   "l2_policy": "inputs larger than L2 (working set %.0f MB per step)" % (
                       4e-6 * (2 * n_tot * CHANNELS + 2 * h_tot * CHANNELS)),
                   "index_dtype": "int64 (reference format)",
                   "launch": ("one CUDA-graph replay per step (forward + backward captured through the module API)"
                              if graph_ms is not None else "eager (graph capture failed: %s)" % graph_err),
                   "tile_plan": {"built_once_ms": plan_ms, "relaxation_sweeps": plan.sweeps, "tiles": plan.n_tiles,
                                 "mean_distinct_rows_per_tile": plan.sum_uniq / max(plan.n_tiles, 1),
                                 "max_distinct_rows_per_tile": plan.max_uniq, "usable": plan.usable,
                                 "note": "per-lattice precomputation from blur_neighbors alone, outside the timed region"},
                   "weight_images": "rebuilt on every call in the timed loops (cache disabled)",
                   "numa_node_rank0": numa_node},
        "eager": {"value": sharding.job_throughput(B * args.steps, world, eager_ms * 1e-3), "ms_per_step": eager_ms / args.steps,
                  "note": "same step launched kernel by kernel from Python"},
        "latency_1cloud_ms": extra.get("latency_1cloud", {}).get("ms_graph"),
        "latency_1cloud": extra.get("latency_1cloud"), "cfg2_scale3": extra.get("cfg2_scale3"),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "aggregate_h2d_gbs": world * h2d * e2e_steps / e2e_s / 1e9,
                "h2d_copy_ceiling_gbs_per_rank_all_ranks_busy": -h2d_gbs_min,
                "int32_tables": {"value": sharding.job_throughput(B * e2e_steps, world, e2e32_s), "unit": "clouds/s",
                                 "h2d_bytes_per_step": h2d32,
                                 "note": "same loop, index tables uploaded in the native int32 format"}},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "clocks": clocks, "lattice_build": lattice, "correlation": corr, "model_forward": model_fwd,
        "train_step": train,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


TRAIN_PAIRS_PER_STEP = 8


def train_leg(dev, rank, world, steps=4, warmup=2):
    """Secondary number (BASELINE configs[4]): one data-parallel optimisation step of the full
    HPLFlowNet -- 8 synthetic 8192+8192-pt pairs per step in total, sharded over the ranks (strong scaling,
    8/N pairs per GPU), lattice built on the GPU per pair, EPE3D loss, backward, ONE flat NCCL all-reduce of
    the 19.3 M gradients, Adam.  Called by every rank; timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from hplflownet_b200 import sharding, train as T
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from hplflownet_b200.synthetic import frustum_pair
    from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1

    class A:
        dim = 3
        evaluate = False
        use_leaky = bcn_use_bias = bcn_use_norm = True
        last_relu = False
        DEVICE = "cuda"
        scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1],
                             [.125, 1, 1, 1], [.0625, 1, 1, 1]]
    if TRAIN_PAIRS_PER_STEP % world:
        return {"skipped": "world size %d does not divide %d pairs" % (world, TRAIN_PAIRS_PER_STEP)}
    ok, err, model, opt, gen, pairs = 1, None, None, None, None, None
    try:
        torch.manual_seed(0)                                     # same initial weights on every rank
        model = HPLFlowNet(A()).to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)     # main.py:118, configs/train_ours.yaml
        gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
        local = TRAIN_PAIRS_PER_STEP // world
        pairs = []
        for i in range(local):
            pc1, pc2 = frustum_pair(N_POINTS, 500 + rank * local + i)
            pairs.append((pc1, pc2, (pc2 - pc1).astype("float32")))
        # dry run of the local part (no collective) so that a failure cannot leave the other ranks waiting
        p1, p2, sf, gd = gen(list(pairs[0]))
        T.epe3d_loss(model(p1[None], p2[None], collate_batch1(gd)), sf[None]).backward()
        torch.cuda.synchronize()
    except Exception as e:                                       # noqa: BLE001 -- secondary leg must not kill the bench
        ok, err = 0, "%s: %s" % (type(e).__name__, e)
    if world > 1:
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item()) if ok else 0
    if not ok:
        return {"error": err or "failed on another rank"}
    model.zero_grad(set_to_none=True)
    buckets = T.GradBuckets(model.parameters(), bucket_mb=16.0)      # overlapped, copy-free gradient averaging
    for _ in range(warmup):
        T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = T.train_step(model, opt, gen, pairs, collate_batch1, buckets)
    e1.record()
    torch.cuda.synchronize()
    (ms,) = sharding.max_over_ranks([e0.elapsed_time(e1)], device=dev)
    return {"workload": "HPLFlowNet train step: %d pairs (8192+8192 pts, 7 scales) per step over %d GPU(s), GPU lattice "
                        "build + forward + EPE3D + backward + bucketed gradient all-reduce overlapped with the last backward + Adam" % (TRAIN_PAIRS_PER_STEP, world),
            "gradient_buckets": len(buckets.buckets),
            "value": TRAIN_PAIRS_PER_STEP * steps / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms / steps,
            "scaling": "strong", "steps": steps, "warmup": warmup, "loss_finite": bool(torch.isfinite(loss)),
            "params": sum(p.numel() for p in model.parameters())}


def model_leg(dev):
    """Secondary number (BASELINE configs[3]): full HPLFlowNet forward, random weights, one synthetic
    FlyingThings3D-shaped 8192-point pair, evaluate mode; with and without the GPU lattice build."""
    import torch
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from hplflownet_b200.synthetic import frustum_pair
    from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1

    class A:
        dim = 3
        evaluate = True
        use_leaky = bcn_use_bias = bcn_use_norm = True
        last_relu = False
        DEVICE = "cuda"
        scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1],
                             [.125, 1, 1, 1], [.0625, 1, 1, 1]]
    torch.manual_seed(0)
    model = HPLFlowNet(A()).to(dev).eval()
    gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
    pc1, pc2 = frustum_pair(N_POINTS, 7)
    a, b = torch.from_numpy(pc1.T.copy()).to(dev), torch.from_numpy(pc2.T.copy()).to(dev)

    from hplflownet_b200 import ops

    def run(build):
        with torch.no_grad():
            gd = collate_batch1(gen.build(a, b)) if build else run.gd
            return model(a[None], b[None], gd)
    run.gd = collate_batch1(gen.build(a, b))
    out = {}
    with ops.weight_cache_scope():                           # evaluation: the weights are constant across calls
        for name, build in (("forward_ms", False), ("build_plus_forward_ms", True)):
            for _ in range(3):
                run(build)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 8
            for _ in range(reps):
                y = run(build)
            torch.cuda.synchronize()
            out[name] = 1e3 * (time.perf_counter() - t0) / reps
    # the same forward (resident lattice) as ONE CUDA-graph replay: what is left is kernel time
    try:
        from hplflownet_b200.graphs import GraphedStep
        with ops.weight_cache_scope():
            g = GraphedStep(lambda: run(False))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                g.replay()
            torch.cuda.synchronize()
            out["forward_graph_ms"] = 1e3 * (time.perf_counter() - t0) / 20
    except Exception as e:                                       # noqa: BLE001
        out["forward_graph_ms"] = None
        out["forward_graph_error"] = "%s: %s" % (type(e).__name__, e)
    out.update({"workload": "HPLFlowNet forward, 8192+8192-pt pair, 7 scales, evaluate mode, random weights",
                "pairs_per_s": 1e3 / out["build_plus_forward_ms"], "output_finite": bool(torch.isfinite(y).all()),
                "note": "reference CPU forward: 13.8 s/pair + 4.1-4.9 s lattice build (SURVEY §6, 8 vCPU)"})
    return out


def corr_leg(dev):
    """Secondary number (BASELINE configs[2]): BilateralCorrelationFlex(3, 1, 1, 64, [32, 32], [64, 64], prev_corr_dim=64)
    forward + backward on two 8192-point clouds at scale 1.0, tables from the GPU lattice builder."""
    import torch
    import hplflownet_b200 as hpl
    from hplflownet_b200.synthetic import frustum_pair
    from hplflownet_b200.transforms import GenerateDataUnsymmetric

    class A:
        dim = 3
        scales_filter_map = [[SCALE, 1, 1, 1]]
    gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
    pc1, pc2 = frustum_pair(N_POINTS, 11)
    a, b = torch.from_numpy(pc1.T.copy()).to(dev), torch.from_numpy(pc2.T.copy()).to(dev)
    d = gen.build(a, b)[0]
    h1, h2 = int(d["pc1_hash_cnt"]), int(d["pc2_hash_cnt"])
    torch.manual_seed(0)
    mod = hpl.BilateralCorrelationFlex(3, 1, 1, CHANNELS, [32, 32], [64, 64], "cuda", use_bias=True, use_leaky=True,
                                       use_norm=True, prev_corr_dim=CHANNELS, last_relu=False, chunk_size=-1).to(dev)
    f1 = torch.randn(1, CHANNELS, h1, device=dev, requires_grad=True)
    f2 = torch.randn(1, CHANNELS, h2, device=dev, requires_grad=True)
    prev = torch.randn(1, CHANNELS, N_POINTS, device=dev, requires_grad=True)
    args = (d["pc1_barycentric"][None], d["pc1_lattice_offset"][None], d["pc1_corr_indices"][None],
            d["pc2_corr_indices"][None], h1, h2)
    gy = None

    def step():
        nonlocal gy
        for t in (f1, f2, prev, *mod.parameters()):
            t.grad = None
        y = mod(f1, f2, prev, *args)
        if gy is None:
            gy = torch.randn_like(y)
        y.backward(gy)
    from hplflownet_b200 import ops
    from hplflownet_b200.graphs import GraphedStep
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"workload": "BilateralCorrelationFlex(3,1,1,64,[32,32],[64,64],prev_corr_dim=64) fwd+bwd, 8192+8192 pts, scale 1.0",
           "ms_per_pair_eager": ms, "H1": h1, "H2": h2,
           "note": "the reference materialises 172.8 KB per vertex for this layer (bnn_flow.py:189-199)"}
    # per-kernel durations of the two patch-correlation kernels (CUDA events, separate pass) and their roofline: both move
    # F * P * width * 4 bytes per vertex between L2 and the SMs (the factored first layer's (H, P * width) operands stay in L2)
    ops.PROFILE_GEMM, ops.PROFILE_ROWS = [], True
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ev = {}
    for tag, a, b in ops.PROFILE_GEMM:
        ev.setdefault(tag, []).append(a.elapsed_time(b))
    ops.PROFILE_GEMM, ops.PROFILE_ROWS = None, False
    peaks = measured_peaks()
    l2_bytes = 15.0 * 15.0 * 32 * 4 * h1 * 2                      # gathers from t1 (P rows) and t2 (F * P rows) per vertex, width 32
    roof = {}
    for tag in ("corr_gather", "corr_scatter"):
        if tag in ev:
            t = sum(ev[tag]) / len(ev[tag])
            roof[tag] = {"kernel_ms": t, "l2_to_sm_bytes": l2_bytes, "achieved_gbs": l2_bytes / (t * 1e-3) / 1e9,
                         "peak_gbs": peaks["hbm_gbs"], "frac": l2_bytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "bound": "hbm (L2-resident gather; the denominator is the measured HBM copy peak)"}
    out["roofline"] = roof
    try:
        g = GraphedStep(step)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out["ms_per_pair"] = e0.elapsed_time(e1) / 50
        out["launch"] = "one CUDA-graph replay per pair"
    except Exception as e:                                       # noqa: BLE001
        out["ms_per_pair"], out["launch"] = ms, "eager (graph capture failed: %s: %s)" % (type(e).__name__, e)
    out["pairs_per_s"] = 1e3 / out["ms_per_pair"]
    return out


def lattice_leg(dev):
    """Secondary number: the index half (GenerateDataUnsymmetric, 7-scale hierarchy of the
    reference's configs) on the GPU vs the oracle's C restatement on one host core."""
    import torch
    from hplflownet_b200.synthetic import frustum_pair
    from hplflownet_b200.transforms import GenerateDataUnsymmetric
    from oracle import lattice as OL

    class A:
        dim = 3
        scales_filter_map = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1],
                             [.125, 1, 1, 1], [.0625, 1, 1, 1]]
    gen = GenerateDataUnsymmetric(A(), device=dev, index_dtype=torch.int32)
    pairs = [frustum_pair(N_POINTS, 100 + s) for s in range(4)]
    dev_pairs = [(torch.from_numpy(a.T.copy()).to(dev), torch.from_numpy(b.T.copy()).to(dev)) for a, b in pairs]
    for a, b in dev_pairs[:2]:
        gen.build(a, b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 12
    for i in range(reps):
        a, b = dev_pairs[i % len(dev_pairs)]
        gen.build(a, b)
    torch.cuda.synchronize()
    gpu = reps / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for a, b in pairs[:2]:
        OL.generate(a, b, A.scales_filter_map)
    cpu = 2 / (time.perf_counter() - t0)
    return {"workload": "7-scale lattice hierarchy of one 8192+8192-pt pair (configs/test_ours_FlyingThings3D.yaml)",
            "value": gpu, "unit": "pairs/s", "cpu_port_value": cpu, "cpu_port_cores": 1,
            "note": "wall clock incl. one host sync per scale; reference Numba/khash build: 4.1-4.9 s/pair (SURVEY §6)"}


def cpu_baseline_leg():
    """Oracle (port of the reference CPU path) timed on the host cores, bounded sample."""
    import torch
    from oracle import bcl as OB
    from oracle import lattice as OL
    from hplflownet_b200.synthetic import frustum_pair
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    mod = make_state()
    state = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
    n_clouds = 6
    data = []
    for s in range(n_clouds):
        pc1, pc2 = frustum_pair(N_POINTS, s)
        d = OL.generate(pc1, pc2, [[SCALE, 1, -1, -1]])[0]
        torch.manual_seed(s)
        data.append((torch.randn(1, CHANNELS, N_POINTS), torch.from_numpy(d["pc1_barycentric"])[None],
                     torch.from_numpy(d["pc1_lattice_offset"])[None],
                     torch.from_numpy(d["pc1_blur_neighbors"])[None], torch.randn(1, CHANNELS, N_POINTS)))

    def one(i):
        feat, bary, off, nbr, gy = data[i]
        f = feat.clone().requires_grad_(True)
        y = OB.bcl_forward(state, f, bary, off, nbr, bary, off, do_splat=True, do_slice=True, use_norm=True,
                           use_leaky=True, use_bias=True)
        y.backward(gy)

    one(0)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 10.0 or reps < n_clouds:            # ~10 s of CPU work
        one(reps % n_clouds)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": reps / dt, "unit": "clouds/s", "cores": cores, "kind": "port",
            "sample": "%d clouds fwd+bwd in %.1f s (oracle/bcl.py, torch CPU, %d threads)" % (reps, dt, cores)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clouds", type=int, default=32, help="distinct clouds per GPU per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
