"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding helpers bench.py uses,
and the reference arm's rank handling under a torchrun-style environment."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hplflownet_b200 import sharding

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.cloud_ids(rank, world, 4)
    # every rank reports a different time; the job time is the slowest
    t_max, t2 = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    dist.barrier()
    with open(os.path.join(out_dir, "r%d.json" % rank), "w") as f:
        json.dump({"ids": ids, "t_max": t_max, "t2": t2, "all": gathered}, f)
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduction(tmp_path):
    world, port = 2, 29613
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / ("r%d.json" % r))) for r in range(world)]
    assert res[0]["ids"] == [0, 1, 2, 3] and res[1]["ids"] == [4, 5, 6, 7]
    flat = sorted(i for ids in res[0]["all"] for i in ids)
    assert flat == list(range(8))                       # disjoint cover, weak scaling
    for r in res:
        assert r["t_max"] == 2.0 and r["t2"] == 5.0     # max over ranks, on every rank
    assert sharding.job_throughput(4, 2, 2.0) == 4.0


def test_split_strong_covers_everything():
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 3, 8):
            parts = [sharding.split_strong(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    with pytest.raises(ValueError):
        sharding.cloud_ids(2, 2, 4)


def test_reference_arm_only_rank0_prints():
    # bench.py --impl reference under a 2-rank environment: rank 0 prints the JSON line, rank 1 exits 0 silently
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29614", OMP_NUM_THREADS="4")
    outs = []
    for rank in (1, 0):
        p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "1"], env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)),
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip())
    assert outs[0] == ""
    line = json.loads(outs[1].splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "clouds/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["value"] > 0


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hplflownet_b200.train import allreduce_mean_grads_
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)
    x = torch.full((2, 5), float(rank + 1))
    net(x).sum().backward()
    net[1].bias.grad = None                                  # a parameter without gradient on ANY rank: stays None
    if rank == 1:
        net[0].bias.grad = None                              # ... on one rank only: averaged with zeros
    n = allreduce_mean_grads_(list(net.parameters()) + [frozen])
    torch.save({"n": n, "grads": [None if p.grad is None else p.grad.clone() for p in net.parameters()]},
               os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_averages_over_ranks(tmp_path):
    world, port = 2, 29615
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / ("g%d.pt" % r)) for r in range(world)]
    assert res[0]["n"] == 5 * 7 + 7 + 7 * 3 + 3
    for a, b in zip(res[0]["grads"], res[1]["grads"]):
        assert (a is None and b is None) or torch.equal(a, b)    # identical on every rank after the all-reduce
    # expected: mean over ranks of the single-rank gradients (inputs 1 and 2 -> weight grads scale linearly)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    grads = []
    for r in range(world):
        net.zero_grad()
        net(torch.full((2, 5), float(r + 1))).sum().backward()
        grads.append([p.grad.clone() for p in net.parameters()])
    want = [(a + b) / 2 for a, b in zip(*grads)]
    want[1] = grads[0][1] / 2                                # rank 1 had no gradient for net[0].bias: mean with zeros
    assert res[0]["grads"][3] is None                        # dropped on both ranks: the optimizer must skip it (reference semantics)
    for k, (got, w) in enumerate(zip(res[0]["grads"], want)):
        if k != 3:
            assert torch.allclose(got, w, rtol=1e-6, atol=1e-7)


def _bucket_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hplflownet_b200.train import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(4))                   # takes part in no loss on any rank
    params = list(net.parameters()) + [unused]
    buckets = GradBuckets(params, bucket_mb=1e-4)                # ~26 floats per bucket: several buckets
    assert len(buckets.buckets) >= 3
    res = []
    for step in range(2):                                        # two steps: the views are reused
        buckets.zero_()
        for i in range(2):                                       # two local "pairs" per step, accumulated
            if i == 1:
                buckets.arm()
            x = torch.full((2, 5), float(rank + 1 + i + step))
            (net(x).sum() / 2).backward()
        buckets.finish()
        res.append([None if p.grad is None else p.grad.clone() for p in params])
    torch.save(res, os.path.join(out_dir, "b%d.pt" % rank))
    dist.destroy_process_group()


def test_bucketed_overlapped_gradient_allreduce(tmp_path):
    world, port = 2, 29617
    mp.spawn(_bucket_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / ("b%d.pt" % r)) for r in range(world)]
    for step in range(2):
        for a, b in zip(res[0][step], res[1][step]):
            assert (a is None and b is None) or torch.equal(a, b)
        assert res[0][step][-1] is None                          # the untouched parameter keeps grad = None
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
        want = [torch.zeros_like(p) for p in net.parameters()]
        for r in range(world):
            for i in range(2):
                net.zero_grad()
                (net(torch.full((2, 5), float(r + 1 + i + step))).sum() / 2).backward()
                for w, p in zip(want, net.parameters()):
                    w += p.grad / world
        for got, w in zip(res[0][step], want):
            assert torch.allclose(got, w, rtol=1e-6, atol=1e-7)
