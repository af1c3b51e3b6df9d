"""Data-parallel training step for HPLFlowNet on the B200 layers (BASELINE configs[4]).

The reference wraps the model in ``torch.nn.DataParallel`` with ``batch_size: 1`` (main.py:104,
configs/train_ours.yaml:17), i.e. one replica works.  Here: one process per GPU, every rank runs its own
pairs (B = 1 API, lattice built on its GPU), gradients are averaged over the ranks and every rank applies the
same Adam step.  No collective sits inside the BCL path.

Gradient exchange: ``GradBuckets`` keeps every ``p.grad`` as a view into a few flat bucket buffers (no flatten /
unflatten copies) and, during the backward of the rank's LAST pair, all-reduces each bucket asynchronously as soon as
all of its parameters have their final gradient -- the up-path layers ``bcn1_`` .. ``bcn3_`` hold 74 % of the
parameters and finish their backward first, so most of the 77 MB travels while the rest of the backward still runs.
``allreduce_mean_grads_`` (one flat all-reduce after the backward) remains for callers without buckets.
"""
import torch
import torch.distributed as dist


def epe3d_loss(pred, target):
    """models/epe3d_loss.py:9 followed by the .mean() of main.py:213."""
    return torch.norm(pred - target, p=2, dim=1).mean()


def allreduce_mean_grads_(params, world_size=None):
    """Average ``p.grad`` over all ranks with a single flat all-reduce (in place).  Works with NCCL (GPU) and
    gloo (CPU tensors, used by the tests).  A parameter that has no gradient on ANY rank keeps ``grad = None`` (the
    optimizer skips it, as in the reference where unused / frozen parameters are never touched); one that has a
    gradient on some ranks only gets the mean with zeros from the others.  A presence mask travels with the
    gradients in the same all-reduce."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world_size == 1:
        return sum(p.numel() for p in params if p.grad is not None)
    like = next((p.grad for p in params if p.grad is not None), params[0])
    present = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], dtype=like.dtype, device=like.device)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] + [present])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    seen = flat[-len(params):].tolist()
    flat.div_(world_size)
    off = 0
    for p, cnt in zip(params, seen):
        n = p.numel()
        if cnt > 0:
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        else:
            p.grad = None
        off += n
    return off


class GradBuckets:
    """Bucketed, overlapped gradient averaging.

    ``buckets = GradBuckets(model.parameters(), bucket_mb=16)`` once; per optimisation step
    ``buckets.zero_()`` -> backward of all pairs but the last -> ``buckets.arm()`` -> backward of the last pair (its
    post-accumulate hooks launch the all-reduces) -> ``buckets.finish()`` -> ``optimizer.step()``.
    Parameters are bucketed in reverse registration order (roughly the order in which autograd finishes them).
    A parameter whose hook never fired on ANY rank ends the step with ``grad = None`` (the optimizer skips it, as in
    the reference); one that fired on some ranks only is averaged with zeros from the others."""

    def __init__(self, params, bucket_mb=16.0, world_size=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size if world_size is not None else (
            dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1)
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets, self.slot = [], {}
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur)
        self.armed = False
        self.handles = []
        self.fired = torch.zeros(len(self.params), dtype=torch.float32)
        self._index = {id(p): i for i, p in enumerate(self.params)}
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _close(self, plist):
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        b = len(self.buckets)
        off = 0
        views = []
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            self.slot[id(p)] = b
            off += p.numel()
        self.buckets.append({"flat": flat, "params": plist, "views": views, "pending": 0})

    def zero_(self):
        """Start of a step: zero the buckets and point every ``p.grad`` at its view."""
        for b in self.buckets:
            b["flat"].zero_()
            for p, v in zip(b["params"], b["views"]):
                p.grad = v
        self.fired.zero_()
        self.armed, self.handles = False, []

    def arm(self):
        """Before the backward of the LAST local pair: from now on a finished bucket is all-reduced at once."""
        self.armed = True
        for b in self.buckets:
            b["pending"] = len(b["params"])

    def _hook(self, p):
        self.fired[self._index[id(p)]] = 1.0
        if not self.armed:
            return
        b = self.buckets[self.slot[id(p)]]
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            self.handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        """After the last backward: launch what is left (buckets with parameters that got no gradient this step), wait,
        average, and drop the gradients of parameters nobody touched."""
        if self.world > 1:
            for b in self.buckets:
                if b["pending"] > 0 or not self.armed:
                    self.handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, async_op=True))
            seen = self.fired.to(self.buckets[0]["flat"].device)
            self.handles.append(dist.all_reduce(seen, op=dist.ReduceOp.SUM, async_op=True))
            for h in self.handles:
                h.wait()
            for b in self.buckets:
                b["flat"].div_(self.world)
            seen = seen.cpu()
        else:
            seen = self.fired
        for p, cnt in zip(self.params, seen.tolist()):
            if cnt == 0:
                p.grad = None
        self.armed, self.handles = False, []


def train_step(model, optimizer, generator, pairs, collate, buckets=None):
    """One optimisation step over this rank's ``pairs`` = [(pc1 (N,3), pc2 (N,3), flow (N,3)), ...] numpy/torch.
    buckets: a ``GradBuckets`` over the model's parameters (overlapped, copy-free gradient averaging); None = one flat
    all-reduce after the backward.  Returns the mean loss of the local pairs (device tensor, not synchronised)."""
    from . import ops
    if buckets is not None:
        buckets.zero_()
    else:
        optimizer.zero_grad(set_to_none=True)
    total = None
    with ops.weight_cache_scope():           # the weights are constant until optimizer.step(): one image per step
        for i, (pc1, pc2, flow) in enumerate(pairs):
            p1, p2, sf, gd = generator([pc1, pc2, flow])
            out = model(p1[None], p2[None], collate(gd))
            loss = epe3d_loss(out, sf[None]) / len(pairs)
            if buckets is not None and i == len(pairs) - 1:
                buckets.arm()                # this backward's hooks all-reduce every bucket as it completes
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
    if buckets is not None:
        buckets.finish()
    else:
        allreduce_mean_grads_(list(model.parameters()))
    optimizer.step()
    return total
