"""Data-parallel training step for HPLFlowNet on the B200 layers (BASELINE configs[4]).

The reference wraps the model in ``torch.nn.DataParallel`` with ``batch_size: 1`` (main.py:104,
configs/train_ours.yaml:17), i.e. one replica works.  Here: one process per GPU, every rank runs its own
pairs (B = 1 API, lattice built on its GPU), gradients are averaged with ONE flat NCCL all-reduce
(19.3 M fp32 = 77 MB) and every rank applies the same Adam step.  No collective sits inside the BCL path.
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


def train_step(model, optimizer, generator, pairs, collate):
    """One optimisation step over this rank's ``pairs`` = [(pc1 (N,3), pc2 (N,3), flow (N,3)), ...] numpy/torch.
    Returns the mean loss of the local pairs (float tensor on the device, not synchronised)."""
    from . import ops
    optimizer.zero_grad(set_to_none=True)
    total = None
    with ops.weight_cache_scope():           # the weights are constant until optimizer.step(): one image per step
        for pc1, pc2, flow in pairs:
            p1, p2, sf, gd = generator([pc1, pc2, flow])
            out = model(p1[None], p2[None], collate(gd))
            loss = epe3d_loss(out, sf[None]) / len(pairs)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
    allreduce_mean_grads_(list(model.parameters()))
    optimizer.step()
    return total
