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
    gloo (CPU tensors, used by the tests).  Parameters without a gradient contribute zeros."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    if world_size == 1:
        return sum(p.numel() for p in params)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off


def train_step(model, optimizer, generator, pairs, collate):
    """One optimisation step over this rank's ``pairs`` = [(pc1 (N,3), pc2 (N,3), flow (N,3)), ...] numpy/torch.
    Returns the mean loss of the local pairs (float tensor on the device, not synchronised)."""
    optimizer.zero_grad(set_to_none=True)
    total = None
    for pc1, pc2, flow in pairs:
        p1, p2, sf, gd = generator([pc1, pc2, flow])
        out = model(p1[None], p2[None], collate(gd))
        loss = epe3d_loss(out, sf[None]) / len(pairs)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    allreduce_mean_grads_(list(model.parameters()))
    optimizer.step()
    from . import ops
    ops.invalidate_weight_cache()            # the step changed every weight (belt and braces next to the version check)
    return total
