"""Pointwise Conv1d (+ (Leaky)ReLU) stacks on the same tensor-core GEMM as the lattice convolution.

``models/HPLFlowNet.py:21-24,234-236,239-240,426-428`` run ``Conv1dReLU`` / ``nn.Conv1d`` with kernel size 1 over
(1, C, N) point or vertex features: a dense GEMM per layer.  The stock op costs ~0.9 ms of the 8192-point
forward in cuDNN's fp32 SGEMM; here a whole stack (conv -> act -> conv -> ...) is one pass through
``_stack`` (filter size 1, no neighbour table), vertex-major in between, with a hand-written backward.
Parameters stay in the reference's modules (same ``state_dict``); only the arithmetic moves.
"""
import torch
import torch.nn as nn

from . import _stack, ops
from .bilateralNN import _act_code, conv_weight_grad, kernel_weight
from .module_utils import Conv1dReLU


class _PointwiseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, x, *params):
        c_in = x.size(1)
        x_amax = ops.amax_slots(x.device, 1) if (ops.fused_stats() and c_in % 4 == 0) else None
        xr = ops.cm_to_rows(x[0].contiguous(), amax=x_amax)
        n = xr.size(0)
        layers = [(kernel_weight(params[2 * l]), params[2 * l + 1].detach(), acts[l]) for l in range(len(acts))]
        xs, chans, out_cm, amaxs = _stack.forward(xr, c_in, n, layers, None, last_channel_major=True, x_amax=x_amax)
        out = out_cm if out_cm is not None else ops.rows_to_cm(xs[-1], chans[-1])
        ctx.xs, ctx.chans, ctx.layers, ctx.n = xs, chans, layers, n
        ctx.amaxs = amaxs
        ctx.shapes = [p.shape for p in params]
        return out.unsqueeze(0)

    @staticmethod
    def backward(ctx, grad_out):
        dy = ops.cm_to_rows(grad_out[0].contiguous())
        n_layers = len(ctx.layers)
        need_p = [ctx.needs_input_grad[2 + 2 * l] or ctx.needs_input_grad[3 + 2 * l] for l in range(n_layers)]
        dx, pg = _stack.backward(dy, ctx.xs, ctx.chans, ctx.layers, ctx.n, None, None, ctx.needs_input_grad[1], need_p,
                                 amaxs=ctx.amaxs)
        grads = []
        for l, g in enumerate(pg):
            grads += [None, None] if g is None else [conv_weight_grad(g[0], ctx.shapes[2 * l]), g[1]]
        d_x = ops.rows_to_cm(dx, ctx.chans[0]).unsqueeze(0) if dx is not None else None
        return (None, d_x, *grads)


def pointwise_stack(modules, x):
    """Apply a sequence of ``Conv1dReLU`` / ``nn.Conv1d`` (kernel size 1) modules to x (1, C, N) on the CUDA path."""
    params, acts = [], []
    for m in modules:
        has_act = isinstance(m, Conv1dReLU)
        conv = m.conv if has_act else m
        if not isinstance(conv, nn.Conv1d) or conv.kernel_size != (1,) or conv.bias is None:
            raise ValueError("pointwise_stack needs kernel-size-1 Conv1d layers with bias")
        params += [conv.weight, conv.bias]
        acts.append(_act_code(has_act, m.use_leaky if has_act else False))
    return _PointwiseFunction.apply(tuple(acts), x, *params)
