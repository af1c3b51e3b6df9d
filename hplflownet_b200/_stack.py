"""A stack of learned lattice convolutions (gather-GEMM + bias + activation) and its backward.

Shared by BilateralConvFlex (blur_conv, models/bilateralNN.py:94-113) and
BilateralCorrelationFlex (corr_conv / blur_conv, models/bnn_flow.py:59-91).  Each layer is
``(w (F, C, Co) fp32, bias (Co) or None, act)``; only the first layer may carry a neighbour table.
"""
from . import ops


def forward(x, c_in, n_rows, layers, first_nbr=None, last_channel_major=False, first_row_scale=None, x_amax=None):
    """Returns (xs, chans, out_cm): xs[l] is the vertex-major input of layer l and xs[-1] the final
    vertex-major output -- unless the last layer is written channel-major directly (only when it has
    no activation), in which case it is returned as out_cm and not kept in xs.
    x_amax: max|x| statistic of the input if its producer already recorded it (ops.amax_slots)."""
    xs, chans, out_cm = [x], [c_in], None
    amaxs = []                     # max|input| of every layer (3xFP16 path), reused by the weight gradient
    # engine 2: every layer's epilogue records max|output| = the next layer's operand scale (no absmax passes)
    slots = ops.amax_slots(x.device, len(layers)) if (ops.fused_stats() and len(layers) > 1) else None
    next_amax = x_amax
    for l, (w, b, act) in enumerate(layers):
        last = l == len(layers) - 1
        direct_cm = last and last_channel_major and act == ops.ACT_NONE
        scale = first_row_scale if l == 0 else None
        split = ops.DEFAULT_PRECISION >= 2 and scale is None and chans[-1] % 4 == 0
        amax = (next_amax if next_amax is not None else ops.absmax(xs[-1])) if split else None
        next_amax = None
        out_amax = slots[l:l + 1] if (slots is not None and not last and w.size(2) % 4 == 0) else None
        x16 = None
        if split and ops.DEFAULT_PRECISION == 3:
            x16 = ops.split16(xs[-1], chans[-1], amax)
        elif split and ops.DEFAULT_PRECISION == 4:
            x16 = ops.h16_split(xs[-1], chans[-1], amax)
        amaxs.append((amax, x16))
        y = ops.blur_gemm(xs[-1], chans[-1], first_nbr if l == 0 else None, n_rows, w, b,
                          act, out_channel_major=direct_cm, row_scale=scale, x_amax=amax, x16=x16, out_amax=out_amax)
        next_amax = out_amax
        if direct_cm:
            out_cm = y
        else:
            xs.append(y)
        chans.append(w.size(2))
    forward.last_amaxs = amaxs
    return xs, chans, out_cm


def backward(dx, xs, chans, layers, n_rows, first_nbr, first_nbr_t, need_input_grad, need_param_grad,
             first_row_scale=None, amaxs=None):
    """dx: gradient w.r.t. the stack's (post-activation) output, vertex-major, modified in place.
    first_nbr_t: callable returning the transposed table of the first layer (built lazily).
    Returns (dx_in or None, [(dw (F, C, Co), db (Co)) or None per layer])."""
    grads = [None] * len(layers)
    for l in range(len(layers) - 1, -1, -1):
        w, b, act = layers[l]
        tbl = first_nbr if l == 0 else None
        split = ops.DEFAULT_PRECISION >= 2 and chans[l + 1] % 4 == 0
        db_fused = None
        if ops.fused_stats():
            # one pass: activation backward, max|dz| (operand scale of wgrad and dgrad) and the bias gradient
            dz_amax = ops.amax_slots(dx.device, 1) if split else None
            if need_param_grad[l] and b is not None:
                db_fused = dx.new_zeros(chans[l + 1])
            if act != ops.ACT_NONE or dz_amax is not None or db_fused is not None:
                ops.act_backward_stats_(dx, xs[l + 1] if act != ops.ACT_NONE else None, chans[l + 1], act, dz_amax, db_fused)
        else:
            if act != ops.ACT_NONE:
                ops.act_backward_(dx, xs[l + 1], chans[l + 1], act)
            dz_amax = ops.absmax(dx) if split else None                      # shared by wgrad and dgrad
        dz16 = None
        if split and ops.DEFAULT_PRECISION == 3:
            dz16 = ops.split16(dx, chans[l + 1], dz_amax)
        elif split and ops.DEFAULT_PRECISION == 4 and (l > 0 or need_input_grad):
            dz16 = ops.h16_split(dx, chans[l + 1], dz_amax)
        x_amax, x16 = amaxs[l] if amaxs else (None, None)
        if need_param_grad[l]:
            grads[l] = ops.blur_wgrad(xs[l], chans[l], tbl, n_rows, dx, chans[l + 1], w.size(0),
                                      want_db=b is not None and db_fused is None,
                                      row_scale=first_row_scale if l == 0 else None,
                                      x_amax=x_amax, dz_amax=dz_amax, x16=x16, dz16=dz16)
            if db_fused is not None:
                grads[l] = (grads[l][0], db_fused)
        if l > 0 or need_input_grad:
            wd = w.transpose(1, 2)                                    # (F, Co, C) view
            owner = getattr(w, "_hpl_owner", None)
            if owner is not None:
                ops.with_owner(wd, owner[0], "dgrad")
            tbl_t = first_nbr_t() if (l == 0 and first_nbr is not None) else None
            n_in = xs[l].size(0)
            dx = ops.blur_gemm(dx, chans[l + 1], tbl_t, n_in, wd, None, ops.ACT_NONE, tag="dgrad", x_amax=dz_amax, x16=dz16)
        else:
            dx = None
    return dx, grads
