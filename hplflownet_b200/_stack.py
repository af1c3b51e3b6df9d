"""A stack of learned lattice convolutions (gather-GEMM + bias + activation) and its backward.

Shared by BilateralConvFlex (blur_conv, models/bilateralNN.py:94-113) and
BilateralCorrelationFlex (corr_conv / blur_conv, models/bnn_flow.py:59-91).  Each layer is
``(w (F, C, Co) fp32, bias (Co) or None, act)``; only the first layer may carry a neighbour table.

Engine 5 (tile plans, csrc/gemm_plan.cu) takes the first, gathered layer when the caller hands in a
``First5`` record: the layer input then exists only as its pre-split h16b image (no fp32 copy), and forward,
data gradient and weight gradient all run on the same per-lattice plan.
"""
from . import ops


class First5:
    """First layer on engine 5: ``x16`` = h16b image of the layer input, ``amax`` its scale slot, ``plan`` the
    neighbour table's tile plan."""

    def __init__(self, x16, amax, plan):
        self.x16, self.amax, self.plan = x16, amax, plan


def forward(x, c_in, n_rows, layers, first_nbr=None, last_channel_major=False, first_row_scale=None, x_amax=None,
            first5=None):
    """Returns (xs, chans, out_cm, amaxs): xs[l] is the vertex-major input of layer l and xs[-1] the final
    vertex-major output -- unless the last layer is written channel-major directly (only when it has
    no activation), in which case it is returned as out_cm and not kept in xs.  amaxs[l] = (scale slot, pre-split
    image or None) of layer l's input, to be handed back to ``backward``.
    x_amax: max|x| statistic of the input if its producer already recorded it (ops.amax_slots).
    first5: run layer 0 on engine 5 (x may then be None)."""
    xs, chans, out_cm = [x], [c_in], None
    amaxs = []                     # max|input| of every layer (3xFP16 path), reused by the weight gradient
    # every layer's epilogue records max|output| = the next layer's operand scale (no absmax passes)
    slots = ops.amax_slots(layers[0][0].device, len(layers)) if (ops.fused_stats() and len(layers) > 1) else None
    next_amax = x_amax
    for l, (w, b, act) in enumerate(layers):
        last = l == len(layers) - 1
        out_amax = slots[l:l + 1] if (slots is not None and not last and w.size(2) % 4 == 0) else None
        if l == 0 and first5 is not None:
            amaxs.append((first5.amax, first5.x16))
            y = ops.conv5(first5.x16, first5.plan, c_in, w, b, act, first5.amax, out_amax=out_amax)
            next_amax = out_amax
            xs.append(y)
            chans.append(w.size(2))
            continue
        direct_cm = last and last_channel_major and act == ops.ACT_NONE
        scale = first_row_scale if l == 0 else None
        split = ops.DEFAULT_PRECISION >= 2 and scale is None and chans[-1] % 4 == 0
        amax = (next_amax if next_amax is not None else ops.absmax(xs[-1])) if split else None
        next_amax = None
        x16 = None
        if split and ops.DEFAULT_PRECISION == 4:
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
    return xs, chans, out_cm, amaxs


def _backward_single_layer5(dx, xs, chans, layer, first_nbr_t, need_input_grad, need_param_grad, amaxs, first5, dz_bound):
    """A one-layer stack on engine 5 whose dz magnitude is bounded by the product of two device scalars (slice backward:
    max|g| x max weight sum): ONE pass applies act', writes the operand image and the bias gradient and hands the
    accumulator back to the zero pool; then the weight and data gradients run on the plan."""
    w, b, act = layer
    plan = first5.plan
    x_amax, x16 = amaxs[0]
    dgrad5 = need_input_grad and plan.symmetric and ops.conv5_supported(w.size(0), chans[1], chans[0])
    keep_fp32 = need_input_grad and not dgrad5                     # engine 2 will gather fp32 rows
    arena = dz_bound[2] if len(dz_bound) > 2 else None                 # pre-zeroed small outputs (one fill for all of them)
    db = None
    if need_param_grad and b is not None:
        db = arena["db"] if arena is not None else ops.small_zeros(chans[1], dx.dtype, dx.device)
    dz_amax = arena["dz_amax"] if arena is not None else ops.amax_slots(dx.device, 1)
    if arena is not None and "_dz16" in arena:                   # (slice backward as a CSR gather: dx was never materialised)
        assert not keep_fp32
        dz16 = arena.pop("_dz16")(xs[1] if act != ops.ACT_NONE else None, act, db, dz_amax)
    else:
        dz16 = ops.h16b_split_ex(dx, chans[1], dz_bound[0], y=xs[1] if act != ops.ACT_NONE else None, act=act,
                                 amax_b=dz_bound[1], amax_out=dz_amax, colsum=db, dispose=1 if keep_fp32 else ops.dispose_mode(dx))
        if not keep_fp32:
            ops.recycle_rows(dx)    # (a large accumulator is zeroed on the side stream, under the two gradient kernels)
    if arena is not None and "_after_split" in arena:
        arena.pop("_after_split")()                               # (side-stream work the caller deferred to this point)
    grads = [None]
    if need_param_grad:
        grads[0] = (ops.wgrad5(x16, dz16, plan, chans[0], chans[1], x_amax, dz_amax, out=arena["dw"] if arena is not None else None), db)
    out = None
    if need_input_grad:
        wd = w.transpose(1, 2)                                    # (F, Co, C) view
        owner = getattr(w, "_hpl_owner", None)
        if owner is not None:
            ops.with_owner(wd, owner[0], "dgrad")
        if dgrad5:
            out = ops.conv5(dz16, plan, chans[1], wd, None, ops.ACT_NONE, dz_amax, mirror=True, tag="dgrad")
        else:
            out = ops.blur_gemm(dx, chans[1], first_nbr_t(), plan.n_in_rows, wd, None, ops.ACT_NONE, tag="dgrad", x_amax=dz_amax)
    return out, grads


def backward(dx, xs, chans, layers, n_rows, first_nbr, first_nbr_t, need_input_grad, need_param_grad,
             first_row_scale=None, amaxs=None, first5=None, dz_bound=None):
    """dx: gradient w.r.t. the stack's (post-activation) output, vertex-major, modified in place.
    first_nbr_t: callable returning the transposed table of the first layer (built lazily).
    dz_bound: (slot a, slot b[, arena]) with max|dx| <= a x b, and dx taken from ops.zero_rows (one-layer stacks on engine 5);
    arena: ops.zero_arena with "dz_amax", "dw", "db" pieces.
    Returns (dx_in or None, [(dw (F, C, Co), db (Co)) or None per layer])."""
    if dz_bound is not None and first5 is not None and len(layers) == 1 and amaxs:
        return _backward_single_layer5(dx, xs, chans, layers[0], first_nbr_t, need_input_grad, need_param_grad[0], amaxs,
                                       first5, dz_bound)
    grads = [None] * len(layers)
    for l in range(len(layers) - 1, -1, -1):
        w, b, act = layers[l]
        tbl = first_nbr if l == 0 else None
        on5 = l == 0 and first5 is not None
        split = ops.DEFAULT_PRECISION >= 2 and chans[l + 1] % 4 == 0
        db_fused = None
        if ops.fused_stats():
            # one pass: activation backward, max|dz| (operand scale of wgrad and dgrad) and the bias gradient
            dz_amax = ops.amax_slots(dx.device, 1) if (split or on5) else None
            if need_param_grad[l] and b is not None:
                db_fused = ops.small_zeros(chans[l + 1], dx.dtype, dx.device)
            if act != ops.ACT_NONE or dz_amax is not None or db_fused is not None:
                ops.act_backward_stats_(dx, xs[l + 1] if act != ops.ACT_NONE else None, chans[l + 1], act, dz_amax, db_fused)
        else:
            if act != ops.ACT_NONE:
                ops.act_backward_(dx, xs[l + 1], chans[l + 1], act)
            dz_amax = ops.absmax(dx) if (split or on5) else None             # shared by wgrad and dgrad
        x_amax, x16 = amaxs[l] if amaxs else (None, None)
        if on5:
            plan = first5.plan
            # the data gradient runs on the same plan with mirrored taps when the table is its own mirrored transpose and
            # the transposed shape (Co -> C) fits the kernel; otherwise engine 2 gathers through an explicit transpose
            dgrad5 = need_input_grad and plan.symmetric and ops.conv5_supported(w.size(0), chans[1], chans[0])
            dz16 = ops.h16b_split(dx, chans[1], dz_amax) if (need_param_grad[0] or dgrad5) else None
            if need_param_grad[0]:
                dw = ops.wgrad5(x16, dz16, plan, chans[0], chans[1], x_amax, dz_amax)
                db = db_fused
                if b is not None and db is None:
                    db = ops.small_zeros(chans[1], dx.dtype, dx.device)
                    ops.column_sums_(dx, chans[1], db)
                grads[0] = (dw, db)
            if need_input_grad:
                wd = w.transpose(1, 2)                                # (F, Co, C) view
                owner = getattr(w, "_hpl_owner", None)
                if owner is not None:
                    ops.with_owner(wd, owner[0], "dgrad")
                if dgrad5:                                            # same table, mirrored taps (plans.py)
                    dx = ops.conv5(dz16, plan, chans[1], wd, None, ops.ACT_NONE, dz_amax, mirror=True, tag="dgrad")
                else:
                    dx = ops.blur_gemm(dx, chans[1], first_nbr_t(), plan.n_in_rows, wd, None, ops.ACT_NONE, tag="dgrad",
                                       x_amax=dz_amax)
            else:
                dx = None
            continue
        dz16 = None
        if split and ops.DEFAULT_PRECISION == 4 and (l > 0 or need_input_grad):
            dz16 = ops.h16_split(dx, chans[l + 1], dz_amax)
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
