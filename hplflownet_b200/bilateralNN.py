"""B200-native ``BilateralConvFlex`` -- drop-in for models/bilateralNN.py:46-238.

Same constructor, forward signature, return shapes and ``state_dict`` layout as the reference
module, so ``models/HPLFlowNet.py`` runs on top of it unchanged.  The forward and the backward
are hand-written sm_100a kernels behind the C ABI (include/hplflownet_b200.h):

  splat   (:150-186)  scatter-add of barycentric-weighted features into vertex-major lattice
                      rows + density normalisation -- replaces two sparse-COO densifications;
  blur    (:198-221)  gather-GEMM over the neighbour table with fused bias + (Leaky)ReLU --
                      the F-times gathered copy (:215-217) never exists, so no chunking;
  slice   (:223-238)  barycentric gather of 4 lattice rows per point (+ bias).

``SparseSum`` / ``sparse_sum`` (:9-43) are kept as public names on top of the same scatter
kernel.  There is no CPU path: tensors must live on a CUDA device.
"""
import torch
import torch.nn as nn

from . import _stack, ops
from .module_utils import Conv2dReLU

__all__ = ["BilateralConvFlex", "SparseSum", "sparse_sum"]


def _act_code(has_act, use_leaky):
    if not has_act:
        return ops.ACT_NONE
    return ops.ACT_LEAKY if use_leaky else ops.ACT_RELU


class SparseSum(torch.autograd.Function):
    """Dense ``out[idx[m], :] += values[m, :]`` (models/bilateralNN.py:9-40).

    Kept for API parity; runs the splat scatter kernel with unit weights on remainder 0.
    ``indices`` (1, M) int64, ``values`` (M, C) -> (size[0], C).
    """

    @staticmethod
    def forward(ctx, indices, values, size, cuda=True):
        ctx.save_for_backward(indices)
        m, c = values.shape
        x = values.t().contiguous()                                        # (C, M)
        off = torch.full((4, m), -1, dtype=torch.int64, device=values.device)
        off[0] = indices.reshape(-1)                                       # remainders 1..3 unused
        ones = torch.ones((4, m), dtype=torch.float32, device=values.device)
        rows, _ = ops.scatter_rows(x, ones, off, int(size[0]), False)
        return rows[:, :c].contiguous()

    @staticmethod
    def backward(ctx, grad_output):
        indices, = ctx.saved_tensors
        grad_values = grad_output[indices.squeeze(0), :] if ctx.needs_input_grad[1] else None
        return None, grad_values, None, None


sparse_sum = SparseSum.apply


def conv2d_layers(module_list, use_leaky):
    """[(conv, act_code)] for a reference-layout Sequential of Conv2dReLU / bare nn.Conv2d."""
    out = []
    for layer in module_list:
        has_act = isinstance(layer, Conv2dReLU)
        out.append((layer.conv if has_act else layer, _act_code(has_act, use_leaky)))
    return out


def kernel_weight(w):
    """(Co, C, F, 1) conv weight -> its (F, C, Co) VIEW, the operand of the gather-GEMM (engine 2 reads it strided;
    ops.blur_gemm makes a contiguous copy for the engines that need one)."""
    return ops.with_owner(w.detach().reshape(w.size(0), w.size(1), -1).permute(2, 1, 0), w, "fwd")


def conv_weight_grad(dw, like):
    """(F, C, Co) -> the conv weight's own shape."""
    return dw.permute(2, 1, 0).reshape(tuple(like)).contiguous()


def _plan_for_first_layer(nbr2, c_in, c_out0):
    """Tile plan of the neighbour table when engine 5 can take the stack's first layer, else None."""
    if not ops.engine5_enabled() or not ops.conv5_supported(nbr2.size(0), c_in, c_out0):
        return None
    from . import plans
    plan = plans.plan_for(nbr2)
    return plan if (plan is not None and plan.usable) else None


class _BCLFunction(torch.autograd.Function):
    """splat -> conv stack -> slice with a hand-written backward (autograd in the reference)."""

    @staticmethod
    def forward(ctx, cfg, features, in_bary, in_off, nbr, out_bary, out_off, slice_bias, *params):
        do_splat, do_slice, use_norm, acts = cfg
        feat = features[0].contiguous()                      # (C, N) -- B = 1 (README.md:57)
        c_in = feat.size(0)
        nbr2 = nbr[0].contiguous()                           # (F, H)
        h = nbr2.size(1)
        layers = [(kernel_weight(params[2 * l]), params[2 * l + 1].detach(), acts[l]) for l in range(len(acts))]
        plan = _plan_for_first_layer(nbr2, c_in, layers[0][0].size(2))
        inv, first5, lat, wsum_amax = None, None, None, None
        # the producer of the lattice rows records max|rows| (or a bound of it) itself: no separate absmax pass
        fused_first = plan is not None and do_splat and use_norm     # (that path takes its slots from one zeroed arena)
        lat_amax = (ops.amax_slots(feat.device, 1)
                    if (ops.fused_stats() and (c_in % 4 == 0 or plan is not None) and not fused_first) else None)
        if do_splat:
            bary_i, off_i = in_bary[0].contiguous(), in_off[0].contiguous()
            if plan is not None and use_norm:
                # engine 5: the normalised rows only ever exist as their pre-split image.  |S[v]| <= max|feat| (a convex
                # combination, bilateralNN.py:150-186), so the splat kernel's fused max|feat| is a valid operand scale.
                # One pass turns the accumulators into the normalised image, writes 1 / (wsum + 1e-5) for the backward, records
                # max wsum and zeroes the accumulator again (it returns to the zero pool: no memset per call).
                if h * c_in * 4 >= ops.SIDE_ZERO_MIN_BYTES:
                    ops.conv5_prepare(layers[0][0])              # weight image on a side stream, under the splat
                from . import plans
                splan = plans.splat_plan_for(off_i, h)
                if splan is not None:
                    # the splat as a deterministic gather (plans.py: splat plan) fused with the split: the features are
                    # transposed to point-major rows once (recording max|feat|), then every lattice row is summed from its
                    # contributions in a fixed order and written straight into the image -- no atomics, no accumulator
                    z = ops.zero_arena(feat.device, [("x_amax", 1, torch.int32), ("w_amax", 1, torch.int32)])
                    lat_amax, wsum_amax = z["x_amax"], z["w_amax"]
                    feat_rows = ops.cm_to_rows(feat, amax=lat_amax)
                    inv = torch.empty(h, dtype=torch.float32, device=feat.device)
                    x16 = ops.h16b_splat_csr(feat_rows, c_in, bary_i, splan, lat_amax, normalize=True, inv_out=inv,
                                             norm_amax_out=wsum_amax)
                else:
                    raw = ops.zero_rows(h, c_in, feat.device)
                    z = ops.zero_arena(feat.device, [("wsum", h, torch.float32), ("x_amax", 1, torch.int32), ("w_amax", 1, torch.int32)])
                    lat_amax, wsum_amax = z["x_amax"], z["w_amax"]
                    raw, wsum = ops.scatter_rows(feat, bary_i, off_i, h, True, in_amax=lat_amax, rows=raw, wsum=z["wsum"])
                    inv = torch.empty_like(wsum)
                    x16 = ops.h16b_split_ex(raw, c_in, lat_amax, norm=wsum, inv_out=inv, norm_amax_out=wsum_amax,
                                            dispose=ops.dispose_mode(raw))
                    ops.recycle_rows(raw)
                first5 = _stack.First5(x16, lat_amax, plan)
            else:
                lat, wsum = ops.scatter_rows(feat, bary_i, off_i, h, use_norm)
                if use_norm:
                    inv = ops.normalize_rows_(lat, c_in, wsum, amax=lat_amax if c_in % 4 == 0 else None)
                else:
                    lat_amax = None
        else:
            bary_i = off_i = None
            lat = ops.cm_to_rows(feat, amax=lat_amax)
        if plan is not None and first5 is None:
            amax = lat_amax if lat_amax is not None else ops.absmax(lat)
            first5 = _stack.First5(ops.h16b_split(lat, c_in, amax), amax, plan)
            lat = None

        xs, chans, out_cm, amaxs = _stack.forward(lat, c_in, h, layers, nbr2, last_channel_major=not do_slice,
                                                  x_amax=lat_amax, first5=first5)

        if do_slice:
            bary_o, off_o = out_bary[0].contiguous(), out_off[0].contiguous()
            out = ops.gather_rows(xs[-1], chans[-1], bary_o, off_o, None,
                                  slice_bias.detach() if slice_bias is not None else None)
        else:
            bary_o = off_o = None
            out = out_cm if out_cm is not None else ops.rows_to_cm(xs[-1], chans[-1])

        ctx.cfg, ctx.chans, ctx.h = cfg, chans, h
        ctx.xs, ctx.layers, ctx.inv, ctx.first5 = xs, layers, inv, first5
        ctx.amaxs = amaxs
        # max of the weight sums, valid for the OUT tables when they are the in tables (slice backward bound, see backward)
        same_tables = (do_splat and do_slice and out_bary.data_ptr() == in_bary.data_ptr() and out_off.data_ptr() == in_off.data_ptr()
                       and out_bary.shape == in_bary.shape)
        ctx.wsum_amax = wsum_amax if same_tables else None
        ctx.idx = (bary_i, off_i, nbr2, bary_o, off_o)
        ctx.has_slice_bias = slice_bias is not None
        ctx.param_shapes = [p.shape for p in params]
        ops.join_side(feat.device)                               # (a large accumulator was zeroed on the side stream meanwhile)
        return out.unsqueeze(0)

    @staticmethod
    def backward(ctx, grad_out):
        do_splat, do_slice, use_norm, acts = ctx.cfg
        bary_i, off_i, nbr2, bary_o, off_o = ctx.idx
        chans, h, xs, layers = ctx.chans, ctx.h, ctx.xs, ctx.layers
        g = grad_out[0].contiguous()
        d_slice_bias = None
        dz_bound = None
        if do_slice:
            if ctx.first5 is not None and len(layers) == 1 and ctx.wsum_amax is not None:
                # |dz[v]| <= max|g| x (sum of barycentric weights at v): the splat's fused max|g| and the forward's max wsum
                # give the operand scale of dz without a pass over it; the accumulator comes from the zero pool
                w0 = layers[0][0]
                if (ctx.needs_input_grad[1] and h * chans[-1] * 4 >= ops.SIDE_ZERO_MIN_BYTES and ctx.first5.plan.symmetric
                        and ops.conv5_supported(w0.size(0), chans[1], chans[0])):
                    ops.conv5_prepare(w0.transpose(1, 2), mirror=True)   # the data gradient's weight image, under the splat of g
                arena = ops.zero_arena(g.device, [("g_amax", 1, torch.int32), ("dz_amax", 1, torch.int32),
                                                  ("dw", w0.numel(), torch.float32), ("db", chans[-1], torch.float32),
                                                  ("dsb", chans[-1], torch.float32)])
                from . import plans
                dgrad5_ok = (not ctx.needs_input_grad[1]) or (ctx.first5.plan.symmetric and ops.conv5_supported(w0.size(0), chans[1], chans[0]))
                splan = plans.splat_plan_for(off_o, h) if dgrad5_ok else None
                if splan is not None:
                    # slice backward as a gather over the out tables' splat plan, fused with act', the bias gradient and the
                    # operand split (see _stack._backward_single_layer5): dz never exists in fp32
                    g_rows = ops.cm_to_rows(g, amax=arena["g_amax"])
                    dx = None
                    arena["_dz16"] = lambda y, act, db, dz_amax: ops.h16b_splat_csr(
                        g_rows, chans[-1], bary_o, splan, arena["g_amax"], y=y, act=act, amax_b=ctx.wsum_amax, amax_out=dz_amax, colsum=db)
                else:
                    dx = ops.zero_rows(h, chans[-1], g.device)
                    dx, _ = ops.scatter_rows(g, bary_o, off_o, h, False, in_amax=arena["g_amax"], rows=dx)
                dz_bound = (arena["g_amax"], ctx.wsum_amax, arena)
                if ctx.has_slice_bias and ctx.needs_input_grad[7]:
                    d_slice_bias = arena["dsb"]
                    if g.numel() * 4 >= ops.SIDE_ZERO_MIN_BYTES and torch.cuda.is_current_stream_capturing():
                        # the slice bias gradient (a reduction over g, DRAM bound) runs on the side stream after the split
                        # pass, under the weight-gradient kernel (latency bound) -- see _stack._backward_single_layer5.  (Only
                        # while a graph is captured: launched eagerly, the fork costs more host time than it hides.)
                        arena["_after_split"] = lambda: ops.channel_sums(g, out=d_slice_bias, side=True)
                    else:
                        ops.channel_sums(g, out=d_slice_bias)
            else:
                dx, _ = ops.scatter_rows(g, bary_o, off_o, h, False)
                if ctx.has_slice_bias and ctx.needs_input_grad[7]:
                    d_slice_bias = ops.channel_sums(g)
        else:
            dx = ops.cm_to_rows(g)

        need_feat = ctx.needs_input_grad[1]
        need_param = [ctx.needs_input_grad[8 + 2 * l] or ctx.needs_input_grad[9 + 2 * l] for l in range(len(layers))]
        dx, pg = _stack.backward(dx, xs, chans, layers, h, nbr2, lambda: ops.transpose_table(nbr2, h),
                                 need_feat, need_param, amaxs=ctx.amaxs, first5=ctx.first5, dz_bound=dz_bound)
        grads = []
        for l, g_l in enumerate(pg):
            if g_l is None:
                grads += [None, None]
            else:
                grads += [conv_weight_grad(g_l[0], ctx.param_shapes[2 * l]), g_l[1]]

        d_feat = None
        if need_feat:
            if do_splat:
                d_feat = ops.gather_rows(dx, chans[0], bary_i, off_i, ctx.inv, None).unsqueeze(0)
            else:
                d_feat = ops.rows_to_cm(dx, chans[0]).unsqueeze(0)
        ops.join_side(g.device)
        return (None, d_feat, None, None, None, None, None, d_slice_bias, *grads)


class BilateralConvFlex(nn.Module):
    """Same interface as the reference module (models/bilateralNN.py:46-125)."""

    def __init__(self, d, neighborhood_size, num_input, num_output, DEVICE, use_bias, use_leaky, use_norm,
                 do_splat, do_slice, last_relu, chunk_size=1024 * 1024 * 25):
        super().__init__()
        self.d, self.d1 = d, d + 1
        self.neighborhood_size = neighborhood_size
        self.filter_size = self.get_filter_size()
        self.num_input, self.num_output = num_input, list(num_output)
        self.DEVICE = DEVICE
        self.use_bias, self.use_leaky, self.use_norm = use_bias, use_leaky, use_norm
        self.do_splat, self.do_slice, self.last_relu = do_splat, do_slice, last_relu
        self.MAX_SIZE = chunk_size      # accepted for compatibility; the fused path never chunks

        c_final = self.num_output[-1]
        # index buffers the reference registers (:90-92); unused here, kept for strict loading
        self.register_buffer("feat_indices", torch.arange(num_input, dtype=torch.long))
        if do_slice:
            self.register_buffer("out_indices", torch.arange(c_final, dtype=torch.long))

        layers, c_prev = [], num_input
        widths = self.num_output
        for i, c_out in enumerate(widths):
            ks = (self.filter_size, 1) if i == 0 else (1, 1)
            is_last = i == len(widths) - 1
            if is_last and not last_relu:
                layers.append(nn.Conv2d(c_prev, c_out, kernel_size=ks))
            else:
                layers.append(Conv2dReLU(c_prev, c_out, ks, use_leaky=use_leaky))
            c_prev = c_out
        self.blur_conv = nn.Sequential(*layers)

        if do_slice and use_bias:
            self.register_parameter("bias", nn.Parameter(torch.zeros(c_final, dtype=torch.float32)))

    def get_filter_size(self):
        return (self.neighborhood_size + 1) ** self.d1 - self.neighborhood_size ** self.d1

    def _layer_params(self):
        params, acts = [], []
        for conv, act in conv2d_layers(self.blur_conv, self.use_leaky):
            params += [conv.weight, conv.bias]
            acts.append(act)
        return params, tuple(acts)

    def forward(self, features, in_barycentric, in_lattice_offset, blur_neighbors, out_barycentric,
                out_lattice_offset):
        """features (1, C_in, N_in | H) -> (1, C_out, N_out | H); see models/bilateralNN.py:122-135."""
        if features.size(0) != 1:
            raise ValueError("batch size must be 1 (reference README.md:57); concatenate clouds instead")
        if not features.is_cuda:
            raise RuntimeError("BilateralConvFlex (B200) has no CPU path: inputs must be CUDA tensors")
        params, acts = self._layer_params()
        cfg = (self.do_splat, self.do_slice, self.use_norm, acts)
        bias = self.bias if (self.do_slice and self.use_bias) else None
        return _BCLFunction.apply(cfg, features, in_barycentric, in_lattice_offset, blur_neighbors,
                                  out_barycentric, out_lattice_offset, bias, *params)
