"""B200-native ``BilateralCorrelationFlex`` -- drop-in for models/bnn_flow.py:10-210.

Same constructor, forward signature, output shape ``(1, num_output[-1], H1)`` and ``state_dict``
layout as the reference.  The computation is restructured (csrc/corr.cu): the reference gathers
``combined_input`` of shape (B, 2C+C', F, P, H1) (bnn_flow.py:189-199, 172.8 KB per vertex) and
runs ``Conv3d (1,P,1)`` on it; here the first, linear, conv layer is applied per *source* vertex
and patch slot by two dense GEMMs, after which the pre-activation of every (vertex, displacement)
is a gather-sum of P short vectors per cloud.  Same numbers (fp32, <= 1e-5 relative), ~16x fewer
FLOPs, no F*P*C intermediate, no chunking.
"""
import torch
import torch.nn as nn

from . import _lib, _stack, ops
from .bilateralNN import _act_code, conv2d_layers, conv_weight_grad, kernel_weight
from .module_utils import Conv2dReLU, Conv3dReLU

__all__ = ["BilateralCorrelationFlex"]


def _pad_cols(t, width):
    """Zero-pad the last dim of a 2-D/3-D tensor to ``width``."""
    if t.size(-1) == width:
        return t
    out = t.new_zeros(t.shape[:-1] + (width,))
    out[..., :t.size(-1)] = t
    return out


class _CorrFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, cfg, feat1, feat2, prev, bary1, off1, i1, i2, *params):
        use_norm, corr_acts, blur_acts = cfg
        n_corr, n_blur = len(corr_acts), len(blur_acts)
        dev = feat1.device
        f1, f2 = feat1[0].contiguous(), feat2[0].contiguous()          # (C, H1), (C, H2)
        c, h1 = f1.shape
        h2 = f2.size(1)
        i1c, _ = ops._idx(i1[0].contiguous(), "pc1_corr_indices")       # (P, H1)
        i2c, _ = ops._idx(i2[0].contiguous(), "pc2_corr_indices")       # (F, P, H1)
        if i1c.dtype != i2c.dtype:                                      # the kernels read both tables at one width
            i2c = i2c.to(i1c.dtype)
        patch, filt = i1c.size(0), i2c.size(0)
        if i1c.size(-1) != h1 or i2c.size(-1) != h1 or i2c.size(1) != patch:
            raise ValueError("correlation tables do not match feat1: pc1 %s, pc2 %s, H1 %d" % (tuple(i1c.shape), tuple(i2c.shape), h1))

        # ---- lattice-1 operand: [splat(prev) | feat1] (bnn_flow.py:119-168); lattice-2: feat2
        rows1 = ops.cm_to_rows(f1)                                      # (H1, ld)
        inv = bary = off = None
        c_prev = 0
        if prev is not None:
            c_prev = prev.size(1)
            bary = ops._f32(bary1[0].contiguous(), "barycentric1")
            off, _ = ops._idx(off1[0].contiguous(), "lattice_offset1")
            if bary.shape != (4, prev.size(-1)) or off.shape != bary.shape:
                raise ValueError("barycentric1 / lattice_offset1 must be (1, 4, N) like prev_corr_feat")
            prow, wsum = ops.scatter_rows(prev[0].contiguous(), bary, off, h1, use_norm)
            if use_norm:
                inv = ops.normalize_rows_(prow, c_prev, wsum)
            s1 = torch.zeros((h1, ops.round4(c_prev + c)), dtype=torch.float32, device=dev)
            s1[:, :c_prev] = prow[:, :c_prev]
            s1[:, c_prev:c_prev + c] = rows1[:, :c]
        else:
            s1 = rows1
        s2 = ops.cm_to_rows(f2)
        c1 = c_prev + c

        # ---- first corr layer, factored:  T = S @ W[:, :, p]  then gather-sum over the patch
        w0, b0 = params[0], params[1]                                   # (O, C1 + C, 1, P, 1)
        o1 = w0.size(0)
        wp = ops.round4(o1)
        w0m = w0.detach()[:, :, 0, :, 0]                                # (O, C1 + C, P)
        wa = _pad_cols(w0m[:, :c1].permute(1, 2, 0), wp).reshape(1, c1, patch * wp).contiguous()
        wb = _pad_cols(w0m[:, c1:].permute(1, 2, 0), wp).reshape(1, c, patch * wp).contiguous()
        t1 = ops.blur_gemm(s1, c1, None, h1, wa, None, ops.ACT_NONE)    # (H1, P*wp)
        t2 = ops.blur_gemm(s2, c, None, h2, wb, None, ops.ACT_NONE)     # (H2, P*wp)
        b0p = _pad_cols(b0.detach()[None], wp)[0].contiguous()
        z = torch.empty((h1 * filt, wp), dtype=torch.float32, device=dev)
        with ops._timed("corr_gather"):
            _lib.call("hpl_corr_gather", t1.data_ptr(), t1.stride(0), i1c.data_ptr(), t2.data_ptr(), t2.stride(0),
                      i2c.data_ptr(), int(i1c.dtype == torch.int64), b0p.data_ptr(), corr_acts[0], z.data_ptr(),
                      z.stride(0), wp, patch, filt, h1, h1, h2, ops._stream())
        del t1, t2

        # ---- remaining 1x1x1 corr layers on (H1*F, O) rows
        corr_layers = [(kernel_weight(params[2 * l]), params[2 * l + 1].detach(), corr_acts[l])
                       for l in range(1, n_corr)]
        zs, zch, _, z_amaxs = _stack.forward(z, o1, h1 * filt, corr_layers)
        o_last = zch[-1]
        ldz = zs[-1].stride(0)

        # ---- displacement filter (bnn_flow.py:205): (H1, F*ldz) @ (F*ldz, Q), then 1x1 layers
        base = 2 * n_corr
        wq = params[base].detach()[:, :, :, 0]                          # (Q, O_last, F)
        wq_k = _pad_cols(wq.permute(2, 0, 1), ldz).permute(0, 2, 1).reshape(1, filt * ldz, wq.size(0)).contiguous()
        blur_layers = [(wq_k, params[base + 1].detach(), blur_acts[0])]
        blur_layers += [(kernel_weight(params[base + 2 * l]), params[base + 2 * l + 1].detach(), blur_acts[l])
                        for l in range(1, n_blur)]
        zf = zs[-1].view(h1, filt * ldz)
        ys, ych, out_cm, y_amaxs = _stack.forward(zf, filt * ldz, h1, blur_layers, last_channel_major=True)
        out = out_cm if out_cm is not None else ops.rows_to_cm(ys[-1], ych[-1])

        ctx.cfg, ctx.dims = cfg, (c, c_prev, c1, h1, h2, patch, filt, o1, wp, o_last, ldz)
        ctx.saved = (s1, s2, wa, wb, z, zs, zch, corr_layers, ys, ych, blur_layers, inv, bary, off, i1c, i2c)
        ctx.amaxs = (z_amaxs, y_amaxs)
        ctx.param_shapes = [p.shape for p in params]
        return out.unsqueeze(0)

    @staticmethod
    def backward(ctx, grad_out):
        use_norm, corr_acts, blur_acts = ctx.cfg
        c, c_prev, c1, h1, h2, patch, filt, o1, wp, o_last, ldz = ctx.dims
        s1, s2, wa, wb, z, zs, zch, corr_layers, ys, ych, blur_layers, inv, bary, off, i1c, i2c = ctx.saved
        n_corr, n_blur = len(corr_acts), len(blur_acts)
        nig = ctx.needs_input_grad
        dev = grad_out.device
        need = lambda k: nig[8 + k]
        grads = [None] * (2 * (n_corr + n_blur))
        base = 2 * n_corr

        # ---- displacement filter stack
        dy = ops.cm_to_rows(grad_out[0].contiguous())
        need_p = [need(base + 2 * l) or need(base + 2 * l + 1) for l in range(n_blur)]
        dzf, pg = _stack.backward(dy, ys, ych, blur_layers, h1, None, None, True, need_p, amaxs=ctx.amaxs[1])
        for l, g in enumerate(pg):
            if g is None:
                continue
            dw, db = g
            if l == 0:     # (1, F*ldz, Q) -> (Q, O_last, F, 1)
                dwq = dw.view(filt, ldz, -1)[:, :o_last].permute(2, 1, 0).unsqueeze(-1).contiguous()
                grads[base], grads[base + 1] = dwq, db
            else:
                grads[base + 2 * l] = conv_weight_grad(dw, ctx.param_shapes[base + 2 * l])
                grads[base + 2 * l + 1] = db

        # ---- 1x1x1 corr layers
        dz = dzf.view(h1 * filt, ldz)
        need_p = [need(2 * l) or need(2 * l + 1) for l in range(1, n_corr)]
        dz, pg = _stack.backward(dz, zs, zch, corr_layers, h1 * filt, None, None, True, need_p, amaxs=ctx.amaxs[0])
        for k, g in enumerate(pg):
            if g is not None:
                l = k + 1
                grads[2 * l] = conv_weight_grad(g[0], ctx.param_shapes[2 * l])
                grads[2 * l + 1] = g[1]

        # ---- first (factored) corr layer
        ops.act_backward_(dz, z, o1, corr_acts[0])
        dt1 = torch.zeros((h1, patch * wp), dtype=torch.float32, device=dev)
        dt2 = torch.zeros((h2, patch * wp), dtype=torch.float32, device=dev)
        with ops._timed("corr_scatter"):
            _lib.call("hpl_corr_scatter", dz.data_ptr(), dz.stride(0), i1c.data_ptr(), i2c.data_ptr(),
                      int(i1c.dtype == torch.int64), dt1.data_ptr(), dt1.stride(0), dt2.data_ptr(), dt2.stride(0),
                      wp, patch, filt, h1, h1, h2, ops._stream())
        if need(0) or need(1):
            dwa, _ = ops.blur_wgrad(s1, c1, None, h1, dt1, patch * wp, 1, want_db=False)     # (1, C1, P*wp)
            dwb, _ = ops.blur_wgrad(s2, c, None, h2, dt2, patch * wp, 1, want_db=False)      # (1, C,  P*wp)
            dwa = dwa.view(c1, patch, wp)[:, :, :o1]
            dwb = dwb.view(c, patch, wp)[:, :, :o1]
            dw0 = torch.cat((dwa, dwb), 0).permute(2, 0, 1)                                   # (O, C1 + C, P)
            grads[0] = dw0.reshape(ctx.param_shapes[0]).contiguous()
            # conv3d bias: every (f, v) output position sees it once
            db0 = ops.small_zeros(wp, torch.float32, dev)
            _lib.call("hpl_column_sums", dz.data_ptr(), dz.stride(0), dz.size(0), wp, db0.data_ptr(), ops._stream())
            grads[1] = db0[:o1].contiguous()

        d_f1 = d_f2 = d_prev = None
        if nig[1] or nig[3]:
            ds1 = ops.blur_gemm(dt1, patch * wp, None, h1, wa.transpose(1, 2).contiguous(), None, ops.ACT_NONE,
                                tag="dgrad")                                                  # (H1, ld(C1))
            if nig[1]:
                d_f1 = ops.rows_to_cm(ds1[:, c_prev:c_prev + c].contiguous() if c_prev else ds1, c).unsqueeze(0)
            if c_prev and nig[3]:
                dp = ds1[:, :c_prev]
                if dp.stride(0) % 4 != 0 or dp.data_ptr() % 16 != 0 or not dp.is_contiguous():
                    dp = _pad_cols(dp, ops.round4(c_prev)).contiguous()
                d_prev = ops.gather_rows(dp, c_prev, bary, off, inv, None).unsqueeze(0)
        if nig[2]:
            ds2 = ops.blur_gemm(dt2, patch * wp, None, h2, wb.transpose(1, 2).contiguous(), None, ops.ACT_NONE,
                                tag="dgrad")
            d_f2 = ops.rows_to_cm(ds2, c).unsqueeze(0)
        return (None, d_f1, d_f2, d_prev, None, None, None, None, *grads)


class BilateralCorrelationFlex(nn.Module):
    """Same interface as the reference module (models/bnn_flow.py:10-99)."""

    def __init__(self, d, corr_filter_radius, corr_corr_radius, num_input, num_corr_output, num_output, DEVICE,
                 use_bias, use_leaky, use_norm, prev_corr_dim, last_relu, chunk_size=1024 * 1024 * 25):
        super().__init__()
        self.d, self.d1 = d, d + 1
        self.corr_size = self.get_filter_size(corr_corr_radius)
        self.filter_size = self.get_filter_size(corr_filter_radius)
        self.num_input, self.prev_corr_dim = num_input, prev_corr_dim
        self.num_corr_output, self.num_output = list(num_corr_output), list(num_output)
        self.DEVICE = DEVICE
        self.use_leaky, self.use_norm, self.last_relu = use_leaky, use_norm, last_relu
        self.MAX_SIZE = chunk_size      # accepted for compatibility; nothing is chunked
        # use_bias is accepted and unused, exactly like the reference (bnn_flow.py:45)

        self.register_buffer("feat_indices", torch.arange(num_input, dtype=torch.long))
        if prev_corr_dim != 0:
            self.register_buffer("feat1_indices", torch.arange(num_input + prev_corr_dim, dtype=torch.long))
        else:
            self.feat1_indices = self.feat_indices
        self.register_buffer("out_indices", torch.arange(self.num_output[-1], dtype=torch.long))

        layers, c_prev = [], num_input * 2 + prev_corr_dim
        for i, c_out in enumerate(self.num_corr_output):
            ks = (1, self.corr_size, 1) if i == 0 else (1, 1, 1)
            layers.append(Conv3dReLU(c_prev, c_out, ks, use_leaky=use_leaky))
            c_prev = c_out
        self.corr_conv = nn.Sequential(*layers)

        layers, widths = [], self.num_output
        for i, c_out in enumerate(widths):
            ks = (self.filter_size, 1) if i == 0 else (1, 1)
            if i == len(widths) - 1 and not last_relu:
                layers.append(nn.Conv2d(c_prev, c_out, kernel_size=ks))
            else:
                layers.append(Conv2dReLU(c_prev, c_out, ks, use_leaky=use_leaky))
            c_prev = c_out
        self.blur_conv = nn.Sequential(*layers)

    def get_filter_size(self, dist):
        return (dist + 1) ** self.d1 - dist ** self.d1

    def forward(self, feat1, feat2, prev_corr_feat, barycentric1, lattice_offset1, pc1_corr_indices,
                pc2_corr_indices, max_hash_cnt1, max_hash_cnt2):
        """(1, C, H1), (1, C, H2), (1, C', N) or None -> (1, num_output[-1], H1); bnn_flow.py:96-113."""
        if feat1.size(0) != 1:
            raise ValueError("batch size must be 1 (reference README.md:57)")
        if not feat1.is_cuda:
            raise RuntimeError("BilateralCorrelationFlex (B200) has no CPU path: inputs must be CUDA tensors")
        if (prev_corr_feat is None) != (self.prev_corr_dim == 0):
            raise ValueError("prev_corr_feat must be given iff prev_corr_dim != 0")
        params, corr_acts, blur_acts = [], [], []
        for layer in self.corr_conv:
            params += [layer.conv.weight, layer.conv.bias]
            corr_acts.append(_act_code(True, self.use_leaky))
        for conv, act in conv2d_layers(self.blur_conv, self.use_leaky):
            params += [conv.weight, conv.bias]
            blur_acts.append(act)
        cfg = (self.use_norm, tuple(corr_acts), tuple(blur_acts))
        return _CorrFunction.apply(cfg, feat1, feat2, prev_corr_feat, barycentric1, lattice_offset1,
                                   pc1_corr_indices, pc2_corr_indices, *params)
