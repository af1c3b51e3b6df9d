"""HPLFlowNet wired on top of the B200 bilateral layers (the caller of the hot path).

Same constructor argument object, ``forward(pc1, pc2, generated_data)`` signature, module names and
``state_dict`` layout as ``models/HPLFlowNet.py:11-430`` so reference checkpoints load with
``strict=True``; the wiring is expressed as data (tables of layer widths) instead of the reference's
unrolled code.  Every layer -- bilateral, correlation and the pointwise ``Conv1d`` stacks (``conv1``,
``conv2-4``, via ``pointwise.py``) -- runs the hand-written CUDA path; the modules only hold parameters.
"""
import torch
import torch.nn as nn

from .bilateralNN import BilateralConvFlex
from .bnn_flow import BilateralCorrelationFlex
from .module_utils import Conv1dReLU
from .pointwise import pointwise_stack

__all__ = ["HPLFlowNet"]

N_SCALES = 7
# up-path layers bcn{k}_ : (extra input channels besides el_minus_gr, output widths)  HPLFlowNet.py:37-221
_UP = {1: (64 + 512, [1024, 1024]), 2: (64 + 256, [512, 512]), 3: (64 * 2 + 256, [256, 256]),
       4: (64 * 2 + 128, [256, 256]), 5: (64 * 2 + 128, [128, 128]), 6: (64 * 2 + 128, [128, 128])}
_FIRST_CORR_SCALE = 2     # corr1 lives on scales_filter_map[2]  (HPLFlowNet.py:92-101)


def _count(v):
    return int(v.item()) if torch.is_tensor(v) else int(v)


class HPLFlowNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.scales_filter_map = args.scales_filter_map
        assert len(self.scales_filter_map) == N_SCALES
        self.chunk_size = -1 if args.evaluate else 1024 * 1024 * 25
        dim, leaky = args.dim, args.use_leaky

        def bcl(radius, c_in, widths, splat, slice_):
            return BilateralConvFlex(dim, radius, c_in, widths, args.DEVICE, use_bias=args.bcn_use_bias,
                                     use_leaky=leaky, use_norm=args.bcn_use_norm, do_splat=splat, do_slice=slice_,
                                     last_relu=args.last_relu, chunk_size=self.chunk_size)

        self.conv1 = nn.Sequential(Conv1dReLU(dim, 32, use_leaky=leaky), Conv1dReLU(32, 32, use_leaky=leaky),
                                   Conv1dReLU(32, 64, use_leaky=leaky))
        # registration order follows the reference so parameter iteration order matches
        for k in range(1, N_SCALES + 1):
            radius = self.scales_filter_map[k - 1][1]
            setattr(self, "bcn%d" % k, bcl(radius, 64 + dim + 1, [64, 64], True, False))
            if k in _UP:
                extra, widths = _UP[k]
                setattr(self, "bcn%d_" % k, bcl(radius, dim + 1 + extra, widths, False, True))
            else:
                setattr(self, "bcn%d_" % k, bcl(radius, 64 * 2, [128, 128], False, True))
            if k - 1 >= _FIRST_CORR_SCALE:
                row = self.scales_filter_map[k - 1]
                setattr(self, "corr%d" % (k - _FIRST_CORR_SCALE),
                        BilateralCorrelationFlex(dim, row[2], row[3], 64, [32, 32], [64, 64], args.DEVICE,
                                                 use_bias=args.bcn_use_bias, use_leaky=leaky, use_norm=args.bcn_use_norm,
                                                 prev_corr_dim=0 if k - 1 == _FIRST_CORR_SCALE else 64,
                                                 last_relu=args.last_relu, chunk_size=self.chunk_size))
        self.conv2 = Conv1dReLU(1024, 1024, use_leaky=leaky)
        self.conv3 = Conv1dReLU(1024, 512, use_leaky=leaky)
        self.conv4 = nn.Conv1d(512, 3, kernel_size=1)

    def forward(self, pc1, pc2, generated_data):
        gd = generated_data
        # pointwise Conv1d stacks run on the same tensor-core GEMM (fp32-accurate; cuDNN would use TF32 by default)
        down1 = [pointwise_stack(self.conv1, pc1)]                   # HPLFlowNet.py:239-240
        down2 = [pointwise_stack(self.conv1, pc2)]
        corr = [None] * N_SCALES
        prev_corr = None
        for k in range(N_SCALES):                                    # :242-369
            layer = getattr(self, "bcn%d" % (k + 1))
            outs = []
            for tag, feats in (("pc1", down1), ("pc2", down2)):
                x = torch.cat((gd[k][tag + "_el_minus_gr"], feats[-1]), dim=1)
                outs.append(layer(x, in_barycentric=gd[k][tag + "_barycentric"],
                                  in_lattice_offset=gd[k][tag + "_lattice_offset"],
                                  blur_neighbors=gd[k][tag + "_blur_neighbors"],
                                  out_barycentric=None, out_lattice_offset=None))
            down1.append(outs[0])
            down2.append(outs[1])
            if k >= _FIRST_CORR_SCALE:
                first = k == _FIRST_CORR_SCALE
                prev_corr = getattr(self, "corr%d" % (k - _FIRST_CORR_SCALE + 1))(
                    outs[0], outs[1], prev_corr,
                    barycentric1=None if first else gd[k]["pc1_barycentric"],
                    lattice_offset1=None if first else gd[k]["pc1_lattice_offset"],
                    pc1_corr_indices=gd[k]["pc1_corr_indices"], pc2_corr_indices=gd[k]["pc2_corr_indices"],
                    max_hash_cnt1=_count(gd[k]["pc1_hash_cnt"]), max_hash_cnt2=_count(gd[k]["pc2_hash_cnt"]))
                corr[k] = prev_corr

        # up path (:372-423): level k consumes [el_minus_gr of level k+1, up, corr_k, skip_k]
        up = None
        for k in range(N_SCALES - 1, -1, -1):
            skip = down1[k + 1]
            if k == N_SCALES - 1:
                parts = (corr[k], skip)
            else:
                parts = (gd[k + 1]["pc1_el_minus_gr"], up) + ((corr[k],) if corr[k] is not None else ()) + (skip,)
            up = getattr(self, "bcn%d_" % (k + 1))(
                torch.cat(parts, dim=1), in_barycentric=None, in_lattice_offset=None,
                blur_neighbors=gd[k]["pc1_blur_neighbors"], out_barycentric=gd[k]["pc1_barycentric"],
                out_lattice_offset=gd[k]["pc1_lattice_offset"])
        return pointwise_stack((self.conv2, self.conv3, self.conv4), up)     # :426-428
