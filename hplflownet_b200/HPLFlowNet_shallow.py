"""HPLFlowNetShallow on the B200 bilateral layers.

Same constructor argument object, ``forward(pc1, pc2, generated_data)`` signature, module names, registration
order and ``state_dict`` layout as ``models/HPLFlowNet_shallow.py:11-311``: five lattice scales, single-conv
bilateral layers (``num_output`` of length 1 -- with ``last_relu=False`` that is one bare convolution,
``bilateralNN.py:104-113``), correlation layers with one conv per stage followed by pointwise ``corr*_refine``
stacks.  Every layer runs the hand-written CUDA path; the modules only hold parameters.
"""
import torch
import torch.nn as nn

from .bilateralNN import BilateralConvFlex
from .bnn_flow import BilateralCorrelationFlex
from .module_utils import Conv1dReLU
from .pointwise import pointwise_stack

__all__ = ["HPLFlowNetShallow"]

N_SCALES = 5
_FIRST_CORR_SCALE = 2          # corr1 lives on scales_filter_map[2]  (HPLFlowNet_shallow.py:86-95)
# up-path layers bcn{k}_ : (input channels, output width)                  (:33-42, :55-64, :77-84, :113-122, :143-151)
_UP_WIDTH = {1: 128, 2: 64, 3: 64, 4: 64, 5: 64}


def _count(v):
    return int(v.item()) if torch.is_tensor(v) else int(v)


class HPLFlowNetShallow(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.scales_filter_map = args.scales_filter_map
        assert len(self.scales_filter_map) == N_SCALES
        dim, leaky = args.dim, args.use_leaky

        def bcl(radius, c_in, width, splat, slice_):
            return BilateralConvFlex(dim, radius, c_in, [width], args.DEVICE, use_bias=args.bcn_use_bias,
                                     use_leaky=leaky, use_norm=args.bcn_use_norm, do_splat=splat, do_slice=slice_,
                                     last_relu=args.last_relu)

        def refine(c_in):
            return nn.Sequential(Conv1dReLU(c_in, 64, use_leaky=leaky), Conv1dReLU(64, 64, use_leaky=leaky),
                                 Conv1dReLU(64, 64, use_leaky=leaky))

        self.conv1 = nn.Sequential(Conv1dReLU(dim, 32, use_leaky=leaky), Conv1dReLU(32, 32, use_leaky=leaky),
                                   Conv1dReLU(32, 64, use_leaky=leaky))
        # registration order follows the reference so parameter iteration order matches
        for k in range(1, N_SCALES + 1):
            row = self.scales_filter_map[k - 1]
            setattr(self, "bcn%d" % k, bcl(row[1], 64 + dim + 1, 64, True, False))
            if k == N_SCALES:
                c_up = 64 + 64                                       # [corr3_refine, skip]
            elif k - 1 >= _FIRST_CORR_SCALE:
                c_up = dim + 1 + 64 * 2 + 64                         # [el_minus_gr, up, corr_refine, skip]
            else:
                c_up = dim + 1 + 64 + 64                             # [el_minus_gr, up, skip]
            setattr(self, "bcn%d_" % k, bcl(row[1], c_up, _UP_WIDTH[k], False, True))
            if k - 1 >= _FIRST_CORR_SCALE:
                j = k - _FIRST_CORR_SCALE
                setattr(self, "corr%d" % j,
                        BilateralCorrelationFlex(dim, row[2], row[3], 64, [32], [32], args.DEVICE,
                                                 use_bias=args.bcn_use_bias, use_leaky=leaky, use_norm=args.bcn_use_norm,
                                                 prev_corr_dim=0 if k - 1 == _FIRST_CORR_SCALE else 64,
                                                 last_relu=args.last_relu))
                # the refined correlation is splatted onto the NEXT scale, whose el_minus_gr it is concatenated with
                setattr(self, "corr%d_refine" % j, refine(32 if k == N_SCALES else 32 + dim + 1))
        self.conv2 = Conv1dReLU(128, 1024, use_leaky=leaky)
        self.conv3 = Conv1dReLU(1024, 512, use_leaky=leaky)
        self.conv4 = nn.Conv1d(512, 3, kernel_size=1)

    def forward(self, pc1, pc2, generated_data):
        gd = generated_data
        down1 = [pointwise_stack(self.conv1, pc1)]                   # HPLFlowNet_shallow.py:172-173
        down2 = [pointwise_stack(self.conv1, pc2)]
        corr = [None] * N_SCALES                                     # refined correlation features per scale
        prev = None
        for k in range(N_SCALES):                                    # :175-268
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
                j = k - _FIRST_CORR_SCALE + 1
                first = k == _FIRST_CORR_SCALE
                c = getattr(self, "corr%d" % j)(
                    outs[0], outs[1], prev,
                    barycentric1=None if first else gd[k]["pc1_barycentric"],
                    lattice_offset1=None if first else gd[k]["pc1_lattice_offset"],
                    pc1_corr_indices=gd[k]["pc1_corr_indices"], pc2_corr_indices=gd[k]["pc2_corr_indices"],
                    max_hash_cnt1=_count(gd[k]["pc1_hash_cnt"]), max_hash_cnt2=_count(gd[k]["pc2_hash_cnt"]))
                if k + 1 < N_SCALES:                                 # vertices of scale k are the points of scale k+1
                    c = torch.cat((gd[k + 1]["pc1_el_minus_gr"], c), dim=1)
                prev = pointwise_stack(getattr(self, "corr%d_refine" % j), c)
                corr[k] = prev

        up = None
        for k in range(N_SCALES - 1, -1, -1):                        # :271-305
            skip = down1[k + 1]
            if k == N_SCALES - 1:
                parts = (corr[k], skip)
            else:
                parts = (gd[k + 1]["pc1_el_minus_gr"], up) + ((corr[k],) if corr[k] is not None else ()) + (skip,)
            up = getattr(self, "bcn%d_" % (k + 1))(
                torch.cat(parts, dim=1), in_barycentric=None, in_lattice_offset=None,
                blur_neighbors=gd[k]["pc1_blur_neighbors"], out_barycentric=gd[k]["pc1_barycentric"],
                out_lattice_offset=gd[k]["pc1_lattice_offset"])
        return pointwise_stack((self.conv2, self.conv3, self.conv4), up)     # :308-311
