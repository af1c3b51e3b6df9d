"""oracle/hplflownet.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of ``HPLFlowNet.forward`` (models/HPLFlowNet.py:238-430) as a pure function of a
reference-format ``state_dict`` and ``generated_data``, composed from oracle/bcl.py.  Used to check the
B200 model wiring.  Parity status: pinned -- tests/test_oracle_golden.py compares it with
an output dumped from the unmodified reference model (oracle/make_golden.py, name-keyed seeded weights).
"""
import torch
import torch.nn.functional as F

from . import bcl as OB


def _sub(state, prefix):
    n = len(prefix) + 1
    return {k[n:]: v for k, v in state.items() if k.startswith(prefix + ".")}


def _conv1d_relu(x, w, b, leaky, act=True):
    y = F.conv1d(x, w, b)
    if not act:
        return y
    return OB._act(y, leaky)


def forward(state, pc1, pc2, gd, *, use_leaky=True, use_norm=True, use_bias=True):
    def conv1(x):
        for i in range(3):                                             # :21-24
            x = _conv1d_relu(x, state["conv1.%d.composed_module.0.weight" % i],
                             state["conv1.%d.composed_module.0.bias" % i], use_leaky)
        return x

    def count(v):
        return int(v.item()) if torch.is_tensor(v) else int(v)

    d1, d2 = [conv1(pc1)], [conv1(pc2)]
    corr, prev = [None] * 7, None
    for k in range(7):                                                 # :242-369
        st = _sub(state, "bcn%d" % (k + 1))
        outs = []
        for tag, feats in (("pc1", d1), ("pc2", d2)):
            x = torch.cat((gd[k][tag + "_el_minus_gr"], feats[-1]), dim=1)
            outs.append(OB.bcl_forward(st, x, gd[k][tag + "_barycentric"], gd[k][tag + "_lattice_offset"],
                                       gd[k][tag + "_blur_neighbors"], None, None, do_splat=True, do_slice=False,
                                       use_norm=use_norm, use_leaky=use_leaky, use_bias=use_bias))
        d1.append(outs[0]); d2.append(outs[1])
        if k >= 2:
            first = k == 2
            prev = OB.corr_forward(_sub(state, "corr%d" % (k - 1)), outs[0], outs[1], prev,
                                   None if first else gd[k]["pc1_barycentric"],
                                   None if first else gd[k]["pc1_lattice_offset"],
                                   gd[k]["pc1_corr_indices"], gd[k]["pc2_corr_indices"],
                                   use_norm=use_norm, use_leaky=use_leaky)
            corr[k] = prev
    up = None
    for k in range(6, -1, -1):                                         # :372-423
        skip = d1[k + 1]
        if k == 6:
            parts = (corr[k], skip)
        else:
            parts = (gd[k + 1]["pc1_el_minus_gr"], up) + ((corr[k],) if corr[k] is not None else ()) + (skip,)
        up = OB.bcl_forward(_sub(state, "bcn%d_" % (k + 1)), torch.cat(parts, dim=1), None, None,
                            gd[k]["pc1_blur_neighbors"], gd[k]["pc1_barycentric"], gd[k]["pc1_lattice_offset"],
                            do_splat=False, do_slice=True, use_norm=use_norm, use_leaky=use_leaky, use_bias=use_bias)
    x = _conv1d_relu(up, state["conv2.composed_module.0.weight"], state["conv2.composed_module.0.bias"], use_leaky)
    x = _conv1d_relu(x, state["conv3.composed_module.0.weight"], state["conv3.composed_module.0.bias"], use_leaky)
    return _conv1d_relu(x, state["conv4.weight"], state["conv4.bias"], use_leaky, act=False)


def forward_shallow(state, pc1, pc2, gd, *, use_leaky=True, use_norm=True, use_bias=True):
    """``HPLFlowNetShallow.forward`` (models/HPLFlowNet_shallow.py:171-311): 5 scales, single-conv bilateral
    layers, ``corr*_refine`` pointwise stacks.  Parity status: pinned by tests/golden/model_shallow_frustum256.npz
    (dumped from the unmodified reference model)."""
    def stack(prefix, n, x):
        for i in range(n):
            x = _conv1d_relu(x, state["%s.%d.composed_module.0.weight" % (prefix, i)],
                             state["%s.%d.composed_module.0.bias" % (prefix, i)], use_leaky)
        return x

    n_scales = 5
    d1, d2 = [stack("conv1", 3, pc1)], [stack("conv1", 3, pc2)]
    corr, prev = [None] * n_scales, None
    for k in range(n_scales):                                          # :175-268
        st = _sub(state, "bcn%d" % (k + 1))
        outs = []
        for tag, feats in (("pc1", d1), ("pc2", d2)):
            x = torch.cat((gd[k][tag + "_el_minus_gr"], feats[-1]), dim=1)
            outs.append(OB.bcl_forward(st, x, gd[k][tag + "_barycentric"], gd[k][tag + "_lattice_offset"],
                                       gd[k][tag + "_blur_neighbors"], None, None, do_splat=True, do_slice=False,
                                       use_norm=use_norm, use_leaky=use_leaky, use_bias=use_bias))
        d1.append(outs[0]); d2.append(outs[1])
        if k >= 2:
            first = k == 2
            c = OB.corr_forward(_sub(state, "corr%d" % (k - 1)), outs[0], outs[1], prev,
                                None if first else gd[k]["pc1_barycentric"],
                                None if first else gd[k]["pc1_lattice_offset"],
                                gd[k]["pc1_corr_indices"], gd[k]["pc2_corr_indices"],
                                use_norm=use_norm, use_leaky=use_leaky)
            if k + 1 < n_scales:
                c = torch.cat((gd[k + 1]["pc1_el_minus_gr"], c), dim=1)
            prev = stack("corr%d_refine" % (k - 1), 3, c)
            corr[k] = prev
    up = None
    for k in range(n_scales - 1, -1, -1):                              # :271-305
        skip = d1[k + 1]
        if k == n_scales - 1:
            parts = (corr[k], skip)
        else:
            parts = (gd[k + 1]["pc1_el_minus_gr"], up) + ((corr[k],) if corr[k] is not None else ()) + (skip,)
        up = OB.bcl_forward(_sub(state, "bcn%d_" % (k + 1)), torch.cat(parts, dim=1), None, None,
                            gd[k]["pc1_blur_neighbors"], gd[k]["pc1_barycentric"], gd[k]["pc1_lattice_offset"],
                            do_splat=False, do_slice=True, use_norm=use_norm, use_leaky=use_leaky, use_bias=use_bias)
    x = _conv1d_relu(up, state["conv2.composed_module.0.weight"], state["conv2.composed_module.0.bias"], use_leaky)
    x = _conv1d_relu(x, state["conv3.composed_module.0.weight"], state["conv3.composed_module.0.bias"], use_leaky)
    return _conv1d_relu(x, state["conv4.weight"], state["conv4.bias"], use_leaky, act=False)
