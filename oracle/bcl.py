"""oracle/bcl.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch fp32) restatement of the reference's value path: the bilateral
convolution layer ``BilateralConvFlex.forward`` (models/bilateralNN.py:122-238)
and the two-cloud ``BilateralCorrelationFlex.forward``
(models/bnn_flow.py:96-210), written as pure functions of a reference-format
``state_dict`` so that the CUDA modules can be checked against them on the same
weights.  Gradients come from autograd, as in the reference.  The same torch
op families as the reference are used (sparse COO densify, advanced-index
gather, ``conv2d``/``conv3d``) so that timing it on host cores is a fair
stand-in for the reference's CPU path.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import
this module.  Parity status: pinned -- tests/test_oracle_golden.py checks it
against outputs and gradients dumped from the unmodified reference modules
(oracle/make_golden.py), tests/test_oracle_vs_reference.py re-checks live when
/root/reference is importable.
"""
import warnings

import torch
import torch.nn.functional as F

warnings.filterwarnings("ignore", message="Sparse invariant checks are implicitly disabled")

LEAKY_RATE = 0.1  # models/module_utils.py:6


# Activation-kink ledger (test infrastructure): when KINK_LOG is a list, every activation call appends
# (units, units with |pre-activation| < KINK_TOL * max|pre-activation|).  (Leaky)ReLU derivatives are discontinuous at 0,
# so only a unit this close to zero can legitimately pick a different slope under another summation order; the parity
# tests allow their loose gradient bound only when the oracle itself reports such units (tests/_util.py).
KINK_LOG = None
KINK_TOL = 1e-5


def _act(x, use_leaky):
    if KINK_LOG is not None:
        m = x.detach().abs()
        KINK_LOG.append((m.numel(), int((m < KINK_TOL * m.max()).sum())))
    return F.leaky_relu(x, LEAKY_RATE) if use_leaky else F.relu(x)


def _conv_stack(state, prefix):
    """Collect [(weight, bias, has_act)] for ``prefix.{i}`` in module order.

    Layer naming follows the reference containers (bilateralNN.py:94-113):
    ``{prefix}.{i}.composed_module.0.*`` is conv+activation (module_utils.py:27-59),
    ``{prefix}.{i}.*`` is a bare conv (last layer when last_relu=False).
    """
    layers, i = [], 0
    while True:
        k_act = "%s.%d.composed_module.0.weight" % (prefix, i)
        k_lin = "%s.%d.weight" % (prefix, i)
        if k_act in state:
            layers.append((state[k_act], state[k_act[:-6] + "bias"], True))
        elif k_lin in state:
            layers.append((state[k_lin], state[k_lin[:-6] + "bias"], False))
        else:
            return layers
        i += 1


def splat(features, barycentric, lattice_offset, n_vertices, use_norm):
    """bilateralNN.py:150-186 (and bnn_flow.py:119-154).

    features (1, C, N) f32, barycentric (1, 4, N) f32, lattice_offset (1, 4, N) i64
    -> (1, C, n_vertices + 1); row 0 is the all-zero "null vertex".
    """
    c = features.size(1)
    idx = (lattice_offset + 1).reshape(1, -1)
    contrib = (barycentric[:, None] * features[:, :, None]).permute(1, 0, 2, 3).reshape(c, -1).t()
    dense = torch.sparse_coo_tensor(idx, contrib, (n_vertices + 1, c), check_invariants=False).to_dense()
    out = dense.reshape(1, n_vertices + 1, c).permute(0, 2, 1)
    if use_norm:
        w = torch.sparse_coo_tensor(idx, barycentric.reshape(-1, 1), (n_vertices + 1, 1),
                                    check_invariants=False).to_dense()
        out = out * (1.0 / (w.reshape(1, n_vertices + 1) + 1e-5))[:, None, :]
    return out


def _pad_null(features):
    """bilateralNN.py:190-196: prepend the null-vertex column."""
    z = torch.zeros((features.size(0), features.size(1), 1), dtype=features.dtype, device=features.device)
    return torch.cat((z, features), dim=-1)


def bcl_forward(state, features, in_barycentric, in_lattice_offset, blur_neighbors,
                out_barycentric, out_lattice_offset, *, do_splat, do_slice, use_norm, use_leaky,
                use_bias):
    """BilateralConvFlex.forward, bilateralNN.py:122-238, B = 1, no chunking
    (chunking only bounds memory; the result is the concatenation, :207-221)."""
    h = blur_neighbors.size(-1)
    lat = splat(features, in_barycentric, in_lattice_offset, h, use_norm) if do_splat \
        else _pad_null(features)
    x = lat[0][:, (blur_neighbors[0] + 1)].unsqueeze(0)             # (1, C, F, H)  :215-217
    for w, b, has_act in _conv_stack(state, "blur_conv"):
        x = F.conv2d(x, w, b)
        if has_act:
            x = _act(x, use_leaky)
    x = x.squeeze(2)                                                # (1, Co, H)
    if not do_slice:
        return x
    g = x[0][:, out_lattice_offset[0]].unsqueeze(0)                 # (1, Co, 4, N) :226-228
    y = (out_barycentric[:, None] * g).sum(dim=2)
    if use_bias:
        y = y + state["bias"][None, :, None]
    return y


def corr_forward(state, feat1, feat2, prev_corr_feat, barycentric1, lattice_offset1,
                 pc1_corr_indices, pc2_corr_indices, *, use_norm, use_leaky):
    """BilateralCorrelationFlex.forward, bnn_flow.py:96-210, B = 1, no chunking."""
    h1 = feat1.size(-1)
    s1, s2 = _pad_null(feat1), _pad_null(feat2)                     # :156-166
    if prev_corr_feat is not None:                                  # :119-154, :167-168
        s1 = torch.cat((splat(prev_corr_feat, barycentric1, lattice_offset1, h1, use_norm), s1), 1)
    fs = pc2_corr_indices.size(1)
    a = s1[0][:, pc1_corr_indices[0] + 1]                           # (C1, P, H)    :189-191
    a = a[None, :, None].expand(-1, -1, fs, -1, -1)                 # repeat over F :192
    b = s2[0][:, pc2_corr_indices[0] + 1].unsqueeze(0)              # (1, C, F, P, H) :195-197
    x = torch.cat((a, b), dim=1)                                    # :199
    for w, bias, has_act in _conv_stack(state, "corr_conv"):        # :202
        x = F.conv3d(x, w, bias)
        if has_act:
            x = _act(x, use_leaky)
    x = x.squeeze(3)                                                # (1, 32, F, H)
    for w, bias, has_act in _conv_stack(state, "blur_conv"):        # :205
        x = F.conv2d(x, w, bias)
        if has_act:
            x = _act(x, use_leaky)
    return x.squeeze(2)
