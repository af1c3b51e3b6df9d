"""Pins the oracle (oracle/) against fixtures dumped from the unmodified
reference by oracle/make_golden.py.  CPU only.

Index path: bit-exact.  Value path: <= 1e-5 relative (REL_TOL) -- the oracle
restates the same torch ops, so it is usually exact, but summation order inside
sparse densify is not part of the contract.
"""
import numpy as np
import pytest
import torch

from oracle import bcl as OB
from oracle import lattice as OL
from tests._util import (assert_close, bits_equal, golden, golden_files, grads_from, state_from, t)

LATTICE_KEYS = ["barycentric", "el_minus_gr", "lattice_offset", "blur_neighbors"]


def test_constants_match_reference_fixture():
    # fp32 bit patterns of elevate_mat, expected_std, canonical table, offsets order
    e, std = OL.constants()
    want = [0x3f3504f3, 0x3ed105eb, 0x3e93cd3a, 0xbf3504f3, 0x3ed105eb, 0x3e93cd3a,
            0x00000000, 0xbf5105eb, 0x3e93cd3a, 0x00000000, 0x00000000, 0xbf5db3d7]
    assert e.view(np.uint32).ravel().tolist() == want
    assert std == 4 * (2.0 / 3.0) ** 0.5
    offs = OL.neighbor_offsets(1)
    assert offs.tolist() == [[0, 0, 0, 0], [-1, -1, -1, 3], [-1, -1, 3, -1], [-2, -2, 2, 2], [-1, 3, -1, -1],
                             [-2, 2, -2, 2], [-2, 2, 2, -2], [-3, 1, 1, 1], [3, -1, -1, -1], [2, -2, -2, 2],
                             [2, -2, 2, -2], [1, -3, 1, 1], [2, 2, -2, -2], [1, 1, -3, 1], [1, 1, 1, -3]]
    assert OL.filter_size(1) == 15 and OL.filter_size(2) == 65
    assert OL.neighbor_offsets(2).shape == (65, 4)
    assert (OL.neighbor_offsets(2).sum(1) == 0).all()


@pytest.mark.parametrize("name", golden_files("lattice_"))
def test_lattice_oracle_bit_exact(name):
    g = golden(name)
    sfm = [[float(r[0]), int(r[1]), int(r[2]), int(r[3])] for r in g["scales_filter_map"]]
    got = OL.generate(g["pc1"], g["pc2"], sfm)
    assert len(got) == len(sfm)
    for k, d in enumerate(got):
        for key, v in d.items():
            ref = g["s%d_%s" % (k, key)]
            if isinstance(v, int):
                assert v == int(ref), (k, key)
            else:
                assert bits_equal(v, ref), "scale %d %s differs from the reference" % (k, key)


def test_lattice_invariants():
    # properties the reference implies, at BASELINE size
    from hplflownet_b200.synthetic import frustum_pair
    pc1, pc2 = frustum_pair(8192, 0)
    d = OL.generate(pc1, pc2, [[1.0, 1, 1, 1]])[0]
    keys, bary, _ = OL.keys_and_barycentric(np.ascontiguousarray(pc1.T))
    assert (keys.sum(0) == 0).all()                      # every key lies on the hyperplane
    assert np.abs(bary.sum(0) - 1).max() < 3e-7 and bary.min() > -1e-6
    off = d["pc1_lattice_offset"]
    flat = off.T.reshape(-1)                             # point outer, remainder inner
    _, first = np.unique(flat, return_index=True)
    assert (np.sort(flat[np.sort(first)]) == np.arange(d["pc1_hash_cnt"])).all()
    assert (flat[np.sort(first)] == np.arange(d["pc1_hash_cnt"])).all()   # ids = first-occurrence order
    nbr = d["pc1_blur_neighbors"]
    assert (nbr[0] == np.arange(nbr.shape[1])).all()     # offset 0 is the vertex itself
    assert nbr.min() >= -1 and nbr.max() < d["pc1_hash_cnt"]


def _bcl_kwargs(g):
    c_in, do_splat, do_slice, use_norm, use_leaky, use_bias, last_relu = [int(x) for x in g["cfg"]]
    return dict(do_splat=bool(do_splat), do_slice=bool(do_slice), use_norm=bool(use_norm),
                use_leaky=bool(use_leaky), use_bias=bool(use_bias))


@pytest.mark.parametrize("name", golden_files("bcl_"))
def test_bcl_oracle_matches_reference(name):
    g = golden(name)
    kw = _bcl_kwargs(g)
    state = {k: v.requires_grad_(v.is_floating_point()) for k, v in state_from(g).items()}
    feat = t(g["features"]).requires_grad_(True)
    bary, off, nbr = t(g["barycentric"]), t(g["lattice_offset"]), t(g["blur_neighbors"])
    y = OB.bcl_forward(state, feat, bary, off, nbr, bary, off, **kw)
    assert_close(y, g["output"], "output")
    y.backward(t(g["grad_output"]))
    assert_close(feat.grad, g["grad_features"], "grad_features")
    for k, ref in grads_from(g).items():
        assert_close(state[k].grad, ref, "grad " + k)


@pytest.mark.parametrize("name", golden_files("corr_"))
def test_corr_oracle_matches_reference(name):
    g = golden(name)
    c, prev_dim, use_leaky, last_relu = [int(x) for x in g["cfg"]]
    state = {k: v.requires_grad_(v.is_floating_point()) for k, v in state_from(g).items()}
    f1, f2 = t(g["feat1"]).requires_grad_(True), t(g["feat2"]).requires_grad_(True)
    prev = t(g["prev_corr_feat"]).requires_grad_(True) if prev_dim else None
    y = OB.corr_forward(state, f1, f2, prev, t(g["barycentric1"]), t(g["lattice_offset1"]),
                        t(g["pc1_corr_indices"]), t(g["pc2_corr_indices"]),
                        use_norm=True, use_leaky=bool(use_leaky))
    assert_close(y, g["output"], "output")
    y.backward(t(g["grad_output"]))
    assert_close(f1.grad, g["grad_feat1"], "grad_feat1")
    assert_close(f2.grad, g["grad_feat2"], "grad_feat2")
    if prev_dim:
        assert_close(prev.grad, g["grad_prev"], "grad_prev")
    for k, ref in grads_from(g).items():
        assert_close(state[k].grad, ref, "grad " + k)


def test_model_oracle_matches_reference_output():
    # full HPLFlowNet forward: oracle composition vs the reference model's own output
    from oracle import hplflownet as OM
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from tests._util import ModelArgs, name_keyed_init_
    g = golden("model_frustum256.npz")
    model = name_keyed_init_(HPLFlowNet(ModelArgs()), int(g["seed"]))
    state = {k: v.detach() for k, v in model.state_dict().items()}
    gd = OL.generate(g["pc1"], g["pc2"], ModelArgs.scales_filter_map)
    gd = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd]
    pc1, pc2 = [torch.from_numpy(np.ascontiguousarray(g[k].T))[None] for k in ("pc1", "pc2")]
    with torch.no_grad():
        out = OM.forward(state, pc1, pc2, gd)
    assert out.shape == (1, 3, 256)
    assert_close(out, g["output"], "flow", tol=1e-5)


def test_shallow_model_oracle_matches_reference_output():
    # HPLFlowNetShallow forward: oracle composition vs the reference model's own output
    from oracle import hplflownet as OM
    from hplflownet_b200.HPLFlowNet_shallow import HPLFlowNetShallow
    from tests._util import ShallowArgs, name_keyed_init_
    g = golden("model_shallow_frustum256.npz")
    model = name_keyed_init_(HPLFlowNetShallow(ShallowArgs()), int(g["seed"]))
    state = {k: v.detach() for k, v in model.state_dict().items()}
    gd = OL.generate(g["pc1"], g["pc2"], ShallowArgs.scales_filter_map)
    gd = [{k: (torch.from_numpy(v)[None] if not isinstance(v, int) else v) for k, v in d.items()} for d in gd]
    pc1, pc2 = [torch.from_numpy(np.ascontiguousarray(g[k].T))[None] for k in ("pc1", "pc2")]
    with torch.no_grad():
        out = OM.forward_shallow(state, pc1, pc2, gd)
    assert out.shape == (1, 3, 256)
    assert_close(out, g["output"], "flow", tol=1e-5)


def test_shallow_model_state_dict_layout():
    # module names / registration order / shapes the reference checkpoint loader relies on (main.py:122, strict=True)
    from hplflownet_b200.HPLFlowNet_shallow import HPLFlowNetShallow
    from tests._util import ShallowArgs
    sd = HPLFlowNetShallow(ShallowArgs()).state_dict()
    keys = list(sd.keys())
    assert len(keys) == 90 and keys[0] == "conv1.0.composed_module.0.weight" and keys[-1] == "conv4.bias"
    assert tuple(sd["bcn1.blur_conv.0.weight"].shape) == (64, 68, 15, 1)            # single bare conv (last_relu=False)
    assert tuple(sd["bcn1_.blur_conv.0.weight"].shape) == (128, 132, 15, 1)
    assert tuple(sd["corr2.corr_conv.0.composed_module.0.weight"].shape) == (32, 192, 1, 15, 1)
    assert tuple(sd["corr1_refine.0.composed_module.0.weight"].shape) == (64, 36, 1)
    assert tuple(sd["corr3_refine.0.composed_module.0.weight"].shape) == (64, 32, 1)
    order = [k.split(".")[0] for k in keys]
    seen = [m for i, m in enumerate(order) if m not in order[:i]]
    assert seen == ["conv1", "bcn1", "bcn1_", "bcn2", "bcn2_", "bcn3", "bcn3_", "corr1", "corr1_refine", "bcn4", "bcn4_",
                    "corr2", "corr2_refine", "bcn5", "bcn5_", "corr3", "corr3_refine", "conv2", "conv3", "conv4"]
