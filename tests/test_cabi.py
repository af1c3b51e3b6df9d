"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/hplflownet_b200.h declares; the host modules keep the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

import hplflownet_b200 as hpl
from hplflownet_b200 import _lib
from tests._util import golden, golden_files, state_from

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "hplflownet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_library_identifies_itself():
    lib = _lib.load()
    assert lib.hpl_version() >= 100
    assert lib.hpl_sm_arch() == 100


def test_argument_errors_are_reported_without_a_gpu():
    # null pointers / bad leading dimension are rejected before any CUDA call
    lib = _lib.load()
    assert lib.hpl_scatter_rows(None, None, None, 1, 8, 4, None, 4, 8, None, None, None) == -1
    with pytest.raises(_lib.HplError):
        _lib.call("hpl_normalize_rows", None, 4, 1, 4, None, None, None)


@pytest.mark.parametrize("name", golden_files("bcl_"))
def test_bcl_state_dict_layout_matches_reference(name):
    g = golden(name)
    c_in, do_splat, do_slice, use_norm, use_leaky, use_bias, last_relu = [int(x) for x in g["cfg"]]
    mod = hpl.BilateralConvFlex(3, 1, c_in, [int(c) for c in g["c_out"]], "cuda", use_bias=bool(use_bias),
                                use_leaky=bool(use_leaky), use_norm=bool(use_norm), do_splat=bool(do_splat),
                                do_slice=bool(do_slice), last_relu=bool(last_relu), chunk_size=-1)
    ref_state = state_from(g)
    mine = mod.state_dict()
    assert sorted(mine) == sorted(ref_state)
    for k in mine:
        assert mine[k].shape == ref_state[k].shape and mine[k].dtype == ref_state[k].dtype, k
    mod.load_state_dict(ref_state, strict=True)
    assert mod.filter_size == 15


def test_cpu_tensors_are_rejected_loudly():
    mod = hpl.BilateralConvFlex(3, 1, 4, [4], "cuda", True, True, True, True, True, False)
    x = torch.zeros(1, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        mod(x, None, None, torch.zeros(1, 15, 3, dtype=torch.long), None, None)


def test_neighbor_offsets_match_reference_order():
    # Traverse.go order for r=1 (conv-weight index f <-> offset f), and vs the oracle for r=2
    from hplflownet_b200.transforms import filter_size, neighbor_offsets
    from oracle import lattice as OL
    assert filter_size(1) == 15
    for r in (1, 2, 3):
        assert (neighbor_offsets(r) == OL.neighbor_offsets(r)).all()


@pytest.mark.parametrize("name", golden_files("corr_"))
def test_corr_state_dict_layout_matches_reference(name):
    g = golden(name)
    c, prev_dim, use_leaky, last_relu = [int(x) for x in g["cfg"]]
    mod = hpl.BilateralCorrelationFlex(3, 1, 1, c, [int(x) for x in g["corr_out"]], [int(x) for x in g["out_ch"]],
                                       "cuda", use_bias=True, use_leaky=bool(use_leaky), use_norm=True,
                                       prev_corr_dim=prev_dim, last_relu=bool(last_relu), chunk_size=-1)
    ref_state = state_from(g)
    mine = mod.state_dict()
    assert sorted(mine) == sorted(ref_state)
    for k in mine:
        assert mine[k].shape == ref_state[k].shape and mine[k].dtype == ref_state[k].dtype, k
    mod.load_state_dict(ref_state, strict=True)
    assert mod.filter_size == 15 and mod.corr_size == 15


def test_model_state_dict_is_reference_layout():
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from tests._util import ModelArgs
    model = HPLFlowNet(ModelArgs())
    sd = model.state_dict()
    assert sum(p.numel() for p in model.parameters()) == 19303843
    for k in ("conv1.0.composed_module.0.weight", "bcn1.blur_conv.0.composed_module.0.weight", "bcn1.blur_conv.1.bias",
              "bcn1_.bias", "bcn1_.out_indices", "corr1.corr_conv.0.composed_module.0.weight", "corr2.feat1_indices",
              "conv4.weight"):
        assert k in sd, k
    assert sd["bcn1_.blur_conv.0.composed_module.0.weight"].shape == (1024, 580, 15, 1)
    assert sd["corr2.corr_conv.0.composed_module.0.weight"].shape == (32, 192, 1, 15, 1)


def test_full_state_dict_layout_matches_the_reference_models():
    """Every key, shape and dtype of the reference HPLFlowNet / HPLFlowNetShallow ``state_dict`` (dumped from the
    unmodified reference by oracle/make_golden.py:dump_state_layout; main.py:122 loads checkpoints with strict=True)."""
    import json
    import os
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from hplflownet_b200.HPLFlowNet_shallow import HPLFlowNetShallow
    from tests._util import GOLDEN, ModelArgs, ShallowArgs
    want = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))
    for name, cls, args in (("HPLFlowNet", HPLFlowNet, ModelArgs()), ("HPLFlowNetShallow", HPLFlowNetShallow, ShallowArgs())):
        sd = cls(args).state_dict()
        got = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
        assert list(got) == sorted(got, key=list(got).index)          # (dict order is insertion order)
        assert set(got) == set(want[name]), (sorted(set(got) ^ set(want[name]))[:10])
        for k, v in want[name].items():
            assert got[k] == v, (name, k, got[k], v)
