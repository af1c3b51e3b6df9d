"""Callers of the hot path built from SURVEY §8f (checkpoint files, evaluation metrics): CPU-only checks against
values produced by the reference's own functions (recorded below with the command that made them)."""
import os

import numpy as np
import torch

from hplflownet_b200 import checkpoint as C
from hplflownet_b200 import evaluation_utils as E
from tests._util import ModelArgs, ShallowArgs

# Produced in the build container by the UNMODIFIED reference (numpy 2 removed np.float, hence the alias):
#   np.float = float; from evaluation_utils import evaluate_3d, evaluate_2d      (/root/reference)
#   rs = np.random.RandomState(0); gt = rs.normal(0, .3, (4096, 3)); pred = gt + rs.normal(0, .08, gt.shape)
#   evaluate_3d(pred.astype(np.float32), gt.astype(np.float32));  evaluate_2d(60 * pred[:, :2], 60 * gt[:, :2]) (fp32)
REF_3D = (0.12639212608337402, 0.056884765625, 0.34716796875, 0.926513671875)
REF_2D = (6.012866020202637, 0.169921875)


def _inputs():
    rs = np.random.RandomState(0)
    gt = rs.normal(0, .3, (4096, 3))
    pred = gt + rs.normal(0, .08, gt.shape)
    return pred.astype(np.float32), gt.astype(np.float32)


def test_evaluate_3d_2d_match_reference_values():
    pred, gt = _inputs()
    got3 = E.evaluate_3d(torch.from_numpy(pred)[None], torch.from_numpy(gt)[None])
    got2 = E.evaluate_2d(60 * pred[:, :2], 60 * gt[:, :2])
    assert np.allclose(got3, REF_3D, rtol=1e-6, atol=1e-9), (got3, REF_3D)
    assert np.allclose(got2, REF_2D, rtol=1e-6, atol=1e-9), (got2, REF_2D)


def test_checkpoint_roundtrip_reference_layout(tmp_path):
    from hplflownet_b200.HPLFlowNet import HPLFlowNet
    from hplflownet_b200.HPLFlowNet_shallow import HPLFlowNetShallow
    for cls, args, arch in ((HPLFlowNet, ModelArgs(), "HPLFlowNet"), (HPLFlowNetShallow, ShallowArgs(), "HPLFlowNetShallow")):
        torch.manual_seed(1)
        model = cls(args)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        state = C.make_state(model, opt, epoch=10, min_loss=0.123, arch=arch)
        # the layout main.py:183-189 writes under DataParallel
        assert set(state) == {"epoch", "arch", "state_dict", "min_loss", "optimizer"} and state["epoch"] == 11
        assert all(k.startswith("module.") for k in state["state_dict"])
        path = C.save_checkpoint(state, True, str(tmp_path))
        assert os.path.exists(os.path.join(str(tmp_path), "model_best.pth.tar"))
        assert os.path.exists(os.path.join(str(tmp_path), "checkpoint_11.pth.tar"))      # epoch % 10 == 1
        torch.manual_seed(2)
        other = cls(args)
        ckpt = C.load_checkpoint(path, other, torch.optim.Adam(other.parameters(), lr=1e-4))
        assert ckpt["epoch"] == 11 and ckpt["arch"] == arch and abs(ckpt["min_loss"] - 0.123) < 1e-12
        for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
            assert torch.equal(a, b), k
        # a bare (un-prefixed) state_dict loads as well
        other.load_state_dict(C.strip_module_prefix(model.state_dict()), strict=True)


def test_dense_permutation_detection():
    """ops._dense_permutation decides whether a weight view can be handed to the strided kernel as is."""
    from hplflownet_b200 import ops
    w = torch.randn(6, 5, 4)                                   # "conv layout" (Co, C, F)
    assert ops._dense_permutation(w)
    assert ops._dense_permutation(w.permute(2, 1, 0))          # (F, C, Co) view: forward operand
    assert ops._dense_permutation(w.permute(2, 1, 0).transpose(1, 2))   # (F, Co, C): data-gradient operand
    assert ops._dense_permutation(torch.randn(1, 5, 4).permute(0, 2, 1))   # size-1 dims carry arbitrary strides
    assert not ops._dense_permutation(w[:, :, ::2])            # holes
    assert not ops._dense_permutation(w[:3])  or w[:3].is_contiguous()     # a leading slice is still dense
    assert not ops._dense_permutation(w.expand(2, 6, 5, 4)[0][:, :, :2])   # partial last dim


def test_numa_binding_is_best_effort_without_gpu():
    from hplflownet_b200 import sharding
    assert sharding.bind_to_gpu_numa_node(0) is None or isinstance(sharding.bind_to_gpu_numa_node(0), int)
