"""Drop-in check (north_star: "drops into models/HPLFlowNet.py unchanged"): the UNMODIFIED reference model files
(models/HPLFlowNet.py, models/HPLFlowNet_shallow.py; staged git-ignored under baseline/_ref/ by __graft_entry__.build())
are executed with their relative imports ``.bilateralNN`` / ``.bnn_flow`` / ``.module_utils`` resolved to this package,
fed by the GPU lattice builder, and compared with the outputs the reference produced on its own modules (fixtures)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from tests._util import ModelArgs, ShallowArgs, assert_close, golden, name_keyed_init_

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODELS = os.path.join(REPO, "baseline", "_ref", "models")


def _reference_model_class(file_name, class_name):
    path = os.path.join(REF_MODELS, file_name)
    if not os.path.exists(path):
        pytest.skip("reference model sources are not staged (run __graft_entry__.build() where /root/reference exists)")
    import hplflownet_b200.bilateralNN
    import hplflownet_b200.bnn_flow
    import hplflownet_b200.module_utils
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = []                                    # a package: relative imports resolve through sys.modules
    sys.modules["refmodels"] = pkg
    for sub in ("bilateralNN", "bnn_flow", "module_utils"):
        sys.modules["refmodels." + sub] = sys.modules["hplflownet_b200." + sub]
    spec = importlib.util.spec_from_file_location("refmodels." + file_name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return getattr(mod, class_name)


@pytest.mark.parametrize("file_name,class_name,fixture,args_cls", [
    ("HPLFlowNet.py", "HPLFlowNet", "model_frustum256.npz", ModelArgs),
    ("HPLFlowNet_shallow.py", "HPLFlowNetShallow", "model_shallow_frustum256.npz", ShallowArgs),
])
def test_unmodified_reference_model_runs_on_the_b200_modules(file_name, class_name, fixture, args_cls):
    from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
    cls = _reference_model_class(file_name, class_name)
    g = golden(fixture)
    args = args_cls()
    model = name_keyed_init_(cls(args), int(g["seed"])).cuda().eval()
    gen = GenerateDataUnsymmetric(args)                   # int64 tables: the reference's own data contract
    pc1, pc2, sf, gd = gen([g["pc1"], g["pc2"], np.zeros_like(g["pc1"])])
    # the caller's own stock layers (the bare nn.Conv1d ``conv4``) run cuDNN, which rounds fp32 operands to TF32 unless told
    # otherwise (the reference predates TF32); everything else is this package's kernels
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out = model(pc1[None], pc2[None], collate_batch1(gd))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert_close(out, g["output"], "output of the unmodified " + class_name)
    # and through autograd (training mode of the caller): finite gradients for every parameter that takes part
    model.train()
    out = model(pc1[None], pc2[None], collate_batch1(gd))
    out.square().mean().backward()
    n = 0
    for name, p in model.named_parameters():
        if p.grad is not None:
            assert bool(torch.isfinite(p.grad).all()), name
            n += 1
    assert n >= 60
