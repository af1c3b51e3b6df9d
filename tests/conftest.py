import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _plan_tables_on_first_use(request):
    """The parity tests call every module once per table; plan on first use so that engine 5 (tile plans) is what they
    exercise wherever it applies.  (Production default: a table is planned when it is used for the second time.)"""
    if "gpu" not in request.keywords:
        yield
        return
    from hplflownet_b200 import plans
    old = plans.PLAN_ON_FIRST_USE
    plans.PLAN_ON_FIRST_USE = True
    yield
    plans.PLAN_ON_FIRST_USE = old


def pytest_terminal_summary(terminalreporter):
    """Which gradient tensors needed the activation-kink bound of tests/_util.py:assert_close_grad, and why."""
    try:
        from tests._util import FALLBACKS
    except Exception:
        return
    tr = terminalreporter
    tr.write_line("gradient tensors checked with the activation-kink bound instead of 1e-5: %d" % len(FALLBACKS))
    for what, e_max, e_l2, amb in FALLBACKS:
        tr.write_line("  %s: max %.2e, L2 %.2e, oracle units near a kink: %s" % (what, e_max, e_l2, amb))
