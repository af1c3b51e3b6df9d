"""hplflownet_b200 -- B200-native (sm_100a) bilateral convolution layers for HPLFlowNet.

Drop-in replacements for the reference's hot path:
``BilateralConvFlex`` (models/bilateralNN.py), ``BilateralCorrelationFlex`` (models/bnn_flow.py)
and the lattice builder ``GenerateDataUnsymmetric`` (transforms/transforms.py), all backed by
hand-written CUDA behind the C ABI in include/hplflownet_b200.h.  No CPU fallback.
"""
from .bilateralNN import BilateralConvFlex, SparseSum, sparse_sum  # noqa: F401
from .bnn_flow import BilateralCorrelationFlex  # noqa: F401
from .module_utils import Conv1dReLU, Conv2dReLU, Conv3dReLU  # noqa: F401

__version__ = "0.1.0"
