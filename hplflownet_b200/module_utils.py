"""Parameter containers with the reference's ``state_dict`` layout.

The reference wraps every learned convolution as ``ConvNd -> (Leaky)ReLU`` inside a
``composed_module`` Sequential (models/module_utils.py:9-59); checkpoints are loaded with
``strict=True`` (main.py:122), so the attribute path ``<name>.composed_module.0.{weight,bias}``
is part of the drop-in contract.  These classes keep that layout and the
default initialisation of ``nn.ConvNd``.  The BCL / correlation modules read ``.weight`` / ``.bias``
directly and run the fused CUDA path; CALLING a ``Conv1dReLU`` with kernel size 1 on a CUDA fp32 tensor
(what the unmodified models/HPLFlowNet.py does for ``conv1`` / ``conv2`` / ``conv3``) runs the same
hand-written fp32-accurate GEMM (``pointwise.py``) -- the stock cuDNN convolution would use TF32 by
default on this hardware (1e-3 relative).  Everything else (CPU, float64, other kernel sizes) runs the
stock op.
"""
import torch.nn as nn

__all__ = ["Conv1dReLU", "Conv2dReLU", "Conv3dReLU", "LEAKY_RATE"]

LEAKY_RATE = 0.1


class _ConvAct(nn.Module):
    conv_cls = None

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, use_leaky=False,
                 bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.use_leaky = use_leaky
        act = nn.LeakyReLU(LEAKY_RATE, inplace=True) if use_leaky else nn.ReLU(inplace=True)
        conv = self.conv_cls(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                             padding=padding, bias=bias)
        self.composed_module = nn.Sequential(conv, act)

    @property
    def conv(self):
        return self.composed_module[0]

    def forward(self, x):
        return self.composed_module(x)


class Conv1dReLU(_ConvAct):
    conv_cls = nn.Conv1d

    def forward(self, x):
        conv = self.conv
        if (x.is_cuda and x.dtype == conv.weight.dtype and str(x.dtype) == "torch.float32" and x.dim() == 3 and x.size(0) == 1
                and conv.kernel_size == (1,) and conv.stride == (1,) and conv.padding == (0,) and conv.bias is not None):
            from .pointwise import pointwise_stack
            return pointwise_stack((self,), x)
        return self.composed_module(x)


class Conv2dReLU(_ConvAct):
    conv_cls = nn.Conv2d


class Conv3dReLU(_ConvAct):
    conv_cls = nn.Conv3d
