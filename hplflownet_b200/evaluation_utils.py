"""Scene-flow metrics of the reference's evaluation loop (``evaluation_utils.py:4-36``), on torch tensors so the
predicted flow never leaves the GPU: one fused reduction per call, one device-to-host read of the results.

``evaluate_3d(sf_pred, sf_gt)`` -> (EPE3D, acc3d_strict, acc3d_relax, outlier), inputs (N, 3) (any leading batch
dims are flattened); ``evaluate_2d(flow_pred, flow_gt)`` -> (EPE2D, acc2d), inputs (N, 2).  numpy arrays are
accepted too.  Thresholds and epsilons are the reference's.
"""
import numpy as np
import torch


def _t(x):
    x = torch.from_numpy(np.asarray(x)) if not torch.is_tensor(x) else x
    return x.reshape(-1, x.shape[-1]).double()


def evaluate_3d(sf_pred, sf_gt):
    p, g = _t(sf_pred), _t(sf_gt)
    l2 = torch.linalg.norm(g - p, dim=-1)
    rel = l2 / (torch.linalg.norm(g, dim=-1) + 1e-4)
    out = torch.stack([l2.mean(),
                       ((l2 < 0.05) | (rel < 0.05)).double().mean(),
                       ((l2 < 0.1) | (rel < 0.1)).double().mean(),
                       ((l2 > 0.3) | (rel > 0.1)).double().mean()])
    return tuple(out.tolist())


def evaluate_2d(flow_pred, flow_gt):
    p, g = _t(flow_pred), _t(flow_gt)
    epe = torch.linalg.norm(g - p, dim=-1)
    rel = epe / (torch.linalg.norm(g, dim=-1) + 1e-5)
    out = torch.stack([epe.mean(), ((epe < 3.0) | (rel < 0.05)).double().mean()])
    return tuple(out.tolist())
