"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for floating-point values and gradients: 1e-5 relative
# fp32, measured per tensor as max|a-b| / max(max|ref|, eps)  (SURVEY §8c).
REL_TOL = 1e-5


def rel_err(got, ref):
    got = torch.as_tensor(got).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.numel() == 0:
        return 0.0
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def assert_close(got, ref, what="", tol=REL_TOL):
    e = rel_err(got, ref)
    assert e <= tol, "%s: rel err %.3e > %.1e" % (what, e, tol)


def golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def golden_files(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def t(a, device="cpu"):
    """numpy -> torch; fixtures narrow index tables to int32, the API takes int64."""
    a = np.asarray(a)
    if a.dtype == np.int32:
        a = a.astype(np.int64)
    return torch.from_numpy(a).to(device)


def state_from(g, device="cpu"):
    return {k[3:]: t(v, device) for k, v in g.items() if k.startswith("sd.")}


def grads_from(g):
    return {k[5:]: t(v) for k, v in g.items() if k.startswith("grad.")}


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
    return np.array_equal(a.astype(np.int64), b.astype(np.int64))
