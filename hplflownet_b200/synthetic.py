"""Seeded FlyingThings3D-shaped synthetic point clouds.

The frustum follows the reference's preprocessing camera
(``data_preprocess/flyingthings3d_utils.py:21-33``: f=1050, 960x540) and the
35 m depth threshold of ``configs/*.yaml``.  Used by tests and bench.py so that
the CUDA path, the oracle and the reference all see identical inputs.
"""
import numpy as np


def frustum_pair(n_points, seed, flow_sigma=0.3):
    """Returns (pc1, pc2) as (n_points, 3) fp32 arrays; pc2 = pc1 + N(0, flow_sigma^2)."""
    rs = np.random.RandomState(seed)
    z = rs.uniform(2.0, 35.0, n_points)
    u = rs.uniform(-0.457, 0.457, n_points)
    v = rs.uniform(-0.257, 0.257, n_points)
    pc1 = np.stack([u * z, v * z, z], axis=1).astype(np.float32)
    pc2 = (pc1 + rs.normal(0.0, flow_sigma, pc1.shape)).astype(np.float32)
    return pc1, pc2


def box_cloud(n_points, seed, half_extent=10.0):
    """Uniform box cloud (n_points, 3) fp32 -- a second distribution for parity tests."""
    rs = np.random.RandomState(seed)
    return rs.uniform(-half_extent, half_extent, (n_points, 3)).astype(np.float32)
