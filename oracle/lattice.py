"""oracle/lattice.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of ``lattice_oracle.c``: the scalar CPU restatement of the
reference's lattice build (``transforms/transforms.py:264-485``,
``GenerateDataUnsymmetric``).  Returns the same per-scale dicts (``:471-483``)
as numpy arrays so tests can compare the CUDA builder against them
bit-for-bit.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs
may import this module.  Parity status: pinned (see lattice_oracle.c header).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "lattice_oracle.c")
_SO = os.path.join(_HERE, "_build", "liblattice_oracle.so")

D = 3
D1 = 4


def build(force=False):
    """gcc the C restatement into oracle/_build/ (a few hundred ms)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -ffp-contract=off: the only fused operations are the explicit fmaf() calls.
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i64, f64, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
        L.hplo_constants.argtypes = [vp, vp]
        L.hplo_keys_barycentric.argtypes = [vp, i64, vp, vp, vp]
        L.hplo_key_range.argtypes = [vp, i64, vp, vp]
        L.hplo_count_vertices.argtypes = [vp, i64, vp, vp]
        L.hplo_count_vertices.restype = i64
        L.hplo_build_unsymmetric.argtypes = [i64, i64, i64, i64, i64] + [vp] * 15 + [i64, i64]
        L.hplo_build_unsymmetric.restype = ctypes.c_int
        L.hplo_next_points.argtypes = [vp, i64, f64, vp]
        L.hplo_neighbor_offsets.argtypes = [ctypes.c_int, vp]
        L.hplo_neighbor_offsets.restype = i64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def constants():
    """(elevate_mat (4,3) fp32, expected_std float) -- transforms.py:271-276."""
    e = np.empty((D1, D), np.float32)
    s = ctypes.c_double()
    lib().hplo_constants(_p(e), ctypes.byref(s))
    return e, s.value


def filter_size(radius):
    """transforms.py:355-356."""
    return (radius + 1) ** D1 - radius ** D1


def neighbor_offsets(radius):
    """(filter_size, 4) int64 in conv-weight order -- transforms.py:112-130,292-298."""
    out = np.empty((filter_size(radius), D1), np.int64)
    n = lib().hplo_neighbor_offsets(int(radius), _p(out))
    assert n == out.shape[0]
    return out


def keys_and_barycentric(pc):
    """pc (3, n) fp32 -> keys (4, n, 4) i64, bary (4, n) f32, el_minus_gr (4, n) f32.
    transforms.py:300-353."""
    pc = np.ascontiguousarray(pc, np.float32)
    assert pc.ndim == 2 and pc.shape[0] == D
    n = pc.shape[1]
    bary = np.empty((D1, n), np.float32)
    emg = np.empty((D1, n), np.float32)
    keys = np.empty((D1, n, D1), np.int64)
    lib().hplo_keys_barycentric(_p(pc), n, _p(bary), _p(emg), _p(keys))
    return keys, bary, emg


def key_range(keys1, keys2):
    """Per-coordinate (mins, maxs) over both clouds -- transforms.py:384-385."""
    mins = np.full(D1, np.iinfo(np.int64).max, np.int64)
    maxs = np.full(D1, np.iinfo(np.int64).min, np.int64)
    for k in (keys1, keys2):
        lib().hplo_key_range(_p(k), k.shape[1], _p(mins), _p(maxs))
    return mins, maxs


def next_points(last, scale):
    """(4, h) fp32 vertex coords -> (3, h) fp32 positions -- transforms.py:461-467."""
    last = np.ascontiguousarray(last, np.float32)
    out = np.empty((D, last.shape[1]), np.float32)
    lib().hplo_next_points(_p(last), last.shape[1], float(scale), _p(out))
    return out


def generate(pc1, pc2, scales_filter_map):
    """Restates GenerateDataUnsymmetric.__call__ (transforms.py:358-485).

    pc1, pc2: (N, 3) fp32 numpy.  Returns the list of per-scale dicts with the
    reference's 12 keys; tensors are numpy arrays with the reference's shapes
    and dtypes (placeholders for disabled tables are ``zeros(1)`` int64).
    """
    last1 = np.ascontiguousarray(np.asarray(pc1, np.float32).T).copy()
    last2 = np.ascontiguousarray(np.asarray(pc2, np.float32).T).copy()
    n1, n2 = last1.shape[1], last2.shape[1]
    out = []
    L = lib()
    for idx, (scale, bcn_r, corr_f_r, corr_c_r) in enumerate(scales_filter_map):
        last1[:3] *= np.float32(scale)          # :377-378
        last2[:3] *= np.float32(scale)
        k1, b1, e1 = keys_and_barycentric(last1)
        k2, b2, e2 = keys_and_barycentric(last2)
        mins, maxs = key_range(k1, k2)
        h1 = L.hplo_count_vertices(_p(k1), n1, _p(mins), _p(maxs))
        h2 = L.hplo_count_vertices(_p(k2), n2, _p(mins), _p(maxs))
        off1 = np.empty((D1, n1), np.int64)
        off2 = np.empty((D1, n2), np.int64)
        if bcn_r != -1:
            bfs = filter_size(bcn_r)
            blur1 = np.full((bfs, h1), -1, np.int64)
            blur2 = np.full((bfs, h2), -1, np.int64)
            boffs = neighbor_offsets(bcn_r)
        else:
            bfs, blur1, blur2, boffs = -1, None, None, None
        if corr_f_r != -1:
            cfs, ccs = filter_size(corr_f_r), filter_size(corr_c_r)
            corr1 = np.full((ccs, h1), -1, np.int64)
            corr2 = np.full((cfs, ccs, h1), -1, np.int64)
            cfo, cco = neighbor_offsets(corr_f_r), neighbor_offsets(corr_c_r)
        else:
            cfs, ccs, corr1, corr2, cfo, cco = -1, -1, None, None, None, None
        keep = idx != len(scales_filter_map) - 1
        lp1 = np.empty((D1, h1), np.float32) if keep else None
        lp2 = np.empty((D1, h2), np.float32) if keep else None
        rc = L.hplo_build_unsymmetric(n1, n2, bfs, cfs, ccs, _p(k1), _p(k2), _p(maxs), _p(mins),
                                      _p(off1), _p(off2), _p(boffs), _p(blur1), _p(blur2),
                                      _p(cfo), _p(cco), _p(corr1), _p(corr2), _p(lp1), _p(lp2),
                                      h1, h2)
        assert rc == 0
        ph = np.zeros(1, np.int64)                # :450-459 placeholders
        out.append({
            "pc1_barycentric": b1, "pc2_barycentric": b2,
            "pc1_el_minus_gr": e1, "pc2_el_minus_gr": e2,
            "pc1_lattice_offset": off1, "pc2_lattice_offset": off2,
            "pc1_blur_neighbors": blur1 if blur1 is not None else ph.copy(),
            "pc2_blur_neighbors": blur2 if blur2 is not None else ph.copy(),
            "pc1_corr_indices": corr1 if corr1 is not None else ph.copy(),
            "pc2_corr_indices": corr2 if corr2 is not None else ph.copy(),
            "pc1_hash_cnt": int(h1), "pc2_hash_cnt": int(h2),
        })
        if keep:
            last1, last2 = next_points(lp1, scale), next_points(lp2, scale)
            n1, n2 = h1, h2
    return out
