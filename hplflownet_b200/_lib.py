"""ctypes binding of the C-ABI library (include/hplflownet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The library is built in-tree by ``python -m hplflownet_b200.build``
(``__graft_entry__.build()`` does it).
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HPL_LIB_PATH") or os.path.join(_PKG, "libhplflownet_b200.so")   # (override: timing experiments)

i64, i32, vp, cint = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int

# name -> argtypes; every function returns int (0 = ok).  Mirrors include/hplflownet_b200.h.
SIGNATURES = {
    "hpl_version": [],
    "hpl_sm_arch": [],
    "hpl_scatter_rows": [vp, vp, vp, cint, i64, i64, vp, i64, i64, vp, vp, vp],
    "hpl_normalize_rows": [vp, i64, i64, i64, vp, vp, vp],
    "hpl_gather_rows": [vp, i64, vp, vp, cint, vp, vp, i64, i64, i64, vp, vp],
    "hpl_blur_gemm": [vp, i64, i64, vp, cint, i64, i64, i64, i64, vp, vp, cint, vp, i64, cint, cint, vp],
    "hpl_blur_wgrad": [vp, i64, i64, vp, cint, i64, i64, i64, i64, vp, i64, vp, vp, vp],
    "hpl_absmax": [vp, i64, vp, vp],
    "hpl_blur_gemm_f16_workspace": [i64, i64, i64],
    "hpl_blur_gemm_f16": [vp, i64, i64, vp, cint, i64, i64, i64, i64, vp, vp, cint, vp, i64, cint, vp, vp, vp],
    "hpl_blur_wgrad_f16": [vp, i64, i64, vp, cint, i64, i64, i64, i64, vp, i64, vp, vp, vp, vp, vp],
    "hpl_blur_gemm_f16_amax": [vp, i64, i64, vp, cint, i64, i64, i64, i64, vp, i64, i64, i64, vp, cint, vp, i64, cint, vp, cint, vp, vp, vp],
    "hpl_normalize_rows_amax": [vp, i64, i64, i64, vp, vp, vp, vp],
    "hpl_cm_to_rows_amax": [vp, i64, i64, i64, vp, i64, vp, vp],
    "hpl_act_backward_stats": [vp, i64, vp, i64, i64, i64, cint, vp, vp, vp],
    "hpl_h16_bytes": [i64, i64],
    "hpl_h16_split": [vp, i64, i64, i64, vp, vp, vp],
    "hpl_blur_gemm_tma_workspace": [i64, i64, i64],
    "hpl_blur_gemm_tma": [vp, i64, vp, cint, i64, i64, i64, i64, vp, vp, cint, vp, i64, cint, vp, vp, vp],
    "hpl_column_sums": [vp, i64, i64, i64, vp, vp],
    "hpl_act_backward": [vp, i64, vp, i64, i64, i64, cint, vp],
    "hpl_transpose_table": [vp, cint, i64, i64, vp, i64, vp, vp],
    "hpl_cm_to_rows": [vp, i64, i64, i64, vp, i64, vp],
    "hpl_rows_to_cm": [vp, i64, i64, i64, vp, i64, vp],
    "hpl_channel_sums": [vp, i64, i64, vp, vp],
    "hpl_corr_gather": [vp, i64, vp, vp, i64, vp, cint, vp, cint, vp, i64, i64, i64, i64, i64, i64, i64, vp],
    "hpl_corr_scatter": [vp, i64, vp, vp, cint, vp, i64, vp, i64, i64, i64, i64, i64, i64, i64, vp],
    "hpl_lattice_init_range": [vp, vp],
    "hpl_lattice_points": [vp, i64, ctypes.c_float, vp, vp, vp, vp, vp, vp],
    "hpl_lattice_table_capacity": [i64],
    "hpl_lattice_scan_blocks": [i64],
    "hpl_lattice_insert": [vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, cint, vp, vp, vp],
    "hpl_lattice_neighbors": [vp, vp, i64, vp, vp, vp, i64, vp, i64, vp, cint, i64, vp],
    "hpl_lattice_corr_table": [vp, vp, i64, vp, vp, vp, i64, vp, i64, vp, i64, vp, cint, i64, vp],
    "hpl_lattice_next_points": [vp, i64, ctypes.c_float, vp, vp],
    "hpl_plan_tiles": [i64],
    "hpl_plan_umax": [],
    "hpl_plan_offset": [i64, cint],
    "hpl_plan_bytes": [i64],
    "hpl_plan_build": [vp, cint, i64, i64, i64, vp, vp, vp, vp, vp],
    "hpl_plan_order_workspace": [i64],
    "hpl_plan_order": [vp, cint, i64, i64, vp, cint, vp, vp, vp, vp],
    "hpl_h16b_bytes": [i64, i64],
    "hpl_h16b_split": [vp, i64, i64, i64, vp, vp, vp, vp],
    "hpl_h16b_split_ex": [vp, i64, i64, i64, vp, vp, vp, vp, i64, cint, vp, vp, vp, vp, cint, vp, vp],
    "hpl_h16b_splat_csr": [vp, i64, vp, i64, vp, vp, i64, i64, cint, vp, vp, vp, i64, cint, vp, vp, vp, vp, vp, vp],
    "hpl_conv5_workspace": [i64],
    "hpl_conv5_supported": [i64, i64, i64],
    "hpl_conv5": [vp, vp, i64, i64, i64, i64, vp, i64, i64, i64, vp, vp, cint, vp, i64, vp, cint, vp, vp, vp],
    "hpl_conv5_weights": [vp, i64, i64, i64, i64, i64, i64, vp, vp, vp],
    "hpl_wgrad5": [vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp],
    "hpl_fill_zero": [vp, i64, vp],
    "hpl_fill_i32": [vp, i64, i32, vp],
}


RETURNS_I64 = {"hpl_lattice_table_capacity", "hpl_lattice_scan_blocks",                "hpl_blur_gemm_f16_workspace", "hpl_h16_bytes", "hpl_blur_gemm_tma_workspace",
               "hpl_plan_tiles", "hpl_plan_umax", "hpl_plan_offset", "hpl_plan_bytes", "hpl_plan_order_workspace",
               "hpl_h16b_bytes", "hpl_conv5_workspace", "hpl_conv5_supported"}   # sizes, not status codes

# kernels enqueued per call (for bench.py's gpu_launches claim)
LAUNCHES = {
    "hpl_scatter_rows": 1, "hpl_normalize_rows": 2, "hpl_gather_rows": 1, "hpl_blur_gemm": 1, 
    "hpl_blur_wgrad": 2, "hpl_absmax": 1, "hpl_blur_gemm_f16_amax": 3, "hpl_normalize_rows_amax": 2, "hpl_cm_to_rows_amax": 1, "hpl_act_backward_stats": 1,
    "hpl_h16_split": 1, "hpl_blur_gemm_tma": 3, "hpl_blur_gemm_f16": 3, "hpl_blur_wgrad_f16": 2, "hpl_act_backward": 1, "hpl_transpose_table": 1, "hpl_cm_to_rows": 1,
    "hpl_rows_to_cm": 1, "hpl_channel_sums": 1, "hpl_fill_zero": 0, "hpl_fill_i32": 1,       # (hpl_fill_zero is a cudaMemsetAsync, not a kernel of this library)
   
    "hpl_corr_gather": 1, "hpl_corr_scatter": 1, "hpl_column_sums": 1,
    "hpl_lattice_init_range": 1, "hpl_lattice_points": 1, "hpl_lattice_insert": 6,
    "hpl_lattice_neighbors": 1, "hpl_lattice_corr_table": 1, "hpl_lattice_next_points": 1,
    "hpl_plan_build": 2, "hpl_h16b_split": 1, "hpl_h16b_split_ex": 1, "hpl_h16b_splat_csr": 1, "hpl_conv5": 3, "hpl_wgrad5": 1,
}
launch_count = 0


class HplError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "hplflownet_b200: %s is missing -- build it with `python -m hplflownet_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so is stale
            fn.argtypes = args
            fn.restype = i64 if name in RETURNS_I64 else cint
        _lib = lib
    return _lib


def call(name, *args):
    """Invoke an ABI entry point; raise HplError on a non-zero return code."""
    global launch_count
    launch_count += LAUNCHES.get(name, 0)
    rc = getattr(load(), name)(*args)
    if rc != 0:
        detail = "argument error" if rc < 0 else "CUDA error %d" % rc
        raise HplError("%s failed: %s (rc=%d)" % (name, detail, rc))
