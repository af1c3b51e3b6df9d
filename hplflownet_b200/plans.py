"""Tile plans of a neighbour table (csrc/plan.cu) -- the per-lattice precomputation behind contraction engine 5.

The reference hands the blur step nothing but ``blur_neighbors`` (models/bilateralNN.py:122-125); a plan is derived
from that tensor alone and cached while the tensor is unchanged (same storage, shape, dtype and version counter), so
a lattice that serves several layers and the forward / data-gradient / weight-gradient kernels is planned once.
"""
import weakref

import numpy as np
import torch

from . import _lib
from .transforms import neighbor_offsets

_offsets = {}
_cache = {}
RELAX_CHUNK = 48          # relaxation sweeps between convergence checks (one host read each)
RELAX_MAX = 4096


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object through
    # several Python layers: ~12 us per call, 270 calls per model forward)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _tap_offsets(filter_size, device):
    """(F, 4) int32 lattice offsets of the taps (transforms.py:112-130) for F = (r+1)^4 - r^4."""
    key = (filter_size, device)
    t = _offsets.get(key)
    if t is None:
        r = 1
        while (r + 1) ** 4 - r ** 4 < filter_size:
            r += 1
        if (r + 1) ** 4 - r ** 4 != filter_size:
            return None
        t = torch.from_numpy(np.ascontiguousarray(neighbor_offsets(r))).to(device)
        _offsets[key] = t
    return t


def mirror_taps(filter_size):
    """tap g -> the tap with the opposite offset (the offset set is closed under negation), or None."""
    r = 1
    while (r + 1) ** 4 - r ** 4 < filter_size:
        r += 1
    if (r + 1) ** 4 - r ** 4 != filter_size:
        return None
    offs = neighbor_offsets(r)
    index = {tuple(o): i for i, o in enumerate(offs.tolist())}
    return [index[tuple((-o).tolist())] for o in offs]


_mirrors = {}


def _mirror_tensor(filter_size, device):
    key = (filter_size, device)
    if key not in _mirrors:
        m = mirror_taps(filter_size)
        _mirrors[key] = torch.tensor(m, dtype=torch.int32, device=device) if m is not None else None
    return _mirrors[key]


class TilePlan:
    """Device buffers + host-side facts of one planned table."""

    def __init__(self, buf, n_rows, n_in_rows, filter_size, stats, order, sweeps):
        self.buf, self.n_rows, self.n_in_rows, self.filter_size = buf, n_rows, n_in_rows, filter_size
        self.max_uniq, self.overflow, self.sum_uniq, violations = stats
        self.order, self.sweeps = order, sweeps
        self.n_tiles = (n_rows + 127) // 128
        # True: nbr[mirror(f), nbr[f, v]] == v everywhere, the data gradient may reuse this plan with mirrored taps
        self.symmetric = violations == 0

    @property
    def usable(self):
        return self.overflow == 0

    def view(self, which):
        """torch view of one plan array (debugging / tests)."""
        L = _lib.load()
        lo, hi = L.hpl_plan_offset(self.n_rows, which), L.hpl_plan_offset(self.n_rows, which + 1)
        raw = self.buf[lo:hi]
        umax = L.hpl_plan_umax()
        if which == 0:
            return raw[:self.n_tiles * 512].view(torch.int32).view(self.n_tiles, 128)
        if which == 1:
            return raw[:self.n_tiles * 4].view(torch.int32)
        if which == 2:
            return raw[:self.n_tiles * umax * 4].view(torch.int32).view(self.n_tiles, umax)
        return raw[:self.n_tiles * 4096].view(torch.int16).view(self.n_tiles, 16, 128)


def spatial_order(nbr2, n_in_rows=None):
    """int32 permutation of the table's columns along a Morton curve of reconstructed lattice coordinates, and the
    number of relaxation sweeps used.  nbr2: (F, H) int32 / int64 CUDA tensor."""
    f, h = nbr2.shape
    offs = _tap_offsets(f, nbr2.device)
    if offs is None or h == 0:
        return None, 0
    L = _lib.load()
    ws = torch.empty(L.hpl_plan_order_workspace(h), dtype=torch.uint8, device=nbr2.device)
    order = torch.empty(h, dtype=torch.int32, device=nbr2.device)
    changed = torch.ones(1, dtype=torch.int32, device=nbr2.device)
    i64 = int(nbr2.dtype == torch.int64)
    # the relaxation restarts from scratch at every call: run it with a growing sweep count until the last sweep
    # changes nothing (one host read per attempt; plans are built once per lattice)
    sweeps = RELAX_CHUNK
    while True:
        _lib.call("hpl_plan_order", nbr2.data_ptr(), i64, f, h, offs.data_ptr(), sweeps, ws.data_ptr(),
                  order.data_ptr(), changed.data_ptr(), _stream())
        if int(changed.item()) == 0 or sweeps >= RELAX_MAX:
            break
        sweeps *= 2
    return order, sweeps


def build(nbr2, n_in_rows=None, order="spatial"):
    """Plan of the table nbr2 (F, n_rows) whose entries index rows [0, n_in_rows) (default n_rows: a same-lattice table)."""
    if not (nbr2.is_cuda and nbr2.dtype in (torch.int64, torch.int32) and nbr2.is_contiguous() and nbr2.dim() == 2):
        raise ValueError("nbr2 must be a contiguous CUDA (F, H) int64/int32 tensor")
    f, h = nbr2.shape
    n_in = h if n_in_rows is None else int(n_in_rows)
    L = _lib.load()
    sweeps = 0
    if isinstance(order, str):
        order, sweeps = spatial_order(nbr2) if order == "spatial" else (None, 0)
    buf = torch.empty(max(L.hpl_plan_bytes(h), 256), dtype=torch.uint8, device=nbr2.device)
    stats = torch.zeros(4, dtype=torch.int32, device=nbr2.device)
    mirror = _mirror_tensor(f, nbr2.device)
    _lib.call("hpl_plan_build", nbr2.data_ptr(), int(nbr2.dtype == torch.int64), f, h, n_in,
              order.data_ptr() if order is not None else None, mirror.data_ptr() if mirror is not None else None,
              buf.data_ptr(), stats.data_ptr(), _stream())
    s = stats.tolist()                                      # (one host read per planned table)
    return TilePlan(buf, h, n_in, f, tuple(s), order, sweeps)


# Planning a table costs a coordinate reconstruction (tens of relaxation sweeps), a radix sort and a host read of the
# statistics -- milliseconds, repaid over many kernel launches on a lattice that stays resident (a cached dataset, the
# forward / data-gradient / weight-gradient of every layer and step that use it), but not on a table that is seen once
# (a DataLoader that ships fresh tables every step).  So a table is planned when it is used for the SECOND time;
# ``PLAN_ON_FIRST_USE`` (or an explicit ``plans.prepare(blur_neighbors)``) plans immediately.
PLAN_ON_FIRST_USE = False


def _key(nbr2):
    return (nbr2.data_ptr(), tuple(nbr2.shape), nbr2.dtype, nbr2.device)


def _remember(nbr2, key, plan):
    base = nbr2._base if nbr2._base is not None else nbr2
    try:
        ref = weakref.ref(base, lambda _r, k=key: _cache.pop(k, None))
    except TypeError:
        return
    _cache[key] = (ref, nbr2._version, plan)


def plan_for(nbr2):
    """Cached plan of a same-lattice table (keyed by the tensor's storage / shape / dtype / version), or None while the
    table has been seen only once (see PLAN_ON_FIRST_USE)."""
    key = _key(nbr2)
    ent = _cache.get(key)
    fresh = ent is None or ent[0]() is None or ent[1] != nbr2._version
    if not fresh and ent[2] is not None:
        return ent[2]
    if fresh and not PLAN_ON_FIRST_USE:
        _remember(nbr2, key, None)
        return None
    plan = build(nbr2)
    _remember(nbr2, key, plan)
    return plan


def prepare(blur_neighbors):
    """Plan a table now (accepts the module-level (1, F, H) tensor or its (F, H) view) and return the plan."""
    nbr2 = blur_neighbors[0] if blur_neighbors.dim() == 3 else blur_neighbors
    nbr2 = nbr2.contiguous()
    key = _key(nbr2)
    ent = _cache.get(key)
    if ent is not None and ent[0]() is not None and ent[1] == nbr2._version and ent[2] is not None:
        return ent[2]
    plan = build(nbr2)
    _remember(nbr2, key, plan)
    return plan


def clear():
    _cache.clear()
    _splat_cache.clear()


# ------------------------------------------------------------------------------------------ splat plans
# The splat's scatter pattern -- lattice_offset (d+1, N): point -> its d+1 lattice rows -- is a per-lattice constant like the
# neighbour table.  Sorted by row it turns the splat (and the backward of the slice over the same tables) into a gather:
# csr_ptr[v] .. csr_ptr[v+1] index the contributions (point, remainder) of lattice row v, in a fixed order (ascending
# remainder, then point), so the fp32 sums are reproducible and no atomics or accumulators are needed
# (hpl_h16b_splat_csr).  Built with a stable device sort on the tensor's SECOND use (or plans.prepare_splat), cached like the
# tile plans.
# Off by default: measured on cfg2 x 32 the gather (166 us: latency-bound loops over 4.3 contributions per row on average but
# up to 220 on dense rows) plus the transposition of the features (50 us) is twice the cost of the RED splat + split (74 +
# 29 us).  Worth it where bitwise reproducible splats matter: plans.SPLAT_PLANS = True.
SPLAT_PLANS = False


class SplatPlan:
    def __init__(self, ptr, ent, n_rows, n_points):
        self.ptr, self.ent, self.n_rows, self.n_points = ptr, ent, n_rows, n_points


_splat_cache = {}


def build_splat(off2, n_rows):
    """off2: (d+1, N) int32 / int64 CUDA tensor of lattice rows (entries outside [0, n_rows) are dropped, as in the kernels)."""
    d1, n = off2.shape
    if n >= (1 << 30):
        return None
    keys = off2.reshape(-1).to(torch.int64)
    keys = torch.where((keys >= 0) & (keys < n_rows), keys, torch.full_like(keys, n_rows))
    order = torch.argsort(keys, stable=True)                              # position r * N + point, grouped by row
    counts = torch.bincount(keys, minlength=n_rows + 1)[:n_rows]
    ptr = torch.zeros(n_rows + 1, dtype=torch.int32, device=off2.device)
    ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    ent = ((order % n) | ((order // n) << 30)).to(torch.int32)            # point | remainder << 30 (dropped entries trail, unused)
    return SplatPlan(ptr.contiguous(), ent.contiguous(), n_rows, n)


def _splat_key(off2, n_rows):
    return (off2.data_ptr(), tuple(off2.shape), off2.dtype, off2.device, int(n_rows))


def splat_plan_for(off2, n_rows):
    """Cached splat plan of an offset tensor, or None while it has been seen only once (see PLAN_ON_FIRST_USE)."""
    if not SPLAT_PLANS:
        return None
    key = _splat_key(off2, n_rows)
    ent = _splat_cache.get(key)
    fresh = ent is None or ent[0]() is None or ent[1] != off2._version
    if not fresh and ent[2] is not None:
        return ent[2]
    plan = None
    if not fresh or PLAN_ON_FIRST_USE:
        plan = build_splat(off2, n_rows)
    base = off2._base if off2._base is not None else off2
    try:
        ref = weakref.ref(base, lambda _r, k=key: _splat_cache.pop(k, None))
    except TypeError:
        return plan
    _splat_cache[key] = (ref, off2._version, plan)
    return plan


def prepare_splat(lattice_offset, n_rows):
    """Plan a lattice_offset tensor now (accepts the module-level (1, d+1, N) tensor or its (d+1, N) view)."""
    off2 = lattice_offset[0] if lattice_offset.dim() == 3 else lattice_offset
    off2 = off2.contiguous()
    key = _splat_key(off2, n_rows)
    ent = _splat_cache.get(key)
    if ent is not None and ent[0]() is not None and ent[1] == off2._version and ent[2] is not None:
        return ent[2]
    plan = build_splat(off2, n_rows)
    base = off2._base if off2._base is not None else off2
    try:
        ref = weakref.ref(base, lambda _r, k=key: _splat_cache.pop(k, None))
        _splat_cache[key] = (ref, off2._version, plan)
    except TypeError:
        pass
    return plan
