"""Tensor-level wrappers over the C ABI (include/hplflownet_b200.h).

PyTorch is used for device memory and streams only: every function checks its arguments,
allocates outputs with torch, and enqueues the hand-written CUDA kernels on the current stream.
Lattice values are vertex-major ``(rows, ld)`` fp32 tensors with ``ld % 4 == 0``.
"""
import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

# Contraction engines (all hand-written sm_100a kernels with fp32-level accuracy; HPL_GEMM_PRECISION overrides):
#   5 (default) = engine 5 for the gathered first layer of a stack whenever its table has a usable tile plan
#                 (csrc/plan.cu + gemm_plan.cu: every distinct neighbour row of a 128-vertex tile is staged once,
#                 pre-split h16b operands, forward / data gradient / weight gradient on the same plan), engine 2
#                 for everything else (1x1 layers, wide layers, tables without locality);
#   2 = tcgen05 with a scaled FP16 hi/lo split done in registers per tap ("3xFP16", csrc/gemm_tc16.cu);
#   4 = tcgen05 3xFP16 with operands pre-split once in HBM and moved by cp.async / TMA row gathers
#       (csrc/gemm_tma.cu; forward / data gradient, dW uses engine 2);
#   0 = fp32 FMA on CUDA cores (parity anchor; also takes shapes with C % 4 != 0).
import os as _os
DEFAULT_PRECISION = int(_os.environ.get("HPL_GEMM_PRECISION", "5"))
_tc_workspace = {}


# Weight images (pre-split, pre-laid-out fp16 hi/lo copies of a conv weight, built by two small kernels per call) are
# reused across calls only while the caller vouches that parameters do not change: inside ``weight_cache_scope()``
# (evaluation loops; ``train.train_step`` wraps one optimisation step in it and the images are dropped when the scope
# ends, i.e. right after the optimizer step), or globally with ``WEIGHT_CACHE = "always"``.  The parameter's version
# counter and storage pointer are checked as well, but they are not sufficient on their own: ``param.data`` writes
# (``init.normal_(m.weight.data)`` in the reference's init_weights, EMA swaps, manual clipping) change neither, which
# is why there is no implicit caching outside a scope.  Only module parameters are cached (the owner travels as
# ``w._hpl_owner = (parameter, role)``); entries are inserted after the kernels that fill them were enqueued without
# error, and a weakref finalizer on the parameter removes them when it dies.
WEIGHT_CACHE = True            # False: never; True: inside weight_cache_scope(); "always": every call
_weight_images = {}
_scope_depth = 0


class weight_cache_scope:
    """Within the scope, weight images are reused across calls (the caller promises not to modify parameters inside
    it); every image is dropped on exit."""

    def __enter__(self):
        global _scope_depth
        _scope_depth += 1
        return self

    def __exit__(self, *a):
        global _scope_depth
        _scope_depth -= 1
        if _scope_depth == 0:
            _weight_images.clear()


def _cache_allowed(param):
    if not WEIGHT_CACHE:
        return False
    if WEIGHT_CACHE == "always" or _scope_depth > 0:
        return True
    return False


def _cached_workspace(w, nbytes, wide_rows):
    """(workspace tensor, valid flag, commit) for the weight view w; valid = the image inside is current;
    commit() must be called once the kernels that build the image have been enqueued without error."""
    owner = getattr(w, "_hpl_owner", None)
    if owner is None or not _cache_allowed(owner[0]):
        return _workspace(w.device, nbytes), 0, _no_commit
    import weakref
    param, role = owner
    key = (id(param), role, wide_rows, _stream_of(w.device))
    sig = (tuple(w.shape), tuple(w.stride()), w.data_ptr())
    ent = _weight_images.get(key)
    if ent is not None and ent[0]() is param and ent[1] == param._version and ent[2].numel() * 4 >= nbytes and ent[3] == sig:
        return ent[2], 1, _no_commit
    _weight_images.pop(key, None)
    ws = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=w.device)

    def commit():
        try:
            ref = weakref.ref(param, lambda _r, k=key: _weight_images.pop(k, None))
        except TypeError:
            return
        _weight_images[key] = (ref, param._version, ws, sig)
    return ws, 0, commit


def _no_commit():
    pass


def invalidate_weight_cache():
    """Drop every cached weight image (after an in-place update the version counter does not see)."""
    _weight_images.clear()


def with_owner(view, param, role):
    """Tag a weight view with the parameter it was derived from (enables the weight-image cache)."""
    view._hpl_owner = (param, role)
    return view


def _workspace(device, nbytes):
    # rewritten by every call that uses it, so it is private to one (device, stream)
    key = (device.index, _stream_of(device), nbytes)
    ws = _tc_workspace.get(key)
    if ws is None:
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        _tc_workspace[key] = ws
    return ws


# bench.py sets this to a list to collect (tag, start_event, end_event) around the contraction
# kernels (CUDA events on the launching stream); None = no instrumentation.
PROFILE_GEMM = None
PROFILE_ROWS = False            # also time the splat / slice kernels (bench.py enables it for a separate, untimed pass)


class _timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        self.on = PROFILE_GEMM is not None and (PROFILE_ROWS or self.tag not in ("scatter", "gather", "corr_gather", "corr_scatter"))
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            PROFILE_GEMM.append((self.tag, self.e0, self.e1))


def _stream_of(device):
    """raw handle of the current stream of `device`"""
    if not isinstance(device, torch.device):
        device = torch.device(device)
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch._C._cuda_getDevice())


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object through
    # several Python layers: ~12 us per call, 270 calls per model forward)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _f32(x, name):
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA float32 tensor" % name)
    return x


def _idx(x, name):
    if not (x.is_cuda and x.dtype in (torch.int64, torch.int32) and x.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA int64/int32 tensor" % name)
    return x, int(x.dtype == torch.int64)


def round4(c):
    return (int(c) + 3) // 4 * 4


def alloc_rows(n_rows, channels, device, zero=False):
    """(n_rows, round4(channels)) fp32; pad columns are always zero."""
    ld = round4(channels)
    if zero or ld != channels:
        return torch.zeros((n_rows, ld), dtype=torch.float32, device=device)
    return torch.empty((n_rows, ld), dtype=torch.float32, device=device)


# Zero pool: splat accumulators (rows that fp32 RED adds into) must start at zero.  A buffer whose last consumer zeroes it
# again (hpl_h16b_split_ex with dispose = 2) goes back to the pool and the next splat of the same shape reuses it without
# a memset (62 MB per cfg2 x 32 step and direction).  Keyed by (device, stream, rows, ld).
_zero_pool = {}


def zero_rows(n_rows, channels, device):
    """(n_rows, round4(channels)) fp32, all zero."""
    ld = round4(channels)
    key = (device.index, _stream_of(device), n_rows, ld)
    free = _zero_pool.get(key)
    if free:
        join_side(device)
        return free.pop()
    return torch.zeros((n_rows, ld), dtype=torch.float32, device=device)


def release_zero_rows(t):
    """Give back a buffer that is all zero again (its consumer ran with dispose = 2 on the current stream)."""
    key = (t.device.index, _stream_of(t.device), t.size(0), t.size(1))
    free = _zero_pool.setdefault(key, [])
    if len(free) < 4:
        free.append(t)


# Large accumulators are not re-zeroed by their consumer (the zero stores were 40 % of the fused split pass's traffic, which
# runs at DRAM speed) but by a fill on a SIDE stream, concurrently with the contraction kernel that follows on the main
# stream (issue / latency bound, it leaves the DRAM bandwidth unused).  The side stream is forked from and joined back into
# the current stream (join_side: end of the module's forward / backward, and before any buffer leaves the pool), so the
# pattern is also legal inside a CUDA-graph capture, where it becomes a parallel branch.
SIDE_ZERO_MIN_BYTES = 16 << 20
_side_streams = {}
_side_dirty = {}


def dispose_mode(t):
    """dispose argument of hpl_h16b_split_ex for an accumulator that goes back to the zero pool: 2 = the kernel zeroes it,
    0 = recycle_rows will (large buffers)."""
    return 0 if t.numel() * 4 >= SIDE_ZERO_MIN_BYTES else 2


_stream_objs = {}


def _cur_stream_obj(device):
    """torch Stream object of the current stream (cached per raw handle: building one costs ~10 us)."""
    raw = _stream_of(device)
    key = (device.index, raw)
    obj = _stream_objs.get(key)
    if obj is None:
        obj = _stream_objs[key] = torch.cuda.current_stream(device)
    return obj, key


def recycle_rows(t):
    """Give back an accumulator whose consumer has been enqueued on the current stream: zeroed here on the side stream
    (large buffers, see above) or already zeroed by the consumer (dispose_mode(t) == 2)."""
    pool_key = (t.device.index, _stream_of(t.device), t.size(0), t.size(1))
    if len(_zero_pool.get(pool_key, ())) >= 4:
        return                       # the pool is full: the buffer is simply dropped (and must not be touched on another stream)
    if t.numel() * 4 >= SIDE_ZERO_MIN_BYTES:
        # (the buffer stays referenced by the pool, so the allocator cannot reuse its memory while the side-stream fill is
        # pending; join_side orders the fill before the next use)
        cur, key = _cur_stream_obj(t.device)
        side = _side_streams.get(key)
        if side is None:
            side = _side_streams[key] = torch.cuda.Stream(t.device)
        side.wait_stream(cur)
        _lib.call("hpl_fill_zero", t.data_ptr(), t.numel() * 4, side.cuda_stream)
        _side_dirty[key] = side
    release_zero_rows(t)


def fork_side(device):
    """Raw handle of the side stream, forked from the current stream here (everything enqueued so far is ordered before
    what the caller launches on it); join_side must follow before the results are used."""
    cur, key = _cur_stream_obj(device)
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device)
    side.wait_stream(cur)
    _side_dirty[key] = side
    return side.cuda_stream


def join_side(device):
    """Order the side stream's fills before whatever follows on the current stream."""
    if _side_dirty:
        cur, key = _cur_stream_obj(device)
        side = _side_dirty.pop(key, None)
        if side is not None:
            cur.wait_stream(side)


def scatter_rows(x, bary, off, n_rows, want_wsum, in_amax=None, rows=None, wsum=None):
    """x (C, N), bary (4, N), off (4, N) -> rows (n_rows, ld) [, wsum (n_rows)].  in_amax: zeroed slot that receives
    max|x| (a bound of the normalised splat's magnitude).  rows / wsum: zeroed accumulators to use (ops.zero_rows,
    ops.zero_arena)."""
    _f32(x, "x"); _f32(bary, "bary")
    off, i64 = _idx(off, "off")
    c, n = x.shape
    if rows is None:
        rows = alloc_rows(n_rows, c, x.device, zero=True)
    if want_wsum and wsum is None:
        wsum = small_zeros(n_rows, torch.float32, x.device)
    with _timed("scatter"):
        _lib.call("hpl_scatter_rows", x.data_ptr(), bary.data_ptr(), off.data_ptr(), i64, n, c,
                  rows.data_ptr(), rows.stride(0), n_rows, wsum.data_ptr() if want_wsum else None,
                  in_amax.data_ptr() if in_amax is not None else None, _stream())
    return rows, wsum


def zero_arena(device, spec):
    """One zero-filled allocation carved into named tensors: spec = [(name, numel, dtype)], every piece 16-byte aligned.
    A step needs a handful of small zeroed outputs (statistic slots, weight sums, dw, bias gradients); one fill kernel
    instead of one per tensor keeps them off the launch path."""
    offs, total = [], 0
    for _, n, _ in spec:
        offs.append(total)
        total += (int(n) + 3) // 4 * 4
    buf = small_zeros(max(total, 4), torch.float32, device)
    return {name: buf[o:o + int(n)].view(dt) for (name, n, dt), o in zip(spec, offs)}


# Small zeroed buffers (statistic slots, bias gradients, weight sums): a training step of the full model asked torch for
# ~2000 of them -- one fill kernel and ~10 us of launch path each.  They are carved from a zeroed chunk instead (a bump
# pointer per device and stream; a chunk is never reused, its views keep it alive).  Not during CUDA-graph capture: there
# every buffer must be zeroed by a node of the graph that uses it.
_ZERO_CHUNK_WORDS = 1 << 18                     # 1 MB
_SMALL_ZERO_MAX = 1 << 14                       # words; larger requests go to torch.zeros
_zero_chunks = {}


def small_zeros(n, dtype, device):
    """n zeroed elements of a 4-byte dtype (16-byte aligned)."""
    n = int(n)
    if n > _SMALL_ZERO_MAX or torch.cuda.is_current_stream_capturing():
        return torch.zeros(n, dtype=dtype, device=device)
    if not isinstance(device, torch.device):
        device = torch.device(device)
    key = (device.index, _stream_of(device))
    st = _zero_chunks.get(key)
    need = (max(n, 1) + 3) // 4 * 4
    if st is None or st[1] + need > _ZERO_CHUNK_WORDS:
        st = [torch.zeros(_ZERO_CHUNK_WORDS, dtype=torch.int32, device=device), 0]
        _zero_chunks[key] = st
    v = st[0][st[1]:st[1] + n]
    st[1] += need
    return v if dtype == torch.int32 else v.view(dtype)


def amax_slots(device, n):
    """n zeroed device scalars for the fused max|x| statistics (kernels RED.MAX into them); index with [i:i+1]."""
    return small_zeros(n, torch.int32, device)


def fused_stats():
    """True when the default engine takes its operand scales from producer-fused statistics (engines 2 / 5)."""
    return DEFAULT_PRECISION in (2, 5)


def engine5_enabled():
    return DEFAULT_PRECISION == 5


def normalize_rows_(rows, channels, wsum, amax=None):
    """In place: rows *= 1/(wsum+1e-5); wsum becomes the reciprocal (returned).  amax: zeroed slot that receives
    max|rows| of the result (saves the separate hpl_absmax pass)."""
    if amax is not None:
        _lib.call("hpl_normalize_rows_amax", rows.data_ptr(), rows.stride(0), rows.size(0), channels,
                  wsum.data_ptr(), wsum.data_ptr(), amax.data_ptr(), _stream())
    else:
        _lib.call("hpl_normalize_rows", rows.data_ptr(), rows.stride(0), rows.size(0), channels,
                  wsum.data_ptr(), wsum.data_ptr(), _stream())
    return wsum


def reciprocal_(wsum):
    """In place: wsum <- 1/(wsum+1e-5), to be applied by the contraction kernels as ``row_scale``."""
    _lib.call("hpl_normalize_rows", None, 0, wsum.numel(), 0, wsum.data_ptr(), wsum.data_ptr(), _stream())
    return wsum


def absmax(t):
    """Device scalar (int32 tensor holding fp32 bits) = max |t| over the whole buffer (pads are zero)."""
    _f32(t, "t")
    out = torch.empty(1, dtype=torch.int32, device=t.device)
    _lib.call("hpl_absmax", t.data_ptr(), t.numel(), out.data_ptr(), _stream())
    return out


def h16_split(x, channels, amax):
    """fp16 hi/lo row image for the TMA-gathered contraction (include/hplflownet_b200.h: hpl_h16_split)."""
    _f32(x, "x")
    n = x.size(0)
    buf = torch.empty(_lib.load().hpl_h16_bytes(n, channels), dtype=torch.uint8, device=x.device)
    _lib.call("hpl_h16_split", x.data_ptr(), x.stride(0), n, channels, amax.data_ptr(), buf.data_ptr(), _stream())
    return buf


def _dense_permutation(w):
    """True when the 3-D tensor w covers one dense buffer of w.numel() elements exactly once (a permuted view of a
    contiguous tensor, possibly with broadcast-free size-1 dims) starting at its own data pointer."""
    dims = sorted(((st, sz) for st, sz in zip(w.stride(), w.shape) if sz > 1))
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


def gather_rows(rows, channels, bary, off, scale=None, bias=None):
    """rows (H, ld) -> y (C, N)."""
    _f32(rows, "rows"); _f32(bary, "bary")
    off, i64 = _idx(off, "off")
    n = bary.size(-1)
    y = torch.empty((channels, n), dtype=torch.float32, device=rows.device)
    with _timed("gather"):
        _lib.call("hpl_gather_rows", rows.data_ptr(), rows.stride(0), bary.data_ptr(), off.data_ptr(), i64,
                  scale.data_ptr() if scale is not None else None,
                  bias.data_ptr() if bias is not None else None, n, channels, rows.size(0), y.data_ptr(), _stream())
    return y


def blur_gemm(x, c_in, nbr, n_out_rows, w, bias, act, out=None, out_channel_major=False, precision=None,
              tag="fwd", row_scale=None, x_amax=None, x16=None, out_amax=None):
    """out[v] = act(bias + sum_f x[nbr[f, v]] @ w[f]);  w (F, C, Co) -- any dense permutation of a contiguous buffer
    (e.g. a view of the conv weight): engine 2 reads it strided, the other engines get a contiguous copy.
    out_amax: zeroed slot that receives max|out|."""
    _f32(x, "x")
    if not (w.is_cuda and w.dtype == torch.float32):
        raise ValueError("w must be a CUDA float32 tensor")
    if row_scale is not None:
        raise ValueError("row_scale went away with the 3xTF32 engine: normalise the rows (ops.normalize_rows_)")
    f, c, co = w.shape
    assert c == c_in
    if nbr is not None:
        nbr, i64 = _idx(nbr, "nbr")
        assert nbr.shape[-2] == f and nbr.shape[-1] == n_out_rows
        nbr_ptr = nbr.data_ptr()
    else:
        assert f == 1
        nbr_ptr, i64 = None, 0
    if out is None:
        out = (torch.empty((co, n_out_rows), dtype=torch.float32, device=x.device) if out_channel_major
               else alloc_rows(n_out_rows, co, x.device))
    if precision is None:
        precision = DEFAULT_PRECISION
    if precision == 5:
        precision = 2
    bias_ptr = bias.data_ptr() if bias is not None else None
    if precision not in (0, 2, 4):
        raise ValueError("unknown contraction engine %r" % (precision,))
    if c % 4 != 0:
        precision = 0
    if precision != 2:
        w = w.contiguous()
    if precision == 4:
        ws = _workspace(x.device, _lib.load().hpl_blur_gemm_tma_workspace(f, c, co))
        if x16 is None:
            x_amax = absmax(x)
            x16 = h16_split(x, c, x_amax)
        with _timed(tag):
            _lib.call("hpl_blur_gemm_tma", x16.data_ptr(), x.size(0), nbr_ptr, i64, f, n_out_rows, c, co,
                      w.data_ptr(), bias_ptr, act, out.data_ptr(), out.stride(0), int(out_channel_major),
                      ws.data_ptr(), x_amax.data_ptr(), _stream())
    elif precision == 2:
        if x_amax is None:
            x_amax = absmax(x)
        if not _dense_permutation(w):
            w = w.contiguous()
        ws, ws_valid, commit = _cached_workspace(w, _lib.load().hpl_blur_gemm_f16_workspace(f, c, co), n_out_rows >= 8192)
        with _timed(tag):
            _lib.call("hpl_blur_gemm_f16_amax", x.data_ptr(), x.stride(0), x.size(0), nbr_ptr, i64, f, n_out_rows, c, co,
                      w.data_ptr(), w.stride(0), w.stride(1), w.stride(2),
                      bias_ptr, act, out.data_ptr(), out.stride(0), int(out_channel_major),
                      ws.data_ptr(), ws_valid, x_amax.data_ptr(), out_amax.data_ptr() if out_amax is not None else None, _stream())
        commit()
        out_amax = None
    else:
        with _timed(tag):
            _lib.call("hpl_blur_gemm", x.data_ptr(), x.stride(0), x.size(0), nbr_ptr, i64, f, n_out_rows, c, co,
                      w.data_ptr(), bias_ptr, act, out.data_ptr(), out.stride(0), int(out_channel_major), 0,
                      _stream())
    if out_amax is not None:                 # an engine without the fused epilogue statistic ran
        _lib.call("hpl_absmax", out.data_ptr(), out.numel(), out_amax.data_ptr(), _stream())
    return out


def blur_wgrad(x, c_in, nbr, n_out_rows, dz, c_out, filter_size, want_db=True, precision=None, row_scale=None,
               x_amax=None, dz_amax=None, x16=None, dz16=None):
    """Returns dw (F, C, Co), db (Co)."""
    _f32(x, "x"); _f32(dz, "dz")
    if row_scale is not None:
        raise ValueError("row_scale went away with the 3xTF32 engine")
    if nbr is not None:
        nbr, i64 = _idx(nbr, "nbr")
        nbr_ptr = nbr.data_ptr()
    else:
        nbr_ptr, i64 = None, 0
    dw = torch.zeros((filter_size, c_in, c_out), dtype=torch.float32, device=x.device)
    db = small_zeros(c_out, torch.float32, x.device) if want_db else None
    if precision is None:
        precision = DEFAULT_PRECISION
    if precision in (4, 5):                  # engines 4 / 5 cover other roles; dW of these layers is register-staged
        precision = 2
    if precision == 2 and c_in % 4 == 0:
        if x_amax is None:
            x_amax = absmax(x)
        if dz_amax is None:
            dz_amax = absmax(dz)
        with _timed("wgrad"):
            _lib.call("hpl_blur_wgrad_f16", x.data_ptr(), x.stride(0), x.size(0), nbr_ptr, i64, filter_size,
                      n_out_rows, c_in, c_out, dz.data_ptr(), dz.stride(0), dw.data_ptr(),
                      db.data_ptr() if want_db else None, x_amax.data_ptr(), dz_amax.data_ptr(), _stream())
        return dw, db
    with _timed("wgrad"):
        _lib.call("hpl_blur_wgrad", x.data_ptr(), x.stride(0), x.size(0), nbr_ptr, i64, filter_size, n_out_rows,
                  c_in, c_out, dz.data_ptr(), dz.stride(0), dw.data_ptr(), db.data_ptr() if want_db else None,
                  _stream())
    return dw, db


def column_sums_(rows, channels, sums):
    """sums[c] += sum_v rows[v, c] (sums zeroed by the caller)."""
    _lib.call("hpl_column_sums", rows.data_ptr(), rows.stride(0), rows.size(0), channels, sums.data_ptr(), _stream())
    return sums


def act_backward_stats_(dz, y, channels, act, amax=None, colsum=None):
    """One pass: dz *= act'(y) in place, max|dz| -> amax (zeroed slot), column sums of dz += colsum (zeroed (Co,))."""
    _lib.call("hpl_act_backward_stats", dz.data_ptr(), dz.stride(0), y.data_ptr() if act != ACT_NONE else None,
              y.stride(0) if act != ACT_NONE else 0, dz.size(0), channels, act,
              amax.data_ptr() if amax is not None else None, colsum.data_ptr() if colsum is not None else None, _stream())
    return dz


def act_backward_(dz, y, channels, act):
    if act != ACT_NONE:
        _lib.call("hpl_act_backward", dz.data_ptr(), dz.stride(0), y.data_ptr(), y.stride(0), dz.size(0),
                  channels, act, _stream())
    return dz


def transpose_table(tbl, n_src_rows):
    """tbl (F, n) -> int32 (F, n_src_rows) with t[f, tbl[f, v]] = v, -1 elsewhere."""
    tbl, i64 = _idx(tbl, "tbl")
    f, n = tbl.shape[-2], tbl.shape[-1]
    out = torch.empty((f, n_src_rows), dtype=torch.int32, device=tbl.device)
    _lib.call("hpl_fill_i32", out.data_ptr(), out.numel(), -1, _stream())
    _lib.call("hpl_transpose_table", tbl.data_ptr(), i64, f, n, out.data_ptr(), n_src_rows, None, _stream())
    return out


def cm_to_rows(cm, amax=None):
    """(C, n) channel-major -> (n, ld) vertex-major.  amax: zeroed slot that receives max|cm|."""
    _f32(cm, "cm")
    c, n = cm.shape
    rows = torch.empty((n, round4(c)), dtype=torch.float32, device=cm.device)
    if amax is not None:
        _lib.call("hpl_cm_to_rows_amax", cm.data_ptr(), cm.stride(0), n, c, rows.data_ptr(), rows.stride(0),
                  amax.data_ptr(), _stream())
    else:
        _lib.call("hpl_cm_to_rows", cm.data_ptr(), cm.stride(0), n, c, rows.data_ptr(), rows.stride(0), _stream())
    return rows


def rows_to_cm(rows, channels):
    _f32(rows, "rows")
    n = rows.size(0)
    cm = torch.empty((channels, n), dtype=torch.float32, device=rows.device)
    _lib.call("hpl_rows_to_cm", rows.data_ptr(), rows.stride(0), n, channels, cm.data_ptr(), cm.stride(0),
              _stream())
    return cm


def channel_sums(x, out=None, side=False):
    """Row sums of x (C, N); out: a zeroed (C,) tensor to accumulate into.  side: launch on the forked side stream
    (ops.fork_side; the caller joins with ops.join_side before the result is used)."""
    _f32(x, "x")
    c, n = x.shape
    s = out if out is not None else small_zeros(c, torch.float32, x.device)
    _lib.call("hpl_channel_sums", x.data_ptr(), c, n, s.data_ptr(), fork_side(x.device) if side else _stream())
    return s


# ------------------------------------------------------------------------------------------ engine 5 (tile plans)
def h16b_split(x, channels, amax, norm=None):
    """h16b image (per row and 32-channel block: 32 fp16 hi | 32 fp16 lo) of a vertex-major matrix; norm (rows,):
    the rows are first multiplied by 1/(norm + 1e-5) (density normalisation fused into the split)."""
    _f32(x, "x")
    n = x.size(0)
    buf = torch.empty(max(_lib.load().hpl_h16b_bytes(n, channels), 16), dtype=torch.uint8, device=x.device)
    _lib.call("hpl_h16b_split", x.data_ptr(), x.stride(0), n, channels, norm.data_ptr() if norm is not None else None,
              amax.data_ptr(), buf.data_ptr(), _stream())
    return buf


def conv5_supported(filter_size, c_in, c_out):
    return bool(_lib.load().hpl_conv5_supported(filter_size, c_in, c_out))


def _mirror_map(filter_size, device):
    from . import plans
    t = plans._mirror_tensor(filter_size, device)
    return t if t is not None else False


# Weight images built ahead of their contraction on a second side stream (ops.conv5_prepare): in a training step the
# weights change every step, so max|w| + the tile image are two small latency-bound kernels in front of every hpl_conv5;
# forked at the start of the module's forward / backward they run under the splat instead.
_weight_side = {}
_prepared = {}


def _wkey(w, mirror):
    return (w.data_ptr(), tuple(w.shape), tuple(w.stride()), bool(mirror))


def conv5_prepare(w, mirror=False):
    """Enqueue the weight image of a coming ops.conv5(w, mirror=...) call on the weight side stream (forked from the
    current stream here); that call joins the stream and uses the image."""
    f, c, co = w.shape
    # (only while a CUDA graph is being captured: there the fork / join are graph edges and cost nothing at replay; launched
    # eagerly, their host-side cost -- three stream operations -- exceeds the two small kernels they hide)
    if not torch.cuda.is_current_stream_capturing() or not _dense_permutation(w):
        return
    nbytes = _lib.load().hpl_conv5_workspace(c) + 256
    ws, ws_valid, commit = _cached_workspace(w, nbytes, ("e5", mirror))
    if ws_valid:
        return
    if commit is _no_commit:                                          # (the shared per-stream scratch may be rewritten before the call)
        ws = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=w.device)
    tap_map = _mirror_map(f, w.device) if mirror else None
    if tap_map is False:
        return
    cur, key = _cur_stream_obj(w.device)
    side = _weight_side.get(key)
    if side is None:
        side = _weight_side[key] = torch.cuda.Stream(w.device)
    side.wait_stream(cur)
    _lib.call("hpl_conv5_weights", w.data_ptr(), w.stride(0), w.stride(1), w.stride(2), f, c, co,
              tap_map.data_ptr() if tap_map is not None else None, ws.data_ptr(), side.cuda_stream)
    _prepared[_wkey(w, mirror)] = (ws, commit, side, cur)


def conv5(x16, plan, c_in, w, bias, act, x_amax, out=None, out_amax=None, mirror=False, tag="fwd"):
    """Engine 5 (csrc/gemm_plan.cu): out[v] = act(bias + sum_f x[nbr[f, v]] @ w[f_or_mirror(f)]) over plan's table.
    x16: h16b image of x; w (F, C, Co) any dense permutation; mirror=True uses w[mirror(f)] (data gradient)."""
    f, c, co = w.shape
    assert c == c_in and plan.usable and f == plan.filter_size
    if out is None:
        out = alloc_rows(plan.n_rows, co, x16.device)
    tap_map = None
    if mirror:
        tap_map = _mirror_map(f, x16.device)
        assert tap_map is not False
    if not _dense_permutation(w):
        w = w.contiguous()
    prep = _prepared.pop(_wkey(w, mirror), None) if _prepared else None
    if prep is not None:                                              # image built on the weight side stream: join it
        ws, commit, side, cur = prep
        cur.wait_stream(side)
        ws_valid = 1
    else:
        ws, ws_valid, commit = _cached_workspace(w, _lib.load().hpl_conv5_workspace(c) + 256, ("e5", mirror))
    with _timed(tag):
        _lib.call("hpl_conv5", x16.data_ptr(), plan.buf.data_ptr(), plan.n_rows, f, c, co,
                  w.data_ptr(), w.stride(0), w.stride(1), w.stride(2),
                  tap_map.data_ptr() if tap_map is not None else None,
                  bias.data_ptr() if bias is not None else None, act, out.data_ptr(), out.stride(0),
                  ws.data_ptr(), ws_valid, x_amax.data_ptr(), out_amax.data_ptr() if out_amax is not None else None,
                  _stream())
    commit()
    return out


def wgrad5(x16, dz16, plan, c_in, c_out, x_amax, dz_amax, out=None):
    """Engine 5 weight gradient: dw (F, C, Co) = sum_v x[nbr[f, v]]^T dz[v] over plan's table (h16b images in).
    out: a zeroed (F * C * Co,) buffer to accumulate into."""
    assert plan.usable
    dw = (out.view(plan.filter_size, c_in, c_out) if out is not None
          else torch.zeros((plan.filter_size, c_in, c_out), dtype=torch.float32, device=x16.device))
    with _timed("wgrad"):
        _lib.call("hpl_wgrad5", x16.data_ptr(), dz16.data_ptr(), plan.buf.data_ptr(), plan.n_rows, plan.filter_size,
                  c_in, c_out, dw.data_ptr(), x_amax.data_ptr(), dz_amax.data_ptr(), _stream())
    return dw


def h16b_splat_csr(src, channels, bary, splan, amax_a, normalize=False, inv_out=None, norm_amax_out=None, y=None, act=ACT_NONE,
                   amax_b=None, amax_out=None, colsum=None):
    """Splat as a deterministic gather fused with the operand split (include/hplflownet_b200.h: hpl_h16b_splat_csr).
    src (N, ld) point-major rows (ops.cm_to_rows), bary (4, N), splan: plans.SplatPlan.  Returns the h16b image."""
    _f32(src, "src"); _f32(bary, "bary")
    h = splan.n_rows
    buf = torch.empty(max(_lib.load().hpl_h16b_bytes(h, channels), 16), dtype=torch.uint8, device=src.device)
    ptr = lambda t: t.data_ptr() if t is not None else None      # noqa: E731
    has_y = y is not None and act != ACT_NONE
    _lib.call("hpl_h16b_splat_csr", src.data_ptr(), src.stride(0), bary.data_ptr(), splan.n_points, splan.ptr.data_ptr(),
              splan.ent.data_ptr(), h, channels, int(bool(normalize)), ptr(inv_out), ptr(norm_amax_out),
              y.data_ptr() if has_y else None, y.stride(0) if has_y else 0, act if has_y else ACT_NONE,
              amax_a.data_ptr(), ptr(amax_b), ptr(amax_out), ptr(colsum), buf.data_ptr(), _stream())
    return buf


def h16b_split_ex(x, channels, amax_a, norm=None, inv_out=None, norm_amax_out=None, y=None, act=ACT_NONE, amax_b=None,
                  amax_out=None, colsum=None, dispose=0):
    """h16b image of x with the surrounding passes folded in (include/hplflownet_b200.h: hpl_h16b_split_ex)."""
    _f32(x, "x")
    n = x.size(0)
    buf = torch.empty(max(_lib.load().hpl_h16b_bytes(n, channels), 16), dtype=torch.uint8, device=x.device)
    ptr = lambda t: t.data_ptr() if t is not None else None      # noqa: E731
    has_y = y is not None and act != ACT_NONE
    _lib.call("hpl_h16b_split_ex", x.data_ptr(), x.stride(0), n, channels, ptr(norm), ptr(inv_out), ptr(norm_amax_out),
              y.data_ptr() if has_y else None, y.stride(0) if has_y else 0, act if has_y else ACT_NONE,
              amax_a.data_ptr(), ptr(amax_b), ptr(amax_out), ptr(colsum), dispose, buf.data_ptr(), _stream())
    return buf
