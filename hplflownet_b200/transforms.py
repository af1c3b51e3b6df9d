"""GPU lattice builder -- drop-in for the reference's ``GenerateDataUnsymmetric``
(transforms/transforms.py:264-491).

Same constructor (reads ``args.dim`` and ``args.scales_filter_map``), same call signature
(``gen([pc1, pc2, sf]) -> (pc1, pc2, sf, generated_data)``) and the same 12-entry dict per scale
(:471-483), but every tensor is built on the GPU by the kernels in csrc/lattice.cu and stays
there -- the reference builds them with torch-CPU + Numba + a cffi khash table inside DataLoader
workers (4-5 s per 8192-point pair) and ships them host-to-device every step.

All outputs are bit-exact with the reference (tests/test_gpu_lattice.py): barycentric weights,
el_minus_gr, lattice offsets (first-occurrence vertex numbering), vertex counts, blur and
correlation tables.  One host synchronisation per scale reads back the two vertex counts
(``pc*_hash_cnt`` are Python ints in the reference as well, :390-391).

Known deviation (oracle/make_golden.py): first-level clouds of 2..11 points.
"""
import math

import numpy as np
import torch

from . import _lib

D = 3
D1 = 4
EXPECTED_STD = (D + 1) * math.sqrt(2.0 / 3.0)        # transforms.py:275


def filter_size(radius, d1=D1):
    """(r+1)^(d+1) - r^(d+1)  (transforms.py:355-356)."""
    return (radius + 1) ** d1 - radius ** d1


def neighbor_offsets(radius, d1=D1):
    """Neighbourhood offsets in conv-weight order (Traverse.go, transforms.py:112-130): every
    step vector (i_0..i_d) in [0, r]^(d+1) with at least one zero, lexicographic with i_0
    slowest; a step of i in dimension j moves the key by -i everywhere and +i*(d+1) at j."""
    out = []
    for steps in np.ndindex(*([radius + 1] * d1)):
        if min(steps) != 0:
            continue
        total = sum(steps)
        out.append([d1 * s - total for s in steps])
    return np.asarray(out, dtype=np.int32)


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object through
    # several Python layers: ~12 us per call, 270 calls per model forward)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _Cloud:
    """Device workspace of one cloud at one scale.  The tensors the caller receives (barycentric weights, el_minus_gr) are
    separate allocations; everything that lives only during the build of this scale -- the hash table, the per-point
    scratch, the vertex coordinates -- is carved out of ONE allocation and addressed by raw pointers (a scale used to make
    ~14 allocator calls per cloud, ~0.3 ms of a 1.4 ms build)."""

    def __init__(self, pc, device):
        self.pc = pc                                    # (3, N) fp32 contiguous
        self.n = pc.size(1)
        n = self.n
        L = _lib.load()
        self.bary = torch.empty((D1, n), dtype=torch.float32, device=device)
        self.emg = torch.empty((D1, n), dtype=torch.float32, device=device)
        self.cap = L.hpl_lattice_table_capacity(n)
        sizes = (("greedy", n * 16), ("rankpack", n * 4), ("table_keys", self.cap * 8), ("table_first", self.cap * 4),
                 ("table_ids", self.cap * 4), ("slot_of", 4 * n * 4), ("scan_ws", L.hpl_lattice_scan_blocks(n) * 4),
                 ("vertex_coords", 4 * n * 16))
        total = 0
        offs = {}
        for name, nbytes in sizes:
            offs[name] = total
            total += (nbytes + 255) // 256 * 256
        self.scratch = torch.empty(total, dtype=torch.uint8, device=device)       # (cudaMalloc alignment: 256 B)
        base = self.scratch.data_ptr()
        self.greedy, self.rankpack = base + offs["greedy"], base + offs["rankpack"]
        self.table_keys, self.table_first, self.table_ids = base + offs["table_keys"], base + offs["table_first"], base + offs["table_ids"]
        self.slot_of, self.scan_ws, self.vertex_coords = base + offs["slot_of"], base + offs["scan_ws"], base + offs["vertex_coords"]
        self.offset = None
        self.h = None


class GenerateDataUnsymmetric(object):
    def __init__(self, args, device="cuda", index_dtype=torch.int64):
        """args.dim must be 3; args.scales_filter_map rows are [scale, bcn_r, corr_filter_r, corr_corr_r]
        with -1 = "table not needed" (transforms.py:265-298).  ``index_dtype``: int64 reproduces the
        reference's tensors; int32 halves the table bytes (the CUDA modules accept both)."""
        if args.dim != D:
            raise ValueError("only dim == 3 is built (all reference configs use 3)")
        assert index_dtype in (torch.int64, torch.int32)
        self.d, self.d1 = D, D1
        self.scales_filter_map = args.scales_filter_map
        self.device = torch.device(device)
        self.index_dtype = index_dtype
        self.expected_std = EXPECTED_STD
        self.radius2offset = {}
        for line in self.scales_filter_map:
            for r in line[1:]:
                if r != -1 and r not in self.radius2offset:
                    self.radius2offset[r] = neighbor_offsets(int(r))
        self._dev_offsets = {}

    def get_filter_size(self, radius):
        return filter_size(radius, self.d1)

    def _offsets_on_device(self, radius):
        if radius not in self._dev_offsets:
            self._dev_offsets[radius] = torch.from_numpy(self.radius2offset[radius]).to(self.device).contiguous()
        return self._dev_offsets[radius]

    # ---- one scale -------------------------------------------------------------------------
    def _points(self, cloud, scale, key_minmax):
        _lib.call("hpl_lattice_points", cloud.pc.data_ptr(), cloud.n, float(scale), cloud.bary.data_ptr(),
                  cloud.emg.data_ptr(), cloud.greedy, cloud.rankpack,
                  key_minmax, _stream())

    def _insert(self, cloud, key_minmax, counts, slot):
        cloud.offset = torch.empty((D1, cloud.n), dtype=self.index_dtype, device=self.device)
        _lib.call("hpl_lattice_insert", cloud.greedy, cloud.rankpack, cloud.n,
                  key_minmax, cloud.table_keys, cloud.table_first,
                  cloud.table_ids, cloud.cap, cloud.slot_of, cloud.scan_ws,
                  cloud.offset.data_ptr(), int(self.index_dtype == torch.int64), cloud.vertex_coords,
                  counts + 4 * slot, _stream())

    def _neighbors(self, src, table, key_minmax, counts, slot, radius):
        offs = self._offsets_on_device(radius)
        f = offs.size(0)
        out = torch.empty((f, src.h), dtype=self.index_dtype, device=self.device)
        _lib.call("hpl_lattice_neighbors", src.vertex_coords, counts + 4 * slot, src.h,
                  key_minmax, table.table_keys, table.table_ids, table.cap,
                  offs.data_ptr(), f, out.data_ptr(), int(self.index_dtype == torch.int64), src.h, _stream())
        return out

    def _corr(self, c1, c2, key_minmax, counts, corr_radius, filt_radius):
        co, fo = self._offsets_on_device(corr_radius), self._offsets_on_device(filt_radius)
        p, f = co.size(0), fo.size(0)
        out = torch.empty((f, p, c1.h), dtype=self.index_dtype, device=self.device)
        _lib.call("hpl_lattice_corr_table", c1.vertex_coords, counts, c1.h,
                  key_minmax, c2.table_keys, c2.table_ids, c2.cap,
                  co.data_ptr(), p, fo.data_ptr(), f, out.data_ptr(), int(self.index_dtype == torch.int64),
                  c1.h, _stream())
        return out

    def _next_points(self, cloud, scale):
        out = torch.empty((D, cloud.h), dtype=torch.float32, device=self.device)
        _lib.call("hpl_lattice_next_points", cloud.vertex_coords, cloud.h,
                  float(np.float32(self.expected_std * scale)), out.data_ptr(), _stream())
        return out

    # ---- public ----------------------------------------------------------------------------
    def build(self, pc1, pc2):
        """pc1, pc2: (3, N) fp32 CUDA tensors (already transposed).  Returns generated_data."""
        dev = self.device
        last1 = pc1.to(dev, torch.float32).contiguous()
        last2 = pc2.to(dev, torch.float32).contiguous()
        generated = []
        n_scales = len(self.scales_filter_map)
        placeholder = lambda: torch.zeros(1, dtype=torch.long, device=dev)      # :450-459
        for idx, (scale, bcn_r, corr_f_r, corr_c_r) in enumerate(self.scales_filter_map):
            small = torch.empty(12, dtype=torch.int32, device=dev)               # key range (8 ints) | vertex counts (2 ints)
            key_minmax = small.data_ptr()                                        # (raw pointers: no view tensors per call)
            counts = key_minmax + 32
            _lib.call("hpl_lattice_init_range", key_minmax, _stream())
            c1, c2 = _Cloud(last1, dev), _Cloud(last2, dev)
            self._points(c1, scale, key_minmax)
            self._points(c2, scale, key_minmax)
            self._insert(c1, key_minmax, counts, 0)
            self._insert(c2, key_minmax, counts, 1)
            c1.h, c2.h = [int(x) for x in small[8:10].tolist()]                  # the one sync per scale

            if bcn_r != -1:
                blur1 = self._neighbors(c1, c1, key_minmax, counts, 0, bcn_r)
                blur2 = self._neighbors(c2, c2, key_minmax, counts, 1, bcn_r)
            else:
                blur1, blur2 = placeholder(), placeholder()
            if corr_f_r != -1:
                corr1 = self._neighbors(c1, c1, key_minmax, counts, 0, corr_c_r)
                corr2 = self._corr(c1, c2, key_minmax, counts, corr_c_r, corr_f_r)
            else:
                corr1, corr2 = placeholder(), placeholder()

            generated.append({
                "pc1_barycentric": c1.bary, "pc2_barycentric": c2.bary,
                "pc1_el_minus_gr": c1.emg, "pc2_el_minus_gr": c2.emg,
                "pc1_lattice_offset": c1.offset, "pc2_lattice_offset": c2.offset,
                "pc1_blur_neighbors": blur1, "pc2_blur_neighbors": blur2,
                "pc1_corr_indices": corr1, "pc2_corr_indices": corr2,
                "pc1_hash_cnt": c1.h, "pc2_hash_cnt": c2.h,
            })
            if idx != n_scales - 1:
                last1, last2 = self._next_points(c1, scale), self._next_points(c2, scale)
        return generated

    def __call__(self, data):
        """data = [pc1, pc2, sf] with (N, 3) numpy arrays or tensors, as the reference takes them
        (transforms.py:358-366).  Returns (pc1, pc2, sf, generated_data), all (3, N) / on the GPU."""
        pc1, pc2, sf = data
        if pc1 is None:
            return None, None, None, None
        with torch.no_grad():
            to_t = lambda a: (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a)
            pc1 = to_t(pc1).to(self.device, torch.float32).t().contiguous()
            pc2 = to_t(pc2).to(self.device, torch.float32).t().contiguous()
            sf = to_t(sf).to(self.device, torch.float32).t().contiguous()
            return pc1, pc2, sf, self.build(pc1, pc2)

    def __repr__(self):
        return "%s\n(scales_filter_map: %s\n)" % (self.__class__.__name__, self.scales_filter_map)


def collate_batch1(generated_data):
    """What torch's default_collate does to one sample: add the B=1 axis to every
    tensor and turn the Python-int vertex counts into ``tensor([H])`` (HPLFlowNet.py:283 calls
    ``.item()`` on them)."""
    out = []
    for d in generated_data:
        out.append({k: (torch.tensor([v]) if isinstance(v, int) else v.unsqueeze(0)) for k, v in d.items()})
    return out
