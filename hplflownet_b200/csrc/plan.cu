// Tile plans for the lattice convolution (engine 5, gemm_plan.cu).
//
// The blur step gathers F = 15 neighbour rows per lattice vertex (models/bilateralNN.py:198-217: the reference
// materialises the F-times gathered copy; engines 2 / 4 re-read every row ~15x from L2).  Lattice neighbourhoods
// overlap heavily once vertices are visited in a spatially coherent order: a tile of 128 such vertices touches
// ~320 distinct rows instead of 1920 (tools/plan_probe.py).  A *plan* is the per-lattice, per-table precomputation
// that lets the contraction kernels exploit this:
//
//   order      (H)                   permutation of the table's columns along a Morton curve of the vertices' lattice
//                                    coordinates (vertices keep their reference ids: only the tile membership changes)
//   tile_rows  (n_tiles, 128) int32  output row (= table column) of every tile slot, -1 = padding
//   n_uniq     (n_tiles)      int32  number of distinct input rows the tile's F x 128 table entries reference
//   uniq       (n_tiles, 464) int32  those rows (staged ONCE per tile into shared memory by the kernels)
//   local      (n_tiles, 16, 128) u16  per (tap, tile slot): index into `uniq`, 464 = "zero row" (missing neighbour)
//
// A table whose tiles reference more than 464 distinct rows (vertex order without locality) is reported through
// `stats`; callers then use the direct-gather engine.
//
// Coordinates are not part of the reference's data contract (transforms.py:471-483 hands out tables only), so they are
// RECONSTRUCTED from the table itself: coord(nbr[f, v]) = coord(v) + offset_f, propagated as a min-label relaxation
// (every connected component ends up in the frame of its smallest vertex id).  Sorting: LSD radix sort, 8 bits a pass,
// of (component, Morton code) keys.
#include "common.cuh"

namespace {

constexpr int TM = 128;
constexpr int kTaps = 16;
constexpr int kUmax = 464;
constexpr int kHash = 4096;                     // >= 2 * 15 * 128 entries: load <= 0.47

// ---------------------------------------------------------------------------------------------- tile build
// One CTA per tile.
template <bool I64>
__global__ void __launch_bounds__(256)
plan_tiles_kernel(const void* __restrict__ nbr, int filter_size, long long n_rows, long long n_in_rows,
                  const int* __restrict__ order, int* __restrict__ tile_rows, int* __restrict__ n_uniq,
                  int* __restrict__ uniq, unsigned short* __restrict__ local, int* __restrict__ stats) {
    __shared__ int keys[kHash];
    __shared__ unsigned short slot_of[kHash];
    __shared__ unsigned short pos[kTaps * TM];
    __shared__ int rows[TM];
    __shared__ int warp_tot[8];
    __shared__ int total;

    const long long tile = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < kHash; i += 256) keys[i] = -1;
    if (tid < TM) {
        const long long s = tile * TM + tid;
        int r = -1;
        if (s < n_rows) r = order != nullptr ? order[s] : (int)s;
        rows[tid] = r;
        tile_rows[tile * TM + tid] = r;
    }
    __syncthreads();

    // insert every referenced row into the shared-memory hash set
    for (int e = tid; e < kTaps * TM; e += 256) {
        const int f = e >> 7, i = e & (TM - 1);
        int u = -1;
        if (f < filter_size && rows[i] >= 0) u = load_idx<I64>(nbr, (long long)f * n_rows + rows[i]);
        unsigned short p = 0xffff;
        if (u >= 0 && u < n_in_rows) {
            unsigned h = ((unsigned)u * 2654435761u) >> 20;               // 12 bits
            for (;;) {
                const int old = atomicCAS(&keys[h], -1, u);
                if (old == -1 || old == u) break;
                h = (h + 1) & (kHash - 1);
            }
            p = (unsigned short)h;
        }
        pos[e] = p;
    }
    __syncthreads();

    // slot = rank of the occupied hash position (block-wide exclusive scan over 4096 flags, 16 per thread)
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt += keys[tid * 16 + k] >= 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int base = incl - cnt;
    for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
    if (tid == 255) total = base + cnt;
    __syncthreads();
    const int n = total;
    const bool fits = n <= kUmax;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int h = tid * 16 + k;
        if (keys[h] >= 0) {
            slot_of[h] = (unsigned short)base;
            if (fits) uniq[tile * kUmax + base] = keys[h];
            ++base;
        }
    }
    __syncthreads();
    for (int e = tid; e < kTaps * TM; e += 256) {
        const unsigned short p = pos[e];
        local[tile * (kTaps * TM) + e] = (p == 0xffff || !fits) ? (unsigned short)kUmax : slot_of[p];
    }
    if (tid == 0) {
        n_uniq[tile] = fits ? n : -n;                   // negative: the tile does not fit (callers fall back)
        atomicMax(&stats[0], n);
        if (!fits) atomicAdd(&stats[1], 1);
        atomicAdd(&stats[2], n);
    }
}

// ------------------------------------------------------------------------------------------- symmetry check
// The data gradient of the convolution gathers through the TRANSPOSED table.  For a lattice table the transpose is a
// row permutation of the table itself: u = nbr[f, v]  <=>  v = nbr[mirror(f), u]  (the offset set is closed under
// negation, transforms.py:112-130).  Counts the entries that violate this (0 for every table the lattice builder makes;
// arbitrary user tables fall back to an explicit transpose).
template <bool I64>
__global__ void plan_symmetry_kernel(const void* __restrict__ nbr, int filter_size, long long n_rows,
                                     const int* __restrict__ mirror, int* __restrict__ violations) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)filter_size * n_rows) return;
    const int f = (int)(t / n_rows);
    const long long v = t - (long long)f * n_rows;
    const int u = load_idx<I64>(nbr, t);
    if (u < 0) return;
    bool bad = u >= n_rows;
    if (!bad) bad = load_idx<I64>(nbr, (long long)mirror[f] * n_rows + u) != (int)v;
    if (bad) atomicAdd(violations, 1);
}

// --------------------------------------------------------------------------------- coordinate reconstruction
// state per vertex: root (component label = smallest vertex id reached so far) and integer coordinates relative to
// that root.  Jacobi relaxation (read buffer `a`, write buffer `b`): a vertex adopts the smallest root among itself
// and its table neighbours together with the coordinates that root's frame implies.  `changed` counts adoptions.
struct VState { int root, x, y, z; };

template <bool I64>
__global__ void plan_relax_kernel(const void* __restrict__ nbr, int filter_size, long long n_rows,
                                  const int4* __restrict__ offs, const VState* __restrict__ a, VState* __restrict__ b,
                                  int* __restrict__ changed) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    VState best = a[v];
    bool ch = false;
    for (int f = 0; f < filter_size; ++f) {
        const int u = load_idx<I64>(nbr, (long long)f * n_rows + v);
        if (u < 0 || u >= n_rows) continue;
        const VState s = a[u];
        if (s.root < best.root) {                        // u = v + offset_f  ->  coord(v) = coord(u) - offset_f
            const int4 o = offs[f];
            best.root = s.root; best.x = s.x - o.x; best.y = s.y - o.y; best.z = s.z - o.z;
            ch = true;
        }
    }
    b[v] = best;
    if (ch) atomicAdd(changed, 1);
}

__global__ void plan_init_state_kernel(VState* __restrict__ a, long long n_rows) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_rows) a[v] = VState{(int)v, 0, 0, 0};
}

// coordinate range over all vertices (per axis), for the Morton code's origin
__global__ void plan_coord_min_kernel(const VState* __restrict__ a, long long n_rows, int* __restrict__ mins) {
    int mx = INT_MAX, my = INT_MAX, mz = INT_MAX;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_rows; v += (long long)gridDim.x * blockDim.x) {
        const VState s = a[v];
        mx = min(mx, s.x); my = min(my, s.y); mz = min(mz, s.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = min(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        my = min(my, __shfl_xor_sync(0xffffffffu, my, o));
        mz = min(mz, __shfl_xor_sync(0xffffffffu, mz, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mins[0], mx); atomicMin(&mins[1], my); atomicMin(&mins[2], mz); }
}

__device__ __forceinline__ unsigned long long spread3(unsigned v) {     // 21 bits -> every third bit
    unsigned long long x = v & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// key = component (root id, high bits) | Morton code of the root-relative coordinates (low 3 * bits_per_axis bits)
// (the first three of the four hyperplane coordinates; tools/plan_probe.py: a Euclidean embedding is no better).
__global__ void plan_keys_kernel(const VState* __restrict__ a, long long n_rows, const int* __restrict__ mins,
                                 int bits_per_axis, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const VState s = a[v];
    const unsigned mask = (1u << bits_per_axis) - 1u;
    const unsigned x = (unsigned)(s.x - mins[0]) & mask, y = (unsigned)(s.y - mins[1]) & mask, z = (unsigned)(s.z - mins[2]) & mask;
    const unsigned long long m = spread3(x) | spread3(y) << 1 | spread3(z) << 2;
    keys[v] = ((unsigned long long)(unsigned)s.root << (3 * bits_per_axis)) | m;
    vals[v] = (int)v;
}

// --------------------------------------------------------------------------------------------- radix sort
constexpr int kSortThreads = 256, kSortItems = 8, kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int n_blocks, int* __restrict__ counts) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const long long i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255], 1);
    }
    __syncthreads();
    counts[threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];        // digit-major
}

// exclusive scan of `count` ints in place, one CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_exclusive_kernel(int* __restrict__ data, int count) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < count; base += 1024) {
        const int i = base + threadIdx.x;
        const int x = i < count ? data[i] : 0;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) off += warp_tot[w];
        if (i < count) data[i] = off + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, const int* __restrict__ vals_in, long long n, int shift,
                     int n_blocks, const int* __restrict__ offsets, unsigned long long* __restrict__ keys_out,
                     int* __restrict__ vals_out) {
    __shared__ int base[256];
    __shared__ int wcnt[8][256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    base[tid] = offsets[tid * n_blocks + blockIdx.x];
    for (int w = 0; w < 8; ++w) wcnt[w][tid] = 0;
    __syncthreads();
    const long long blk = (long long)blockIdx.x * kSortTile;
    for (int r = 0; r < kSortItems; ++r) {
        const long long i = blk + r * kSortThreads + tid;
        const bool ok = i < n;
        unsigned long long k = 0;
        int v = 0, d = 0;
        if (ok) { k = keys_in[i]; v = vals_in[i]; d = (int)((k >> shift) & 255); }
        // stable rank inside the warp among equal digits
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256 + lane) & act;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) wcnt[warp][d] = __popc(peers);
        __syncthreads();
        if (ok) {
            int off = base[d] + rank;
            for (int w = 0; w < warp; ++w) off += wcnt[w][d];
            keys_out[off] = k;
            vals_out[off] = v;
        }
        __syncthreads();
        int add = 0;
        for (int w = 0; w < 8; ++w) { add += wcnt[w][tid]; wcnt[w][tid] = 0; }
        base[tid] += add;
        __syncthreads();
    }
}

int ceil_log2(long long x) { int b = 0; while ((1LL << b) < x) ++b; return b; }

}  // namespace

extern "C" {

int64_t hpl_plan_tiles(int64_t n_rows) { return (n_rows + TM - 1) / TM; }
int64_t hpl_plan_umax(void) { return kUmax; }

/* byte offsets of the plan's arrays inside one allocation (each 256-byte aligned) */
static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }
int64_t hpl_plan_offset(int64_t n_rows, int which) {
    const int64_t t = hpl_plan_tiles(n_rows);
    const int64_t o_rows = 0, o_nuniq = align256(o_rows + t * TM * 4), o_uniq = align256(o_nuniq + t * 4);
    const int64_t o_local = align256(o_uniq + t * kUmax * 4), o_end = align256(o_local + t * kTaps * TM * 2);
    switch (which) {
        case 0: return o_rows;
        case 1: return o_nuniq;
        case 2: return o_uniq;
        case 3: return o_local;
        default: return o_end;
    }
}
int64_t hpl_plan_bytes(int64_t n_rows) { return hpl_plan_offset(n_rows, 4); }

int hpl_plan_build(const void* nbr, int idx64, int64_t filter_size, int64_t n_rows, int64_t n_in_rows, const int32_t* order,
                   const int32_t* tap_mirror, void* plan, int32_t* stats, void* stream) {
    HPL_CHECK_ARG(nbr && plan && stats && filter_size > 0 && filter_size <= kTaps && n_rows >= 0 && n_in_rows >= 0);
    HPL_CHECK_ARG(n_rows < (1LL << 31) - TM && n_in_rows < (1LL << 31) && ((uintptr_t)plan & 255) == 0);
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(stats, 0, 16, s);
    if (e != cudaSuccess) return (int)e;
    if (n_rows == 0) return 0;
    uint8_t* base = reinterpret_cast<uint8_t*>(plan);
    int* tile_rows = reinterpret_cast<int*>(base + hpl_plan_offset(n_rows, 0));
    int* n_uniq = reinterpret_cast<int*>(base + hpl_plan_offset(n_rows, 1));
    int* uniq = reinterpret_cast<int*>(base + hpl_plan_offset(n_rows, 2));
    unsigned short* local = reinterpret_cast<unsigned short*>(base + hpl_plan_offset(n_rows, 3));
    const unsigned grid = (unsigned)hpl_plan_tiles(n_rows);
    if (idx64)
        plan_tiles_kernel<true><<<grid, 256, 0, s>>>(nbr, (int)filter_size, n_rows, n_in_rows, order, tile_rows, n_uniq, uniq, local, stats);
    else
        plan_tiles_kernel<false><<<grid, 256, 0, s>>>(nbr, (int)filter_size, n_rows, n_in_rows, order, tile_rows, n_uniq, uniq, local, stats);
    if (tap_mirror != nullptr && n_in_rows == n_rows) {
        const unsigned g = (unsigned)((filter_size * n_rows + 255) / 256);
        if (idx64) plan_symmetry_kernel<true><<<g, 256, 0, s>>>(nbr, (int)filter_size, n_rows, tap_mirror, stats + 3);
        else plan_symmetry_kernel<false><<<g, 256, 0, s>>>(nbr, (int)filter_size, n_rows, tap_mirror, stats + 3);
    } else {
        cudaMemsetAsync(stats + 3, 0xff, 4, s);                           // -1: not checked
    }
    HPL_RETURN_LAST();
}

/* Vertex order along a Morton curve, from the table alone.
 * offsets (F, 4) int32: lattice offset of every tap (transforms.py:112-130,292-298; the 4th coordinate is redundant).
 * workspace: hpl_plan_order_workspace(n_rows) bytes.  order (n_rows) int32 out.  `iterations` relaxation sweeps are
 * run (>= the largest component's hop diameter for an exact reconstruction; fewer only costs locality, never
 * correctness).  changed_out (int32, may be NULL): adoptions in the LAST sweep (0 = converged). */
int64_t hpl_plan_order_workspace(int64_t n_rows) {
    const int64_t nb = (n_rows + kSortTile - 1) / kSortTile;
    return align256(n_rows * 16) * 2 + align256(n_rows * 8) * 2 + align256(n_rows * 4) * 2 + align256(256 * nb * 4) + 256 + 256;
}

int hpl_plan_order(const void* nbr, int idx64, int64_t filter_size, int64_t n_rows, const int32_t* offsets, int iterations,
                   void* workspace, int32_t* order, int32_t* changed_out, void* stream) {
    HPL_CHECK_ARG(nbr && offsets && workspace && order && filter_size > 0 && filter_size <= kTaps && iterations >= 0);
    HPL_CHECK_ARG(n_rows >= 0 && n_rows < (1LL << 31) - kSortTile && ((uintptr_t)workspace & 255) == 0);
    if (n_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
    VState* st_a = reinterpret_cast<VState*>(p); p += align256(n_rows * 16);
    VState* st_b = reinterpret_cast<VState*>(p); p += align256(n_rows * 16);
    unsigned long long* k_a = reinterpret_cast<unsigned long long*>(p); p += align256(n_rows * 8);
    unsigned long long* k_b = reinterpret_cast<unsigned long long*>(p); p += align256(n_rows * 8);
    int* v_a = reinterpret_cast<int*>(p); p += align256(n_rows * 4);
    int* v_b = reinterpret_cast<int*>(p); p += align256(n_rows * 4);
    const int nb = (int)((n_rows + kSortTile - 1) / kSortTile);
    int* counts = reinterpret_cast<int*>(p); p += align256(256LL * nb * 4);
    int* mins = reinterpret_cast<int*>(p); p += 256;
    int* changed = reinterpret_cast<int*>(p);

    const unsigned blocks = (unsigned)((n_rows + 255) / 256);
    plan_init_state_kernel<<<blocks, 256, 0, s>>>(st_a, n_rows);
    const int4* offs = reinterpret_cast<const int4*>(offsets);
    for (int it = 0; it < iterations; ++it) {
        if (it == iterations - 1) cudaMemsetAsync(changed, 0, 4, s);
        if (idx64) plan_relax_kernel<true><<<blocks, 256, 0, s>>>(nbr, (int)filter_size, n_rows, offs, st_a, st_b, changed);
        else plan_relax_kernel<false><<<blocks, 256, 0, s>>>(nbr, (int)filter_size, n_rows, offs, st_a, st_b, changed);
        VState* t = st_a; st_a = st_b; st_b = t;
    }
    if (changed_out != nullptr) {
        if (iterations == 0) cudaMemsetAsync(changed_out, 0, 4, s);
        else cudaMemcpyAsync(changed_out, changed, 4, cudaMemcpyDeviceToDevice, s);
    }
    cudaMemsetAsync(mins, 0x7f, 12, s);
    plan_coord_min_kernel<<<(unsigned)min((long long)blocks, 4LL * num_sms()), 256, 0, s>>>(st_a, n_rows, mins);
    const int root_bits = ceil_log2(n_rows > 1 ? n_rows : 2);
    int bits_per_axis = (64 - root_bits) / 3;
    if (bits_per_axis > 12) bits_per_axis = 12;
    plan_keys_kernel<<<blocks, 256, 0, s>>>(st_a, n_rows, mins, bits_per_axis, k_a, v_a);
    const int total_bits = root_bits + 3 * bits_per_axis;
    int* v_src = v_a;
    int* v_dst = v_b;
    for (int shift = 0; shift < total_bits; shift += 8) {
        const bool last = shift + 8 >= total_bits;
        radix_hist_kernel<<<nb, kSortThreads, 0, s>>>(k_a, n_rows, shift, nb, counts);
        scan_exclusive_kernel<<<1, 1024, 0, s>>>(counts, 256 * nb);
        radix_scatter_kernel<<<nb, kSortThreads, 0, s>>>(k_a, v_src, n_rows, shift, nb, counts, k_b, last ? order : v_dst);
        unsigned long long* tk = k_a; k_a = k_b; k_b = tk;
        int* tv = v_src; v_src = v_dst; v_dst = tv;
    }
    HPL_RETURN_LAST();
}

}  // extern "C"
