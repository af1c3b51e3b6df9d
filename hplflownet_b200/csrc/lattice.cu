// Permutohedral-lattice index path on the GPU (sm_100a) -- replaces the reference's CPU builder
// (transforms/transforms.py:264-485: torch-CPU elevation + Numba loops over a cffi khash table).
//
// Parity contract: barycentric weights, el_minus_gr, vertex ids, vertex
// counts, neighbour / correlation tables and next-scale points are BIT-EXACT with the reference.
// That pins the fp32 arithmetic (explicit *_rn intrinsics so nvcc never contracts or reorders:
// the elevation is a k-ordered FMA chain from a +0 accumulator, division is IEEE, rounding is
// half-to-even, the descending sort is stable) and the vertex numbering: ids are the order of
// FIRST OCCURRENCE in a point-outer / remainder-inner scan (transforms.py:179-192).  The
// reference gets that order for free from a sequential insert loop; here every (point,
// remainder) inserts its packed key in parallel with atomicCAS and races an atomicMin of its
// flat scan position, and a prefix sum over "I am the first occurrence" flags yields the ids.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int D1 = 4;
constexpr unsigned long long kEmpty = 0x8000000000000000ULL;   // never a packed key
constexpr int kScanThreads = 256;                             // one thread = one point = 4 scan slots

// transforms/transforms.py:271-276 -- elevate_mat, fp32 bit patterns (row-major 4x3).
__constant__ unsigned int c_elevate_bits[12] = {0x3f3504f3u, 0x3ed105ebu, 0x3e93cd3au, 0xbf3504f3u,
                                                0x3ed105ebu, 0x3e93cd3au, 0x00000000u, 0xbf5105ebu,
                                                0x3e93cd3au, 0x00000000u, 0x00000000u, 0xbf5db3d7u};
__device__ __forceinline__ float elevate(int i, int k) { return __uint_as_float(c_elevate_bits[i * 3 + k]); }

struct KeyRange {   // per-coordinate min and radix of the mixed-radix packing (transforms.py:70-86)
    long long mn[D1], radix[D1];
};
__device__ __forceinline__ KeyRange load_range(const int* __restrict__ minmax) {
    KeyRange r;
#pragma unroll
    for (int i = 0; i < D1; ++i) {
        r.mn[i] = minmax[i];
        r.radix[i] = (long long)minmax[D1 + i] - minmax[i] + 1;
    }
    return r;
}
// key2int: (((k0)*s1 + k1)*s2 + k2)*s3 + k3, no range check (transforms.py:79-86)
__device__ __forceinline__ long long pack_key(const int* key, const KeyRange& kr) {
    long long res = 0;
#pragma unroll
    for (int i = 0; i < D1 - 1; ++i) {
        res += key[i] - kr.mn[i];
        res *= kr.radix[i + 1];
    }
    return res + (key[D1 - 1] - kr.mn[D1 - 1]);
}
__device__ __forceinline__ unsigned int hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (unsigned int)x;
}

// ------------------------------------------------------------------ L2: per-point simplex
// transforms/transforms.py:300-353.  One thread per point.
__global__ void lattice_points_kernel(const float* __restrict__ pc, long long n, float scale,
                                      float* __restrict__ bary, float* __restrict__ emg,
                                      int4* __restrict__ greedy, unsigned int* __restrict__ rankpack,
                                      int* __restrict__ key_minmax) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lo[D1], hi[D1];
#pragma unroll
    for (int i = 0; i < D1; ++i) { lo[i] = INT_MAX; hi[i] = INT_MIN; }
    if (p < n) {
        // :377 last_pc[:3] *= scale
        const float x = __fmul_rn(pc[p], scale), y = __fmul_rn(pc[n + p], scale), z = __fmul_rn(pc[2 * n + p], scale);
        const float std32 = __uint_as_float(0x405105ecu);   // fp32(expected_std = 4*sqrt(2/3)), :275,:309
        float el[D1], gr[D1], em[D1];
        int rank[D1];
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            float acc = __fmaf_rn(elevate(i, 0), x, 0.0f);
            acc = __fmaf_rn(elevate(i, 1), y, acc);
            acc = __fmaf_rn(elevate(i, 2), z, acc);
            el[i] = __fmul_rn(acc, std32);
            gr[i] = __fmul_rn(rintf(__fdiv_rn(el[i], 4.0f)), 4.0f);   // :312
            em[i] = __fsub_rn(el[i], gr[i]);
        }
#pragma unroll
        for (int i = 0; i < D1; ++i) {   // :315-319 inverse permutation of a stable descending sort
            int r = 0;
#pragma unroll
            for (int j = 0; j < D1; ++j) r += (em[j] > em[i]) || (em[j] == em[i] && j < i);
            rank[i] = r;
        }
        const float rsum = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(gr[0], gr[1]), gr[2]), gr[3]), 4.0f);   // :322
        const float sign = rsum > 0.f ? -1.f : (rsum < 0.f ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < D1; ++i) {   // :324-334
            const float rf = (float)rank[i];
            const bool cond = ((rf >= __fsub_rn(4.0f, rsum)) && rsum > 0.f) || ((rf < -rsum) && rsum < 0.f);
            const float step = __fmul_rn(__fmul_rn(4.0f, sign), cond ? 1.0f : 0.0f);
            gr[i] = __fadd_rn(gr[i], step);
            rank[i] += (int)step + (int)rsum;
        }
        float b[D1 + 1] = {0.f, 0.f, 0.f, 0.f, 0.f};   // :337-345
#pragma unroll
        for (int i = 0; i < D1; ++i) em[i] = __fsub_rn(el[i], gr[i]);
        // dynamic indexing kept out of local memory: scatter with compile-time slots
#pragma unroll
        for (int i = 0; i < D1; ++i)
#pragma unroll
            for (int s = 0; s <= D1; ++s)
                if (s == 3 - rank[i]) b[s] = __fadd_rn(b[s], em[i]);
#pragma unroll
        for (int i = 0; i < D1; ++i)
#pragma unroll
            for (int s = 0; s <= D1; ++s)
                if (s == 4 - rank[i]) b[s] = __fsub_rn(b[s], em[i]);
#pragma unroll
        for (int s = 0; s <= D1; ++s) b[s] = __fdiv_rn(b[s], 4.0f);
        b[0] = __fadd_rn(b[0], __fadd_rn(1.0f, b[D1]));
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            bary[i * n + p] = b[i];
            emg[i * n + p] = em[i];
        }
        int g[D1];
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            g[i] = (int)gr[i];
            // keys over remainders r: g + r (r < 4-rank) or g + r - 4 (:281-285,:347)
            lo[i] = rank[i] > 0 ? g[i] - rank[i] : g[i];
            hi[i] = g[i] + 3 - rank[i];
        }
        greedy[p] = make_int4(g[0], g[1], g[2], g[3]);
        rankpack[p] = (unsigned)rank[0] | ((unsigned)rank[1] << 8) | ((unsigned)rank[2] << 16) | ((unsigned)rank[3] << 24);
    }
    // :384-385 key range, folded over every cloud that shares key_minmax
#pragma unroll
    for (int i = 0; i < D1; ++i) {
        int a = lo[i], c = hi[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
            c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
        }
        if ((threadIdx.x & 31) == 0 && a != INT_MAX) {
            atomicMin(key_minmax + i, a);
            atomicMax(key_minmax + D1 + i, c);
        }
    }
}

__device__ __forceinline__ void simplex_vertex(const int4 g, unsigned int rp, int r, int* key) {
    const int gi[D1] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < D1; ++i) {
        const int rank = (rp >> (8 * i)) & 0xff;
        key[i] = gi[i] + (rank < D1 - r ? r : r - D1);
    }
}

// ------------------------------------------------------------------ L5/L6: parallel insert
__global__ void hash_clear_kernel(unsigned long long* keys, int* first_pos, long long cap) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cap) { keys[t] = kEmpty; first_pos[t] = INT_MAX; }
}

__global__ void hash_insert_kernel(const int4* __restrict__ greedy, const unsigned int* __restrict__ rankpack,
                                   long long n, const int* __restrict__ key_minmax,
                                   unsigned long long* __restrict__ keys, int* __restrict__ first_pos,
                                   unsigned int mask, int* __restrict__ slot_of) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // flat scan position
    if (t >= n * D1) return;
    const long long p = t >> 2;
    const int r = (int)(t & 3);
    const KeyRange kr = load_range(key_minmax);
    int key[D1];
    simplex_vertex(greedy[p], rankpack[p], r, key);
    const unsigned long long packed = (unsigned long long)pack_key(key, kr);
    unsigned int slot = hash64(packed) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(keys + slot, kEmpty, packed);
        if (prev == kEmpty || prev == packed) break;
        slot = (slot + 1) & mask;
    }
    atomicMin(first_pos + slot, (int)t);
    slot_of[t] = (int)slot;
}

// flags of one point (4 consecutive scan positions): bit r set = first occurrence of its key
__device__ __forceinline__ unsigned int first_flags(long long p, long long n, const int* __restrict__ slot_of,
                                                   const int* __restrict__ first_pos) {
    unsigned int f = 0;
    if (p < n) {
        const int4 s = *reinterpret_cast<const int4*>(slot_of + 4 * p);
        const int base = (int)(4 * p);
        f |= (first_pos[s.x] == base + 0) ? 1u : 0u;
        f |= (first_pos[s.y] == base + 1) ? 2u : 0u;
        f |= (first_pos[s.z] == base + 2) ? 4u : 0u;
        f |= (first_pos[s.w] == base + 3) ? 8u : 0u;
    }
    return f;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const int s = warp_sums[w];
        if (w < warp) before += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return before + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
count_first_kernel(long long n, const int* __restrict__ slot_of, const int* __restrict__ first_pos,
                   int* __restrict__ block_counts) {
    const long long p = (long long)blockIdx.x * kScanThreads + threadIdx.x;
    int tot;
    block_exclusive_scan(__popc(first_flags(p, n, slot_of, first_pos)), &tot);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = tot;
}

// exclusive scan of the block counts in place (single CTA), total -> *n_vertices
__global__ void __launch_bounds__(kScanThreads)
scan_blocks_kernel(int* __restrict__ block_counts, int n_blocks, int* __restrict__ n_vertices) {
    int carry = 0;
    for (int base = 0; base < n_blocks; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_counts[i] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, &tot);
        if (i < n_blocks) block_counts[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *n_vertices = carry;
}

// ids of first occurrences; vertex coordinates in id order (the reference's last_pc, :188-189)
__global__ void __launch_bounds__(kScanThreads)
assign_ids_kernel(const int4* __restrict__ greedy, const unsigned int* __restrict__ rankpack, long long n,
                  const int* __restrict__ slot_of, const int* __restrict__ first_pos,
                  const int* __restrict__ block_offsets, int* __restrict__ slot_ids, int4* __restrict__ vertex_coords) {
    const long long p = (long long)blockIdx.x * kScanThreads + threadIdx.x;
    const unsigned int flags = first_flags(p, n, slot_of, first_pos);
    int tot;
    int id = block_offsets[blockIdx.x] + block_exclusive_scan(__popc(flags), &tot);
    if (flags == 0) return;
    const int4 g = greedy[p];
    const unsigned int rp = rankpack[p];
#pragma unroll
    for (int r = 0; r < D1; ++r)
        if (flags & (1u << r)) {
            int key[D1];
            simplex_vertex(g, rp, r, key);
            slot_ids[slot_of[4 * p + r]] = id;
            vertex_coords[id] = make_int4(key[0], key[1], key[2], key[3]);
            ++id;
        }
}

template <typename OutT>
__global__ void write_offsets_kernel(long long n, const int* __restrict__ slot_of, const int* __restrict__ slot_ids,
                                     OutT* __restrict__ lattice_offset) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // output position r*n + p
    if (t >= n * D1) return;
    const long long r = t / n, p = t - r * n;
    lattice_offset[t] = (OutT)slot_ids[slot_of[4 * p + r]];
}

// ------------------------------------------------------------------ lookups
__device__ __forceinline__ int lookup(const unsigned long long* __restrict__ keys, const int* __restrict__ slot_ids,
                                      unsigned int mask, long long packed) {
    const unsigned long long k = (unsigned long long)packed;
    unsigned int slot = hash64(k) & mask;
    while (true) {
        const unsigned long long cur = keys[slot];
        if (cur == k) return slot_ids[slot];
        if (cur == kEmpty) return -1;
        slot = (slot + 1) & mask;
    }
}

// blur_neighbors[f, h] = id of (key_h + offset_f) or -1   (transforms.py:209-221,:243-255)
template <typename OutT>
__global__ void neighbor_table_kernel(const int4* __restrict__ vertex_coords, const int* __restrict__ n_vertices_dev,
                                      long long h_cap, const int* __restrict__ key_minmax,
                                      const unsigned long long* __restrict__ keys, const int* __restrict__ slot_ids,
                                      unsigned int mask, const int* __restrict__ offsets, int filter_size,
                                      OutT* __restrict__ out, long long ld) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    const long long h_count = min((long long)*n_vertices_dev, h_cap);
    if (h >= h_count) return;
    const KeyRange kr = load_range(key_minmax);
    const int4 v = vertex_coords[h];
    int key[D1] = {v.x + offsets[4 * f], v.y + offsets[4 * f + 1], v.z + offsets[4 * f + 2], v.w + offsets[4 * f + 3]};
    out[(long long)f * ld + h] = (OutT)lookup(keys, slot_ids, mask, pack_key(key, kr));
}

// pc2_corr_indices[f, p, h] = id IN TABLE 2 of (key1_h + corr_offset_p + filter_offset_f)  (:223-241)
template <typename OutT>
__global__ void corr_table_kernel(const int4* __restrict__ vertex_coords1, const int* __restrict__ n_vertices_dev,
                                  long long h_cap, const int* __restrict__ key_minmax,
                                  const unsigned long long* __restrict__ keys2, const int* __restrict__ slot_ids2,
                                  unsigned int mask2, const int* __restrict__ corr_offsets, int corr_size,
                                  const int* __restrict__ filt_offsets, int filter_size, OutT* __restrict__ out,
                                  long long ld) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int fp = blockIdx.y;                 // f * corr_size + p
    const int f = fp / corr_size, p = fp - f * corr_size;
    const long long h_count = min((long long)*n_vertices_dev, h_cap);
    if (h >= h_count) return;
    const KeyRange kr = load_range(key_minmax);
    const int4 v = vertex_coords1[h];
    int key[D1] = {v.x + corr_offsets[4 * p] + filt_offsets[4 * f], v.y + corr_offsets[4 * p + 1] + filt_offsets[4 * f + 1],
                   v.z + corr_offsets[4 * p + 2] + filt_offsets[4 * f + 2],
                   v.w + corr_offsets[4 * p + 3] + filt_offsets[4 * f + 3]};
    out[(long long)fp * ld + h] = (OutT)lookup(keys2, slot_ids2, mask2, pack_key(key, kr));
}

// next-scale points: p = E^T . (key / fp32(expected_std*scale))   (transforms.py:461-467)
__global__ void next_points_kernel(const int4* __restrict__ vertex_coords, long long h, float divisor,
                                   float* __restrict__ out) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= h) return;
    const int4 c = vertex_coords[q];
    const float v[D1] = {__fdiv_rn((float)c.x, divisor), __fdiv_rn((float)c.y, divisor),
                         __fdiv_rn((float)c.z, divisor), __fdiv_rn((float)c.w, divisor)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < D1; ++i) acc = __fmaf_rn(elevate(i, k), v[i], acc);
        out[k * h + q] = acc;
    }
}

__global__ void init_minmax_kernel(int* key_minmax) {
    if (threadIdx.x < D1) key_minmax[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 2 * D1) key_minmax[threadIdx.x] = INT_MIN;
}

inline unsigned nblk(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

}  // namespace

extern "C" {

int hpl_lattice_init_range(int32_t* key_minmax, void* stream) {
    HPL_CHECK_ARG(key_minmax);
    init_minmax_kernel<<<1, 32, 0, as_stream(stream)>>>(key_minmax);
    HPL_RETURN_LAST();
}

int hpl_lattice_points(const float* pc, int64_t n_points, float scale, float* bary, float* el_minus_gr,
                       int32_t* greedy, uint32_t* rankpack, int32_t* key_minmax, void* stream) {
    HPL_CHECK_ARG(pc && bary && el_minus_gr && greedy && rankpack && key_minmax && n_points >= 0);
    HPL_CHECK_ARG(((uintptr_t)greedy & 15) == 0);
    if (n_points == 0) return 0;
    lattice_points_kernel<<<nblk(n_points, 256), 256, 0, as_stream(stream)>>>(
        pc, n_points, scale, bary, el_minus_gr, reinterpret_cast<int4*>(greedy), rankpack, key_minmax);
    HPL_RETURN_LAST();
}

int64_t hpl_lattice_table_capacity(int64_t n_points) {
    int64_t cap = 1024;
    while (cap < 2 * 4 * n_points) cap <<= 1;
    return cap;
}

int64_t hpl_lattice_scan_blocks(int64_t n_points) { return (n_points + kScanThreads - 1) / kScanThreads; }

int hpl_lattice_insert(const int32_t* greedy, const uint32_t* rankpack, int64_t n_points, const int32_t* key_minmax,
                       uint64_t* table_keys, int32_t* table_first, int32_t* table_ids, int64_t table_cap,
                       int32_t* slot_of, int32_t* scan_ws, void* lattice_offset, int idx64, int32_t* vertex_coords,
                       int32_t* n_vertices, void* stream) {
    HPL_CHECK_ARG(greedy && rankpack && key_minmax && table_keys && table_first && table_ids && slot_of && scan_ws);
    HPL_CHECK_ARG(lattice_offset && vertex_coords && n_vertices && n_points > 0);
    HPL_CHECK_ARG(table_cap >= 2 * 4 * n_points && (table_cap & (table_cap - 1)) == 0 && table_cap <= (1LL << 31));
    HPL_CHECK_ARG(n_points * 4 < INT_MAX && ((uintptr_t)slot_of & 15) == 0 && ((uintptr_t)vertex_coords & 15) == 0);
    cudaStream_t s = as_stream(stream);
    const unsigned int mask = (unsigned int)(table_cap - 1);
    const int4* g4 = reinterpret_cast<const int4*>(greedy);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(table_keys);
    hash_clear_kernel<<<nblk(table_cap, 256), 256, 0, s>>>(keys, table_first, table_cap);
    hash_insert_kernel<<<nblk(n_points * 4, 256), 256, 0, s>>>(g4, rankpack, n_points, key_minmax, keys, table_first, mask, slot_of);
    const int blocks = (int)hpl_lattice_scan_blocks(n_points);
    count_first_kernel<<<blocks, kScanThreads, 0, s>>>(n_points, slot_of, table_first, scan_ws);
    scan_blocks_kernel<<<1, kScanThreads, 0, s>>>(scan_ws, blocks, n_vertices);
    assign_ids_kernel<<<blocks, kScanThreads, 0, s>>>(g4, rankpack, n_points, slot_of, table_first, scan_ws, table_ids,
                                                     reinterpret_cast<int4*>(vertex_coords));
    if (idx64)
        write_offsets_kernel<long long><<<nblk(n_points * 4, 256), 256, 0, s>>>(n_points, slot_of, table_ids, (long long*)lattice_offset);
    else
        write_offsets_kernel<int><<<nblk(n_points * 4, 256), 256, 0, s>>>(n_points, slot_of, table_ids, (int*)lattice_offset);
    HPL_RETURN_LAST();
}

int hpl_lattice_neighbors(const int32_t* vertex_coords, const int32_t* n_vertices, int64_t h_cap,
                          const int32_t* key_minmax, const uint64_t* table_keys, const int32_t* table_ids,
                          int64_t table_cap, const int32_t* offsets, int64_t filter_size, void* out, int idx64,
                          int64_t ld, void* stream) {
    HPL_CHECK_ARG(vertex_coords && n_vertices && key_minmax && table_keys && table_ids && offsets && out);
    HPL_CHECK_ARG(filter_size > 0 && filter_size < 65536 && ld >= h_cap);
    if (h_cap == 0) return 0;
    dim3 grid(nblk(h_cap, 256), (unsigned)filter_size);
    const unsigned int mask = (unsigned int)(table_cap - 1);
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(table_keys);
    const int4* vc = reinterpret_cast<const int4*>(vertex_coords);
    if (idx64)
        neighbor_table_kernel<long long><<<grid, 256, 0, as_stream(stream)>>>(vc, n_vertices, h_cap, key_minmax, keys, table_ids, mask, offsets, (int)filter_size, (long long*)out, ld);
    else
        neighbor_table_kernel<int><<<grid, 256, 0, as_stream(stream)>>>(vc, n_vertices, h_cap, key_minmax, keys, table_ids, mask, offsets, (int)filter_size, (int*)out, ld);
    HPL_RETURN_LAST();
}

int hpl_lattice_corr_table(const int32_t* vertex_coords1, const int32_t* n_vertices1, int64_t h_cap,
                           const int32_t* key_minmax, const uint64_t* table_keys2, const int32_t* table_ids2,
                           int64_t table_cap2, const int32_t* corr_offsets, int64_t corr_size,
                           const int32_t* filter_offsets, int64_t filter_size, void* out, int idx64, int64_t ld,
                           void* stream) {
    HPL_CHECK_ARG(vertex_coords1 && n_vertices1 && key_minmax && table_keys2 && table_ids2 && corr_offsets && filter_offsets && out);
    HPL_CHECK_ARG(corr_size > 0 && filter_size > 0 && corr_size * filter_size < 65536 && ld >= h_cap);
    if (h_cap == 0) return 0;
    dim3 grid(nblk(h_cap, 256), (unsigned)(corr_size * filter_size));
    const unsigned int mask = (unsigned int)(table_cap2 - 1);
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(table_keys2);
    const int4* vc = reinterpret_cast<const int4*>(vertex_coords1);
    if (idx64)
        corr_table_kernel<long long><<<grid, 256, 0, as_stream(stream)>>>(vc, n_vertices1, h_cap, key_minmax, keys, table_ids2, mask, corr_offsets, (int)corr_size, filter_offsets, (int)filter_size, (long long*)out, ld);
    else
        corr_table_kernel<int><<<grid, 256, 0, as_stream(stream)>>>(vc, n_vertices1, h_cap, key_minmax, keys, table_ids2, mask, corr_offsets, (int)corr_size, filter_offsets, (int)filter_size, (int*)out, ld);
    HPL_RETURN_LAST();
}

int hpl_lattice_next_points(const int32_t* vertex_coords, int64_t n_vertices, float divisor, float* out, void* stream) {
    HPL_CHECK_ARG(vertex_coords && out && n_vertices >= 0);
    if (n_vertices == 0) return 0;
    next_points_kernel<<<nblk(n_vertices, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(vertex_coords), n_vertices, divisor, out);
    HPL_RETURN_LAST();
}

}  // extern "C"
