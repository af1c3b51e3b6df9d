// Bandwidth-bound kernels of the BCL value path (sm_100a): splat / slice as scatter-add /
// gather over vertex-major lattice rows, density normalisation, layout changes at the module
// boundary, activation backward, table transposition.
//
// Layout recap (include/hplflownet_b200.h): point features are channel-major (C, N), lattice
// values are vertex-major (H, ld) with 16-byte aligned rows, so one lattice row is one
// contiguous, vectorisable burst and a scatter/gather touches 4 rows per point.
#include "common.cuh"

namespace {

constexpr int kPts = 32;   // points per CTA tile
constexpr int kCh = 64;    // channels per CTA tile
constexpr int kThreads = 256;

// max |x| of a block -> one RED.MAX on the device scalar (non-negative floats order as unsigned integers)
__device__ __forceinline__ void block_absmax_to(float m, uint32_t* out) {
    __shared__ float warp_max[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    if ((tid & 31) == 0) warp_max[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        const int n_warps = (blockDim.x * blockDim.y + 31) >> 5;
        for (int w = 1; w < n_warps; ++w) m = fmaxf(m, warp_max[w]);
        // (the plain read may be stale, but the scalar only grows: skipping is safe and keeps thousands of CTAs from
        // serialising on one L2 address)
        if (m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile uint32_t*>(out)) atomicMax(out, __float_as_uint(m));
    }
}

// ---------------------------------------------------------------------------------- splat
// One CTA: a (64 channel x 32 point) tile of x, staged through shared memory so that the
// global read is coalesced along points and the RED traffic is 16-byte vectors along channels.
template <bool I64>
__global__ void __launch_bounds__(kThreads)
scatter_rows_kernel(const float* __restrict__ x, const float* __restrict__ bary,
                    const void* __restrict__ off, long long n_points, int channels,
                    float* __restrict__ rows, long long ld, long long n_rows, float* __restrict__ wsum,
                    uint32_t* __restrict__ in_amax) {
    __shared__ float tile[kCh][kPts + 1];
    __shared__ float s_bary[4][kPts];
    __shared__ int s_off[4][kPts];

    const long long n0 = (long long)blockIdx.x * kPts;
    const int c0 = blockIdx.y * kCh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float x_max = 0.f;
#pragma unroll
    for (int i = 0; i < kCh / 8; ++i) {
        const int c = c0 + warp + 8 * i;
        const long long n = n0 + lane;
        const float v = (c < channels && n < n_points) ? __ldg(x + (long long)c * n_points + n) : 0.f;
        tile[warp + 8 * i][lane] = v;
        x_max = fmaxf(x_max, fabsf(v));
    }
    if (threadIdx.x < 4 * kPts) {
        const int r = threadIdx.x >> 5;
        const long long n = n0 + lane;
        const bool ok = n < n_points;
        s_bary[r][lane] = ok ? __ldg(bary + r * n_points + n) : 0.f;
        int row = ok ? load_idx<I64>(off, r * n_points + n) : -1;
        if (row >= n_rows) row = -1;                     // out-of-range offsets are dropped (the reference would raise)
        s_off[r][lane] = row;
    }
    if (in_amax != nullptr) {                                // fused max|x| (operand-scale bound of the splatted rows), per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x_max = fmaxf(x_max, __shfl_xor_sync(0xffffffffu, x_max, o));
        // (the plain read may be stale, but the scalar only grows: skipping is safe and almost every warp skips)
        if (lane == 0 && x_max > 0.f && __float_as_uint(x_max) > *reinterpret_cast<volatile uint32_t*>(in_amax))
            atomicMax(in_amax, __float_as_uint(x_max));
    }
    __syncthreads();

    if (wsum != nullptr && blockIdx.y == 0 && threadIdx.x < 4 * kPts) {
        const int r = threadIdx.x >> 5;
        const int row = s_off[r][lane];
        if (row >= 0) atomicAdd(wsum + row, s_bary[r][lane]);
    }

    const int grp = threadIdx.x >> 4, j = threadIdx.x & 15;
    if (c0 + 4 * j >= channels) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = grp + 16 * h;
        const float v0 = tile[4 * j + 0][p], v1 = tile[4 * j + 1][p];
        const float v2 = tile[4 * j + 2][p], v3 = tile[4 * j + 3][p];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = s_off[r][p];
            if (row < 0) continue;
            const float b = s_bary[r][p];
            red_add_f32x4(rows + (long long)row * ld + c0 + 4 * j, make_float4(b * v0, b * v1, b * v2, b * v3));
        }
    }
}

// ---------------------------------------------------------------------------------- slice
template <bool I64>
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float* __restrict__ rows, long long ld, const float* __restrict__ bary,
                   const void* __restrict__ off, const float* __restrict__ scale,
                   const float* __restrict__ bias, long long n_points, int channels,
                   long long n_rows, float* __restrict__ y) {
    __shared__ float tile[kCh][kPts + 1];
    __shared__ float s_w[4][kPts];
    __shared__ int s_off[4][kPts];

    const long long n0 = (long long)blockIdx.x * kPts;
    const int c0 = blockIdx.y * kCh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x < 4 * kPts) {
        const int r = threadIdx.x >> 5;
        const long long n = n0 + lane;
        int row = -1;
        float w = 0.f;
        if (n < n_points) {
            row = load_idx<I64>(off, r * n_points + n);
            if (row >= n_rows) row = -1;
            w = __ldg(bary + r * n_points + n);
            if (scale != nullptr && row >= 0) w *= __ldg(scale + row);
        }
        s_w[r][lane] = w;
        s_off[r][lane] = row;
    }
    __syncthreads();

    const int grp = threadIdx.x >> 4, j = threadIdx.x & 15;
    const bool live = c0 + 4 * j < channels;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = grp + 16 * h;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int row = s_off[r][p];
                v[r] = row >= 0 ? __ldg(reinterpret_cast<const float4*>(rows + (long long)row * ld + c0 + 4 * j))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float w = s_w[r][p];
                acc.x = fmaf(w, v[r].x, acc.x);
                acc.y = fmaf(w, v[r].y, acc.y);
                acc.z = fmaf(w, v[r].z, acc.z);
                acc.w = fmaf(w, v[r].w, acc.w);
            }
        }
        tile[4 * j + 0][p] = acc.x;
        tile[4 * j + 1][p] = acc.y;
        tile[4 * j + 2][p] = acc.z;
        tile[4 * j + 3][p] = acc.w;
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < kCh / 8; ++i) {
        const int c = c0 + warp + 8 * i;
        const long long n = n0 + lane;
        if (c < channels && n < n_points)
            y[(long long)c * n_points + n] = tile[warp + 8 * i][lane] + (bias != nullptr ? __ldg(bias + c) : 0.f);
    }
}

// ------------------------------------------------------------------------------ normalise
__global__ void normalize_rows_kernel(float* __restrict__ rows, long long ld, long long n_rows, int quads,
                                      const float* __restrict__ wsum, float* __restrict__ inv, uint32_t* __restrict__ amax) {
    float m = 0.f;
    const long long total = n_rows * quads;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long v = t / quads;
        const int q = (int)(t - v * quads);
        const float s = 1.0f / (wsum[v] + 1e-5f);   // bilateralNN.py:185
        float4* p = reinterpret_cast<float4*>(rows + v * ld) + q;
        float4 a = *p;
        a.x *= s; a.y *= s; a.z *= s; a.w *= s;
        *p = a;
        m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
        // every quad-thread of a row read wsum[v] before any of them can have written inv[v] only
        // if inv does not alias wsum; when it does, the q == quads-1 thread may race with readers.
        // So the in-place case is handled by writing inv in a second kernel (see launcher).
        if (inv != nullptr && q == 0 && inv != wsum) inv[v] = s;
    }
    if (amax != nullptr) block_absmax_to(m, amax);          // fused hpl_absmax of the normalised rows (pads are 0)
}
__global__ void reciprocal_kernel(float* __restrict__ w, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) w[t] = 1.0f / (w[t] + 1e-5f);
}

// --------------------------------------------------------------------------- act backward
__global__ void act_backward_kernel(float* __restrict__ dz, long long ld_dz, const float* __restrict__ y,
                                    long long ld_y, long long n_rows, int quads, float slope) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * quads) return;
    const long long v = t / quads;
    const int q = (int)(t - v * quads);
    float4* pd = reinterpret_cast<float4*>(dz + v * ld_dz) + q;
    const float4 yy = __ldg(reinterpret_cast<const float4*>(y + v * ld_y) + q);
    float4 d = *pd;
    d.x *= yy.x > 0.f ? 1.f : slope;
    d.y *= yy.y > 0.f ? 1.f : slope;
    d.z *= yy.z > 0.f ? 1.f : slope;
    d.w *= yy.w > 0.f ? 1.f : slope;
    *pd = d;
}

// One pass over dz (n_rows, ld): dz *= act'(y) (skipped when y == nullptr), max|dz| -> amax, sum_v dz[v, :] -> colsum.
// Replaces act_backward + hpl_absmax + hpl_column_sums (three passes over the same 4 * H * Co bytes) in the backward of
// every convolution layer.  Block = 64 columns x 4 row lanes over a slab of rows.
template <bool HAS_Y>
__global__ void __launch_bounds__(256)
act_backward_stats_kernel(float* __restrict__ dz, long long ld_dz, const float* __restrict__ y, long long ld_y,
                          long long n_rows, int channels, float slope, long long rows_per_block,
                          uint32_t* __restrict__ amax, float* __restrict__ colsum) {
    __shared__ float part[4][64];
    const int col = threadIdx.x & 63, lane_r = threadIdx.x >> 6;
    const int o = blockIdx.x * 64 + col;
    const long long lo = rows_per_block * blockIdx.y, hi = min(n_rows, lo + rows_per_block);
    float acc = 0.f, m = 0.f;
    if (o < channels) {
        for (long long v = lo + lane_r; v < hi; v += 4) {
            float d = dz[v * ld_dz + o];
            if (HAS_Y) {
                d *= __ldg(y + v * ld_y + o) > 0.f ? 1.f : slope;
                dz[v * ld_dz + o] = d;
            }
            acc += d;
            m = fmaxf(m, fabsf(d));
        }
    }
    part[lane_r][col] = acc;
    __syncthreads();
    if (colsum != nullptr && lane_r == 0 && o < channels && lo < hi)
        atomicAdd(colsum + o, (part[0][col] + part[1][col]) + (part[2][col] + part[3][col]));
    if (amax != nullptr) block_absmax_to(m, amax);
}

// ------------------------------------------------------------------------------ transposes
// cm (C, ld_cm) -> rows (n, ld); pad columns [C, ld) of rows are written as zeros.
__global__ void cm_to_rows_kernel(const float* __restrict__ cm, long long ld_cm, long long n, int channels,
                                  float* __restrict__ rows, long long ld, uint32_t* __restrict__ amax) {
    __shared__ float tile[32][33];
    const long long v0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    float m = 0.f;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long v = v0 + threadIdx.x;
        const float x = (c < channels && v < n) ? __ldg(cm + (long long)c * ld_cm + v) : 0.f;
        tile[i][threadIdx.x] = x;
        m = fmaxf(m, fabsf(x));
    }
    if (amax != nullptr) {
        block_absmax_to(m, amax);                            // (contains the __syncthreads the tile needs)
    } else {
        __syncthreads();
    }
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long v = v0 + i;
        const int c = c0 + threadIdx.x;
        if (v < n && c < ld) rows[v * ld + c] = tile[threadIdx.x][i];
    }
}
__global__ void rows_to_cm_kernel(const float* __restrict__ rows, long long ld, long long n, int channels,
                                  float* __restrict__ cm, long long ld_cm) {
    __shared__ float tile[32][33];
    const long long v0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long v = v0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (v < n && c < channels) ? __ldg(rows + v * ld + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long v = v0 + threadIdx.x;
        if (c < channels && v < n) cm[(long long)c * ld_cm + v] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------------------- channel sums
__global__ void channel_sums_kernel(const float* __restrict__ x, long long n, float* __restrict__ sums) {
    const int c = blockIdx.x;
    const long long chunk = (n + gridDim.y - 1) / gridDim.y;
    const long long lo = chunk * blockIdx.y, hi = min(n, lo + chunk);
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += __ldg(x + (long long)c * n + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
        atomicAdd(sums + c, s);
    }
}

// ------------------------------------------------------------------------- table transpose
template <bool I64>
__global__ void transpose_table_kernel(const void* __restrict__ tbl, int filter_size, long long n_rows,
                                       int* __restrict__ tbl_t, long long n_src_rows, int* collisions) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)filter_size * n_rows) return;
    const long long f = t / n_rows, v = t - f * n_rows;
    const int u = load_idx<I64>(tbl, t);
    if (u < 0 || u >= n_src_rows) return;
    const int old = atomicCAS(tbl_t + f * n_src_rows + u, -1, (int)v);
    if (old != -1 && collisions != nullptr) atomicAdd(collisions, 1);
}

__global__ void fill_i32_kernel(int* p, long long n, int v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}

inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

}  // namespace

extern "C" {

int hpl_version(void) { return 100; }
int hpl_sm_arch(void) {
#ifdef HPL_SM_ARCH
    return HPL_SM_ARCH;
#else
    return 100;
#endif
}

int hpl_scatter_rows(const float* x, const float* bary, const void* off, int idx64, int64_t n_points,
                     int64_t channels, float* rows, int64_t ld, int64_t n_rows, float* wsum, uint32_t* in_amax,
                     void* stream) {
    HPL_CHECK_ARG(x && bary && off && rows && n_points >= 0 && channels > 0 && n_rows >= 0);
    HPL_CHECK_ARG(ld % 4 == 0 && ld >= channels && ((uintptr_t)rows & 15) == 0);
    if (n_points == 0) return 0;
    dim3 grid(blocks_for(n_points, kPts), blocks_for(channels, kCh));
    if (idx64)
        scatter_rows_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(x, bary, off, n_points, (int)channels, rows, ld, n_rows, wsum, in_amax);
    else
        scatter_rows_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(x, bary, off, n_points, (int)channels, rows, ld, n_rows, wsum, in_amax);
    HPL_RETURN_LAST();
}

int hpl_normalize_rows(float* rows, int64_t ld, int64_t n_rows, int64_t channels, const float* wsum,
                       float* inv, void* stream) {
    return hpl_normalize_rows_amax(rows, ld, n_rows, channels, wsum, inv, nullptr, stream);
}

int hpl_normalize_rows_amax(float* rows, int64_t ld, int64_t n_rows, int64_t channels, const float* wsum,
                            float* inv, uint32_t* amax, void* stream) {
    HPL_CHECK_ARG(wsum);
    if (n_rows == 0) return 0;
    if (rows == nullptr) {   // reciprocal only
        HPL_CHECK_ARG(inv == wsum);
        reciprocal_kernel<<<blocks_for(n_rows, 256), 256, 0, as_stream(stream)>>>(inv, n_rows);
        HPL_RETURN_LAST();
    }
    HPL_CHECK_ARG(ld % 4 == 0 && ld >= channels && ((uintptr_t)rows & 15) == 0);
    const int quads = (int)((channels + 3) / 4);
    unsigned blocks = blocks_for(n_rows * quads, 256);
    if (blocks > 16u * num_sms()) blocks = 16u * num_sms();            // grid-stride: one RED.MAX per block
    normalize_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(rows, ld, n_rows, quads, wsum, inv, amax);
    if (inv != nullptr && inv == wsum)
        reciprocal_kernel<<<blocks_for(n_rows, 256), 256, 0, as_stream(stream)>>>(inv, n_rows);
    HPL_RETURN_LAST();
}

int hpl_gather_rows(const float* rows, int64_t ld, const float* bary, const void* off, int idx64,
                    const float* scale, const float* bias, int64_t n_points, int64_t channels, int64_t n_rows,
                    float* y, void* stream) {
    HPL_CHECK_ARG(rows && bary && off && y && n_points >= 0 && channels > 0 && n_rows >= 0);
    HPL_CHECK_ARG(ld % 4 == 0 && ld >= channels && ((uintptr_t)rows & 15) == 0);
    if (n_points == 0) return 0;
    dim3 grid(blocks_for(n_points, kPts), blocks_for(channels, kCh));
    if (idx64)
        gather_rows_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(rows, ld, bary, off, scale, bias, n_points, (int)channels, n_rows, y);
    else
        gather_rows_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(rows, ld, bary, off, scale, bias, n_points, (int)channels, n_rows, y);
    HPL_RETURN_LAST();
}

int hpl_act_backward(float* dz, int64_t ld_dz, const float* y, int64_t ld_y, int64_t n_rows, int64_t channels,
                     int act, void* stream) {
    HPL_CHECK_ARG(dz && y && ld_dz % 4 == 0 && ld_y % 4 == 0 && ld_dz >= channels && ld_y >= channels);
    if (act == HPL_ACT_NONE || n_rows == 0) return 0;
    const int quads = (int)((channels + 3) / 4);
    const float slope = act == HPL_ACT_LEAKY ? HPL_LEAKY_RATE : 0.f;
    act_backward_kernel<<<blocks_for(n_rows * quads, 256), 256, 0, as_stream(stream)>>>(dz, ld_dz, y, ld_y, n_rows, quads, slope);
    HPL_RETURN_LAST();
}

int hpl_act_backward_stats(float* dz, int64_t ld_dz, const float* y, int64_t ld_y, int64_t n_rows, int64_t channels,
                           int act, uint32_t* amax, float* colsum, void* stream) {
    HPL_CHECK_ARG(dz && ld_dz >= channels && channels > 0 && (act == HPL_ACT_NONE || (y && ld_y >= channels)));
    if (n_rows == 0) return 0;
    long long blocks_y = (n_rows + 127) / 128;
    if (blocks_y > 2048) blocks_y = 2048;
    const long long rpb = (n_rows + blocks_y - 1) / blocks_y;
    dim3 grid((unsigned)((channels + 63) / 64), (unsigned)blocks_y);
    const float slope = act == HPL_ACT_LEAKY ? HPL_LEAKY_RATE : 0.f;
    if (act == HPL_ACT_NONE)
        act_backward_stats_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(dz, ld_dz, nullptr, 0, n_rows, (int)channels, 1.f, rpb, amax, colsum);
    else
        act_backward_stats_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(dz, ld_dz, y, ld_y, n_rows, (int)channels, slope, rpb, amax, colsum);
    HPL_RETURN_LAST();
}

int hpl_transpose_table(const void* tbl, int idx64, int64_t filter_size, int64_t n_rows, int32_t* tbl_t,
                        int64_t n_src_rows, int32_t* collisions, void* stream) {
    HPL_CHECK_ARG(tbl && tbl_t && filter_size > 0);
    if (n_rows == 0) return 0;
    const unsigned g = blocks_for(filter_size * n_rows, 256);
    if (idx64)
        transpose_table_kernel<true><<<g, 256, 0, as_stream(stream)>>>(tbl, (int)filter_size, n_rows, tbl_t, n_src_rows, collisions);
    else
        transpose_table_kernel<false><<<g, 256, 0, as_stream(stream)>>>(tbl, (int)filter_size, n_rows, tbl_t, n_src_rows, collisions);
    HPL_RETURN_LAST();
}

int hpl_cm_to_rows(const float* cm, int64_t ld_cm, int64_t n, int64_t channels, float* rows, int64_t ld,
                   void* stream) {
    return hpl_cm_to_rows_amax(cm, ld_cm, n, channels, rows, ld, nullptr, stream);
}

int hpl_cm_to_rows_amax(const float* cm, int64_t ld_cm, int64_t n, int64_t channels, float* rows, int64_t ld,
                        uint32_t* amax, void* stream) {
    HPL_CHECK_ARG(cm && rows && ld >= channels && ld_cm >= n);
    if (n == 0) return 0;
    dim3 grid(blocks_for(n, 32), blocks_for(ld, 32)), block(32, 8);
    cm_to_rows_kernel<<<grid, block, 0, as_stream(stream)>>>(cm, ld_cm, n, (int)channels, rows, ld, amax);
    HPL_RETURN_LAST();
}

int hpl_rows_to_cm(const float* rows, int64_t ld, int64_t n, int64_t channels, float* cm, int64_t ld_cm,
                   void* stream) {
    HPL_CHECK_ARG(cm && rows && ld >= channels && ld_cm >= n);
    if (n == 0) return 0;
    dim3 grid(blocks_for(n, 32), blocks_for(channels, 32)), block(32, 8);
    rows_to_cm_kernel<<<grid, block, 0, as_stream(stream)>>>(rows, ld, n, (int)channels, cm, ld_cm);
    HPL_RETURN_LAST();
}

int hpl_channel_sums(const float* x, int64_t channels, int64_t n, float* sums, void* stream) {
    HPL_CHECK_ARG(x && sums && channels > 0);
    if (n == 0) return 0;
    int split = (int)((n + 8191) / 8192);
    if (split < 1) split = 1;
    if (split > 64) split = 64;
    dim3 grid((unsigned)channels, split);
    channel_sums_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n, sums);
    HPL_RETURN_LAST();
}

int hpl_fill_zero(void* ptr, int64_t bytes, void* stream) {
    if (bytes == 0) return 0;
    HPL_CHECK_ARG(ptr);
    return (int)cudaMemsetAsync(ptr, 0, (size_t)bytes, as_stream(stream));
}

int hpl_fill_i32(int32_t* ptr, int64_t count, int32_t value, void* stream) {
    if (count == 0) return 0;
    HPL_CHECK_ARG(ptr);
    fill_i32_kernel<<<blocks_for(count, 256), 256, 0, as_stream(stream)>>>(ptr, count, value);
    HPL_RETURN_LAST();
}

}  // extern "C"
