// Gather-GEMM on CUDA cores (fp32 FMA) -- the parity anchor of the learned lattice convolution
// (models/bilateralNN.py:198-221): out[v,:] = act(b + sum_f in[nbr[f,v],:] . W_f).
//
// The reference materialises the F-times gathered copy (bilateralNN.py:215-217) and hands it
// to cuDNN.  Here the A operand is gathered row by row straight into shared memory (one
// lattice row = one contiguous burst), so the copy never exists in HBM.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int kGemmThreads = 128;     // 16 (m) x 8 (n) threads, 8x8 outputs each
constexpr int APAD = 4;

template <bool I64>
__device__ __forceinline__ int src_row(const void* nbr, int f, long long n_out_rows, long long v, long long n_in_rows) {
    if (v >= n_out_rows) return -1;
    if (nbr == nullptr) return v < n_in_rows ? (int)v : -1;
    const int r = load_idx<I64>(nbr, (long long)f * n_out_rows + v);
    return r < n_in_rows ? r : -1;
}

template <bool I64>
__global__ void __launch_bounds__(kGemmThreads)
gather_gemm_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows,
                   const void* __restrict__ nbr, int filter_size, long long n_out_rows, int c_in, int c_out,
                   const float* __restrict__ w, const float* __restrict__ bias, int act,
                   float* __restrict__ out, long long ld_out, int out_cm) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int t = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = t & 7, ty = t >> 3;

    const int n_cc = (c_in + BK - 1) / BK;          // channel chunks per filter tap
    const int steps = filter_size * n_cc;

    // A loader: thread covers rows (t>>2) + 32*i, 16-byte quad (t&3) of the 64-byte chunk.
    const int a_kq = t & 3;
    int a_row[4];
    float4 a_reg[4], b_reg[2];

    auto load_rows = [&](int f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a_row[i] = src_row<I64>(nbr, f, n_out_rows, m0 + (t >> 2) + 32 * i, n_in_rows);
    };
    auto load_global = [&](int s) {
        const int f = s / n_cc, cc = (s - f * n_cc) * BK;
        const int c = cc + 4 * a_kq;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            a_reg[i] = (a_row[i] >= 0 && c < c_in)
                           ? __ldg(reinterpret_cast<const float4*>(in + (long long)a_row[i] * ld_in + c))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        // B: rows k = cc..cc+15 of w[f] (c_in x c_out), 64 columns from n0
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int e = t + kGemmThreads * i;   // 0..255
            const int k = e >> 4, q = e & 15;
            const int cr = cc + k, n = n0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cr < c_in) {
                const float* p = w + ((long long)f * c_in + cr) * c_out + n;
                if (n + 3 < c_out && (c_out & 3) == 0) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (n + 0 < c_out) v.x = __ldg(p + 0);
                    if (n + 1 < c_out) v.y = __ldg(p + 1);
                    if (n + 2 < c_out) v.z = __ldg(p + 2);
                    if (n + 3 < c_out) v.w = __ldg(p + 3);
                }
            }
            b_reg[i] = v;
        }
    };
    auto store_shared = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = (t >> 2) + 32 * i;
            As[buf][4 * a_kq + 0][m] = a_reg[i].x;
            As[buf][4 * a_kq + 1][m] = a_reg[i].y;
            As[buf][4 * a_kq + 2][m] = a_reg[i].z;
            As[buf][4 * a_kq + 3][m] = a_reg[i].w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int e = t + kGemmThreads * i;
            *reinterpret_cast<float4*>(&Bs[buf][e >> 4][4 * (e & 15)]) = b_reg[i];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_rows(0);
    load_global(0);
    store_shared(0);
    __syncthreads();

    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        if (s + 1 < steps) {
            if ((s + 1) % n_cc == 0) load_rows((s + 1) / n_cc);
            load_global(s + 1);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][32 + 4 * tx]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (s + 1 < steps) {
            store_shared(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: bias + activation, vertex-major or channel-major store
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
        const int nb = n0 + 32 * jh + 4 * tx;
        float bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = (bias != nullptr && nb + j < c_out) ? __ldg(bias + nb + j) : 0.f;
#pragma unroll
        for (int ih = 0; ih < 2; ++ih)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long m = m0 + 64 * ih + 4 * ty + i;
                if (m >= n_out_rows) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[4 * ih + i][4 * jh + j] + bv[j], act);
                if (!out_cm) {
                    float* p = out + m * ld_out + nb;
                    if (nb + 3 < c_out && (ld_out & 3) == 0 && ((uintptr_t)out & 15) == 0) {
                        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (nb + j < c_out) p[j] = v[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (nb + j < c_out) out[(long long)(nb + j) * ld_out + m] = v[j];
                }
            }
    }
}

// ------------------------------------------------------------------------------- wgrad
// dw[f] (c_in x c_out) += gathered(in)^T . dz over a vertex range.  Tile 64 x 64, k = 16 vertices.
constexpr int WM = 64, WN = 64, WK = 16;
constexpr int kWgradThreads = 256;    // 16 x 16 threads, 4x4 outputs each

template <bool I64>
__global__ void __launch_bounds__(kWgradThreads)
wgrad_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows, const void* __restrict__ nbr,
             long long n_out_rows, int c_in, int c_out, const float* __restrict__ dz, long long ld_dz,
             float* __restrict__ dw, int tiles_n, long long rows_per_split) {
    __shared__ __align__(16) float As[2][WK][WM];
    __shared__ __align__(16) float Bs[2][WK][WN];

    const int t = threadIdx.x;
    const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x - tile_m * tiles_n;
    const int c0 = tile_m * WM, o0 = tile_n * WN;
    const int f = blockIdx.y;
    const long long v_lo = rows_per_split * blockIdx.z;
    const long long v_hi = min(n_out_rows, v_lo + rows_per_split);
    if (v_lo >= v_hi) return;
    const int steps = (int)((v_hi - v_lo + WK - 1) / WK);

    const int lk = t >> 4, lq = t & 15;   // loader: vertex lk of the step, quad lq
    float4 a_reg, b_reg;
    auto load_global = [&](int s) {
        const long long v = v_lo + (long long)s * WK + lk;
        a_reg = make_float4(0.f, 0.f, 0.f, 0.f);
        b_reg = a_reg;
        if (v < v_hi) {
            const int row = src_row<I64>(nbr, f, n_out_rows, v, n_in_rows);
            const int c = c0 + 4 * lq;
            if (row >= 0 && c < c_in) a_reg = __ldg(reinterpret_cast<const float4*>(in + (long long)row * ld_in + c));
            const int o = o0 + 4 * lq;
            if (o < c_out) b_reg = __ldg(reinterpret_cast<const float4*>(dz + v * ld_dz + o));
        }
    };
    auto store_shared = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][lk][4 * lq]) = a_reg;
        *reinterpret_cast<float4*>(&Bs[buf][lk][4 * lq]) = b_reg;
    };

    const int tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_global(0);
    store_shared(0);
    __syncthreads();
    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        if (s + 1 < steps) load_global(s + 1);
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (s + 1 < steps) {
            store_shared(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + 4 * ty + i;
        if (c >= c_in) continue;
        float* p = dw + ((long long)f * c_in + c) * c_out + o0 + 4 * tx;
        const int o = o0 + 4 * tx;
        if (o + 3 < c_out && (c_out & 3) == 0) {
            red_add_f32x4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (o + j < c_out) atomicAdd(p + j, acc[i][j]);
        }
    }
}

// db[o] += sum_v dz[v, o].  Block = 64 columns x 4 row lanes; each block reduces a slab of rows.
__global__ void __launch_bounds__(256)
column_sums_kernel(const float* __restrict__ dz, long long ld_dz, long long n_rows, int c_out,
                   float* __restrict__ db, long long rows_per_block) {
    __shared__ float part[4][64];
    const int col = threadIdx.x & 63, lane_r = threadIdx.x >> 6;
    const int o = blockIdx.x * 64 + col;
    const long long lo = rows_per_block * blockIdx.y, hi = min(n_rows, lo + rows_per_block);
    float acc0 = 0.f, acc1 = 0.f;
    if (o < c_out) {
        long long v = lo + lane_r;
        for (; v + 4 < hi; v += 8) {
            acc0 += __ldg(dz + v * ld_dz + o);
            acc1 += __ldg(dz + (v + 4) * ld_dz + o);
        }
        if (v < hi) acc0 += __ldg(dz + v * ld_dz + o);
    }
    part[lane_r][col] = acc0 + acc1;
    __syncthreads();
    if (lane_r == 0 && o < c_out && lo < hi) atomicAdd(db + o, (part[0][col] + part[1][col]) + (part[2][col] + part[3][col]));
}

}  // namespace

extern "C" {

int hpl_blur_gemm(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                  int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* w,
                  const float* bias, int act, float* out, int64_t ld_out, int out_channel_major, int precision,
                  void* stream) {
    HPL_CHECK_ARG(in && w && out && c_in > 0 && c_out > 0 && filter_size > 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    HPL_CHECK_ARG(out_channel_major ? ld_out >= n_out_rows : ld_out >= c_out);
    HPL_CHECK_ARG(precision == 0);
    if (n_out_rows == 0) return 0;
    dim3 grid((unsigned)((n_out_rows + BM - 1) / BM), (unsigned)((c_out + BN - 1) / BN));
    if (idx64)
        gather_gemm_kernel<true><<<grid, kGemmThreads, 0, as_stream(stream)>>>(
            in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, w, bias, act, out, ld_out,
            out_channel_major);
    else
        gather_gemm_kernel<false><<<grid, kGemmThreads, 0, as_stream(stream)>>>(
            in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, w, bias, act, out, ld_out,
            out_channel_major);
    HPL_RETURN_LAST();
}

int hpl_column_sums(const float* rows, int64_t ld, int64_t n_rows, int64_t channels, float* sums, void* stream) {
    HPL_CHECK_ARG(rows && sums && ld >= channels && channels > 0);
    if (n_rows == 0) return 0;
    long long blocks_y = (n_rows + 127) / 128;
    if (blocks_y > 4096) blocks_y = 4096;
    const long long rpb = (n_rows + blocks_y - 1) / blocks_y;
    dim3 g2((unsigned)((channels + 63) / 64), (unsigned)blocks_y);
    column_sums_kernel<<<g2, 256, 0, as_stream(stream)>>>(rows, ld, n_rows, (int)channels, sums, rpb);
    HPL_RETURN_LAST();
}

int hpl_blur_wgrad(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                   int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* dz,
                   int64_t ld_dz, float* dw, float* db, void* stream) {
    HPL_CHECK_ARG(in && dz && dw && c_in > 0 && c_out > 0 && filter_size > 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0);
    HPL_CHECK_ARG(ld_dz % 4 == 0 && ld_dz >= c_out && ((uintptr_t)dz & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    if (n_out_rows == 0) return 0;
    const int tiles_m = (int)((c_in + WM - 1) / WM), tiles_n = (int)((c_out + WN - 1) / WN);
    const long long base = (long long)tiles_m * tiles_n * filter_size;
    long long splits = (4LL * num_sms() + base - 1) / base;
    const long long max_splits = (n_out_rows + 8 * WK - 1) / (8 * WK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long rows_per_split = (n_out_rows + splits - 1) / splits;
    rows_per_split = (rows_per_split + WK - 1) / WK * WK;
    splits = (n_out_rows + rows_per_split - 1) / rows_per_split;
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)filter_size, (unsigned)splits);
    if (idx64)
        wgrad_kernel<true><<<grid, kWgradThreads, 0, as_stream(stream)>>>(in, ld_in, n_in_rows, nbr, n_out_rows, (int)c_in,
                                                                          (int)c_out, dz, ld_dz, dw, tiles_n, rows_per_split);
    else
        wgrad_kernel<false><<<grid, kWgradThreads, 0, as_stream(stream)>>>(in, ld_in, n_in_rows, nbr, n_out_rows, (int)c_in,
                                                                           (int)c_out, dz, ld_dz, dw, tiles_n, rows_per_split);
    if (db != nullptr) return hpl_column_sums(dz, ld_dz, n_out_rows, c_out, db, stream);
    HPL_RETURN_LAST();
}

}  // extern "C"
