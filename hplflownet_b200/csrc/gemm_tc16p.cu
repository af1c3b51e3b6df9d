// "Split once, copy many": tcgen05 3xFP16 contractions whose operands are PRE-SPLIT in HBM.
//
// ncu on gemm_tc16.cu (the register-staged variant) showed the contraction kernels bound by
// instruction issue (56-68 % of peak, tensor pipe 15-20 %): every gathered element is scaled and split
// into fp16 hi/lo (~5 instructions) once per filter tap, i.e. 15 times.  Here the split happens once per
// tensor (hpl_split16: x -> [hi | lo] fp16 planes, scale from hpl_absmax), and the contraction producers
// only COPY: cp.async (LDGSTS, 16 bytes, zero-fill for missing neighbours) straight from HBM/L2 into the
// UMMA operand layout in shared memory -- no registers, ~25 instructions per thread and stage instead of
// ~300, prefetch depth bounded by the shared-memory ring instead of the register file.
//
// x16 layout (one row per lattice vertex, same 4 bytes per element as fp32): for every block of 32
// channels a 128-byte line [32 hi halves | 32 lo halves].  One line = everything one row contributes to one
// 32-wide K block (forward) or to 4+4 MN chunks (weight gradient), so a quarter warp copies exactly one
// line.  Operand strides are padded (LBO / SBO + 32 B, lo plane + 16 B) so that the 8 chunks of a line land
// in 8 different 16-byte bank groups.
//
// STATUS (measured on B200, cfg2 x 32 clouds): parity-green (tests/test_gpu_gemm.py precision = 3) but not
// faster than the register-staged kernels yet -- forward 0.30 ms vs 0.24 ms, weight gradient 0.44 vs 0.33 ms --
// so ops.DEFAULT_PRECISION stays 2.  The copy producers are cheap (72 registers, ~25 instructions per stage)
// but the kernel is then paced by the full/empty round trip of a 4-stage ring; next step is a deeper A ring
// with a separate B ring (or a persistent CTA), see DESIGN.md section 8.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TM = 128, TN = 64, TK = 32;
constexpr int kStages = 4;
constexpr int kProducerWarps = 8;
constexpr int kThreads = kProducerWarps * 32 + 64;
constexpr int kLookahead = 2;                       // cp.async groups in flight per thread; kStages - kLookahead slots of slack
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

// ---- forward operand geometry (K-major, no swizzle)
constexpr uint32_t kA_LBO = TM * 16 + 32;           // 2080
constexpr uint32_t kALoOff = (TK / 8) * kA_LBO + 16; // lo plane of the A stage: 8336 (== 16 mod 128)
constexpr uint32_t kABytes = 2 * (TK / 8) * kA_LBO + 128;   // hi + lo (+ pad) = 16768
constexpr uint32_t kB_LBO = TN * 16, kSBO = 128;
constexpr int kBHalf = TN * TK * 2;                 // 4 KB
constexpr int kStageBytes = kABytes + 2 * kBHalf;   // 24960
constexpr int kSmemBytes = kStages * kStageBytes + 1024;
constexpr uint32_t kIdescK = instr_desc(0, TM, TN, 0, 0);
constexpr uint32_t kIdescMN = instr_desc(0, TM, TN, 1, 1);

__device__ __forceinline__ void scale_from_amax(uint32_t bits, float& scale, float& inv_scale) {
    int e = (int)((bits >> 23) & 0xff) - 127;
    if (bits == 0) e = 13;
    int se = e - 13;
    se = se < -100 ? -100 : (se > 100 ? 100 : se);
    scale = __uint_as_float((uint32_t)(se + 127) << 23);
    inv_scale = __uint_as_float((uint32_t)(127 - se) << 23);
}
__device__ __forceinline__ void split4h(const float4 a, float inv_s, uint32_t* hi, uint32_t* lo) {
    const float x[4] = {a.x * inv_s, a.y * inv_s, a.z * inv_s, a.w * inv_s};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((x[2 * i] - f.x) * kLoScale, (x[2 * i + 1] - f.y) * kLoScale);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
// 16-byte async copy global -> shared; src_bytes = 0 writes zeros (missing neighbour)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------- split
// x (rows, ld) fp32 -> x16 (rows, blocks * 128 bytes): per 32-channel block [32 hi | 32 lo] halves.
__global__ void split16_kernel(const float* __restrict__ x, long long ld, long long n_rows, int channels, int blocks,
                               const uint32_t* __restrict__ amax, uint8_t* __restrict__ x16) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread = 4 channels
    const int quads = blocks * 8;
    if (t >= n_rows * quads) return;
    const long long row = t / quads;
    const int q = (int)(t - row * quads), c = 4 * q;
    float s, inv_s;
    scale_from_amax(__ldg(amax), s, inv_s);
    const float4 v = c < channels ? __ldg(reinterpret_cast<const float4*>(x + row * ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 vv = v;
    if (c + 1 >= channels) vv.y = 0.f;      // ragged tail (channels % 4 != 0): pad columns are zero by convention anyway
    if (c + 2 >= channels) vv.z = 0.f;
    if (c + 3 >= channels) vv.w = 0.f;
    uint32_t hi[2], lo[2];
    split4h(vv, inv_s, hi, lo);
    uint8_t* line = x16 + (row * blocks + (q >> 3)) * 128 + (q & 7) * 8;
    *reinterpret_cast<uint2*>(line) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(line + 64) = make_uint2(lo[0], lo[1]);
}

// weight image (same as gemm_tc16.cu): w (F, C, Co) -> per (N tile, K block) [hi 4 KB | lo 4 KB]
__global__ void weight_image16p_kernel(const float* __restrict__ w, int filter_size, int c_in, int c_out, int kb_per_tap,
                                       const uint32_t* __restrict__ w_amax, uint8_t* __restrict__ image) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int chunks = TN * (TK / 8);
    const long long n_kb = (long long)filter_size * kb_per_tap;
    const long long n_tiles = (c_out + TN - 1) / TN;
    if (t >= n_tiles * n_kb * chunks) return;
    float s, inv_s;
    scale_from_amax(*w_amax, s, inv_s);
    const int chunk = (int)(t % chunks);
    const long long blk = t / chunks;
    const long long kb = blk % n_kb, tile = blk / n_kb;
    const int f = (int)(kb / kb_per_tap), c0 = (int)(kb % kb_per_tap) * TK;
    const int n = chunk & (TN - 1), kc = chunk / TN;
    const int o = (int)tile * TN + n;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + 8 * kc + i;
        v[i] = (c < c_in && o < c_out) ? __ldg(w + ((long long)f * c_in + c) * c_out + o) : 0.f;
    }
    uint32_t hi[4], lo[4];
    split4h(make_float4(v[0], v[1], v[2], v[3]), inv_s, hi, lo);
    split4h(make_float4(v[4], v[5], v[6], v[7]), inv_s, hi + 2, lo + 2);
    uint8_t* dst = image + blk * (2 * kBHalf) + kc * (TN * 16) + (n >> 3) * 128 + (n & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + kBHalf) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------ forward / dgrad
template <bool I64>
__global__ void __launch_bounds__(kThreads, 2)
gather_gemm_p16_kernel(const uint8_t* __restrict__ in16, int blocks_in, long long n_in_rows, const void* __restrict__ nbr,
                       int filter_size, long long n_out_rows, int c_out, const uint8_t* __restrict__ w_image,
                       const float* __restrict__ bias, int act, float* __restrict__ out, long long ld_out, int out_cm, int n_main,
                       const uint32_t* __restrict__ in_amax, const uint32_t* __restrict__ w_amax) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n_tile = blockIdx.y;
    const int kb_per_tap = blocks_in;                        // one 32-channel block per K block
    const int n_kb = filter_size * kb_per_tap;
    const uint32_t tmem_cols = (uint32_t)(TN * (n_main + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], kProducerWarps + 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == kProducerWarps) tmem_alloc(&tmem_slot, tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), accum_a = smem_u32(&accum_bar);
    const bool lane0 = lane == 0;

    if (warp < kProducerWarps) {
        // lane = (row within a group of 4, 16-byte chunk of the row's 128-byte line): chunks 0-3 hi, 4-7 lo
        const int rq = lane >> 3, c8 = lane & 7;
        const long long row_bytes = (long long)blocks_in * 128;
        const uint8_t* rowp[4];                              // line of the current tap's row (+ chunk offset), or NULL
        int row_next[4];
        uint32_t off[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ml = warp * 16 + b * 4 + rq;
            off[b] = (c8 >> 2) * kALoOff + (c8 & 3) * kA_LBO + (ml >> 3) * 128 + (ml & 7) * 16;
        }
        const long long v_first = m0 + warp * 16 + rq;
        auto fetch_rows = [&](int f) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const long long v = v_first + b * 4;
                int r = -1;
                if (v < n_out_rows && f < filter_size) r = nbr != nullptr ? load_idx<I64>(nbr, (long long)f * n_out_rows + v) : (int)v;
                row_next[b] = r;
            }
        };
        int tap_i = 0, kt_i = 0;
        fetch_rows(0);
        int stage_w = 0, stage_r = 0;
        uint32_t phase_w = 0;
        for (int kb = 0; kb < n_kb + kLookahead; ++kb) {
            if (kb >= kLookahead) {                          // publish stage kb - kLookahead (its group is the oldest pending one)
                cp_async_wait<kLookahead - 1>();
                fence_proxy_async();
                __syncwarp();
                if (lane0) mbar_arrive_a(full_a + 8 * stage_r);
                if (++stage_r == kStages) stage_r = 0;
            }
            if (kb < n_kb) {
                if (kt_i == 0) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        rowp[b] = (row_next[b] >= 0 && row_next[b] < n_in_rows) ? in16 + (long long)row_next[b] * row_bytes + c8 * 16 : nullptr;
                    fetch_rows(tap_i + 1);
                }
                if (lane0) mbar_wait_a(empty_a + 8 * stage_w, phase_w ^ 1);     // slot freed kStages - kLookahead MMAs ago
                __syncwarp();
                const uint32_t dst = smem_base + stage_w * kStageBytes;
                const long long boff = (long long)kt_i * 128;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const bool ok = rowp[b] != nullptr;
                    cp_async16(dst + off[b], ok ? (const void*)(rowp[b] + boff) : (const void*)in16, ok ? 16u : 0u);
                }
                if (++kt_i == kb_per_tap) { kt_i = 0; ++tap_i; }
                if (++stage_w == kStages) { stage_w = 0; phase_w ^= 1; }
            }
            cp_async_commit();
        }
    } else if (warp == kProducerWarps) {
        if (lane == 0) {
            int last_g = -1, stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int g = (int)((long long)kb * n_main / n_kb);
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait_a(full_a + 8 * stage, phase);
                fence_after();
                const uint32_t a_hi = smem_base + stage * kStageBytes;
                const uint32_t a_lo = a_hi + kALoOff, b_hi = a_hi + kABytes, b_lo = b_hi + kBHalf;
#pragma unroll
                for (int j = 0; j < TK / 16; ++j) {
                    const uint64_t dah = smem_desc(a_hi + j * 2 * kA_LBO, kA_LBO, kSBO);
                    const uint64_t dal = smem_desc(a_lo + j * 2 * kA_LBO, kA_LBO, kSBO);
                    const uint64_t dbh = smem_desc(b_hi + j * 2 * kB_LBO, kB_LBO, kSBO);
                    const uint64_t dbl = smem_desc(b_lo + j * 2 * kB_LBO, kB_LBO, kSBO);
                    umma_f16(tmem_d, dal, dbh, kIdescK, (kb | j) != 0);
                    umma_f16(tmem_d, dah, dbl, kIdescK, 1);
                    umma_f16(tmem_main, dah, dbh, kIdescK, g == last_g);
                    last_g = g;
                }
                umma_commit_a(empty_a + 8 * stage);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_a(accum_a);
        }
    } else {
        if (lane == 0) {
            const uint8_t* src = w_image + (long long)n_tile * n_kb * (2 * kBHalf);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait_a(empty_a + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx_a(full_a + 8 * stage, 2 * kBHalf);
                bulk_load_a(smem_base + stage * kStageBytes + kABytes, src, 2 * kBHalf, full_a + 8 * stage);
                src += 2 * kBHalf;
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }

    if (warp < 4) {
        float s_in, inv_in, s_w, inv_w;
        scale_from_amax(__ldg(in_amax), s_in, inv_in);
        scale_from_amax(__ldg(w_amax), s_w, inv_w);
        mbar_wait_a(accum_a, 0);
        fence_after();
        const long long m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const int o0 = n_tile * TN;
        const float s_ab = s_in * s_w;
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            float sum[16];
            uint32_t v[16];
            tmem_ld16(taddr + cb, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]) * kLoInv;
            for (int g = 1; g <= n_main; ++g) {
                tmem_ld16(taddr + g * TN + cb, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (m < n_out_rows) {
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int o = o0 + cb + j;
                    const float b = (bias != nullptr && o < c_out) ? __ldg(bias + o) : 0.f;
                    y[j] = apply_act(fmaf(sum[j], s_ab, b), act);
                }
                if (!out_cm) {
                    float* p = out + m * ld_out + o0 + cb;
                    if (o0 + cb + 15 < c_out && (ld_out & 3) == 0 && ((uintptr_t)out & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (o0 + cb + j < c_out) p[j] = y[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (o0 + cb + j < c_out) out[(long long)(o0 + cb + j) * ld_out + m] = y[j];
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == kProducerWarps) {
        fence_after();
        tmem_dealloc(tmem_d, tmem_cols);
    }
}

// ------------------------------------------------------------------------------ weight gradient
// dw[(f,c), o] += sum_v in[nbr[f,v], c] * dz[v, o];  operands MN-major no-swizzle fp16, copied from the
// pre-split planes: element (m, k) at (k / 8) * LBO + (m / 8) * SBO + (k % 8) * 16 + (m % 8) * 2.
// Requires c_in % 32 == 0 (an M tile is 4 whole 32-channel blocks).
constexpr int WG_MAIN = 3;
constexpr int kWStages = 3;   // 31 KB each -> 2 CTAs / SM
constexpr uint32_t kW_SBO = 128 + 32;
constexpr uint32_t kWA_LBO = (TM / 8) * kW_SBO;                 // 2560
constexpr uint32_t kWB_LBO = (TN / 8) * kW_SBO;                 // 1280
constexpr uint32_t kWALoOff = (TK / 8) * kWA_LBO + 16;          // 10256
constexpr uint32_t kWABytes = 2 * (TK / 8) * kWA_LBO + 128;     // 20608
constexpr uint32_t kWBLoOff = (TK / 8) * kWB_LBO + 16;          // 5136
constexpr uint32_t kWBBytes = 2 * (TK / 8) * kWB_LBO + 128;     // 10368
constexpr int kWStageBytes = kWABytes + kWBBytes;               // 30976
constexpr int kWSmemBytes = kWStages * kWStageBytes + 1024;
constexpr int kWLookahead = 2;

template <bool I64>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_p16_kernel(const uint8_t* __restrict__ in16, int blocks_in, long long n_in_rows, const void* __restrict__ nbr, int filter_size,
                 long long n_out_rows, int c_in, int c_out, const uint8_t* __restrict__ dz16, int blocks_dz, float* __restrict__ dw,
                 long long rows_per_split, const uint32_t* __restrict__ in_amax, const uint32_t* __restrict__ dz_amax) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kWStages], empty_bar[kWStages], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TM, o0 = blockIdx.z * TN;
    const int m_total = filter_size * c_in;
    const long long v_lo = rows_per_split * blockIdx.x;
    const long long v_hi = min(n_out_rows, v_lo + rows_per_split);
    const int n_kb = v_lo < v_hi ? (int)((v_hi - v_lo + TK - 1) / TK) : 0;
    const uint32_t tmem_cols = (uint32_t)(TN * (WG_MAIN + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(&full_bar[s], kProducerWarps);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == kProducerWarps) tmem_alloc(&tmem_slot, tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), accum_a = smem_u32(&accum_bar);
    const bool lane0 = lane == 0;

    if (warp < kProducerWarps) {
        // warp w owns vertex quad w of the stage; lane = (vertex within the quad, chunk of the 128-byte line)
        const int rq = lane >> 3, c8 = lane & 7;
        const int kk = warp * 4 + rq;                                  // vertex inside the stage
        const int span = (int)(v_hi - v_lo);
        // destination of chunk c8 of block mb: plane (hi/lo) + K group + MN chunk + row in group
        const uint32_t dst_common = (c8 >> 2) * 1u;                    // 0 = hi, 1 = lo
        const uint32_t in_grp = (uint32_t)((warp >> 1) * 1), in_row = (uint32_t)(((warp & 1) * 4 + rq) * 16);
        const uint32_t dst_a = dst_common * kWALoOff + in_grp * kWA_LBO + (c8 & 3) * kW_SBO + in_row;      // + mb * 4 * kW_SBO
        const uint32_t dst_b = kWABytes + dst_common * kWBLoOff + in_grp * kWB_LBO + (c8 & 3) * kW_SBO + in_row;
        // the 4 blocks of this M tile: tap and 32-channel block inside the gathered row
        int tap[4], blk[4];
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int m = m0 + 32 * mb;
            tap[mb] = m < m_total ? m / c_in : -1;
            blk[mb] = m < m_total ? (m - tap[mb] * c_in) >> 5 : 0;
        }
        const long long row_bytes = (long long)blocks_in * 128;
        const int nb_live = (c_out - o0 + 31) / 32;                    // 32-output blocks of dz inside this N tile (1 or 2)
        int idx_next[4];
        int fetched = 0;
        auto fetch_idx = [&]() {
            const int v = fetched * TK + kk;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                int r = -1;
                if (v < span && tap[mb] >= 0)
                    r = nbr != nullptr ? load_idx<I64>(nbr, (long long)tap[mb] * n_out_rows + v_lo + v) : (int)(v_lo + v);
                idx_next[mb] = r;
            }
            ++fetched;
        };
        fetch_idx();
        int stage_w = 0, stage_r = 0;
        uint32_t phase_w = 0;
        for (int kb = 0; kb < n_kb + kWLookahead; ++kb) {
            if (kb >= kWLookahead) {
                cp_async_wait<kWLookahead - 1>();
                fence_proxy_async();
                __syncwarp();
                if (lane0) mbar_arrive_a(full_a + 8 * stage_r);
                if (++stage_r == kWStages) stage_r = 0;
            }
            if (kb < n_kb) {
                int rr[4];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) rr[mb] = idx_next[mb] < n_in_rows ? idx_next[mb] : -1;
                fetch_idx();
                if (lane0) mbar_wait_a(empty_a + 8 * stage_w, phase_w ^ 1);
                __syncwarp();
                const uint32_t dst = smem_base + stage_w * kWStageBytes;
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) {
                    const bool ok = rr[mb] >= 0;
                    const uint8_t* src = in16 + (long long)(ok ? rr[mb] : 0) * row_bytes + blk[mb] * 128 + c8 * 16;
                    cp_async16(dst + dst_a + mb * 4 * kW_SBO, src, ok ? 16u : 0u);
                }
                const int v = kb * TK + kk;
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    const bool ok = v < span && nb < nb_live;
                    const uint8_t* src = dz16 + (ok ? (v_lo + v) : 0) * ((long long)blocks_dz * 128) + ((o0 >> 5) + nb) * 128 + c8 * 16;
                    cp_async16(dst + dst_b + nb * 4 * kW_SBO, ok ? src : dz16, ok ? 16u : 0u);
                }
                if (++stage_w == kWStages) { stage_w = 0; phase_w ^= 1; }
            }
            cp_async_commit();
        }
    } else if (warp == kProducerWarps) {
        if (lane == 0) {
            int last_g = -1, stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int g = (int)((long long)kb * WG_MAIN / n_kb);
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait_a(full_a + 8 * stage, phase);
                fence_after();
                const uint32_t a_hi = smem_base + stage * kWStageBytes;
                const uint32_t a_lo = a_hi + kWALoOff, b_hi = a_hi + kWABytes, b_lo = b_hi + kWBLoOff;
#pragma unroll
                for (int j = 0; j < TK / 16; ++j) {
                    const uint64_t dah = smem_desc(a_hi + j * 2 * kWA_LBO, kWA_LBO, kW_SBO);
                    const uint64_t dal = smem_desc(a_lo + j * 2 * kWA_LBO, kWA_LBO, kW_SBO);
                    const uint64_t dbh = smem_desc(b_hi + j * 2 * kWB_LBO, kWB_LBO, kW_SBO);
                    const uint64_t dbl = smem_desc(b_lo + j * 2 * kWB_LBO, kWB_LBO, kW_SBO);
                    umma_f16(tmem_d, dal, dbh, kIdescMN, (kb | j) != 0);
                    umma_f16(tmem_d, dah, dbl, kIdescMN, 1);
                    umma_f16(tmem_main, dah, dbh, kIdescMN, g == last_g);
                    last_g = g;
                }
                umma_commit_a(empty_a + 8 * stage);
                if (++stage == kWStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_a(accum_a);
        }
    }

    if (warp < 4 && n_kb > 0) {
        float s_in, inv_in, s_dz, inv_dz;
        scale_from_amax(__ldg(in_amax), s_in, inv_in);
        scale_from_amax(__ldg(dz_amax), s_dz, inv_dz);
        mbar_wait_a(accum_a, 0);
        fence_after();
        const int m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const float s_ab = s_in * s_dz;
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            float sum[16];
            uint32_t v[16];
            tmem_ld16(taddr + cb, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]) * kLoInv;
            for (int g = 1; g <= WG_MAIN; ++g) {
                const int first_kb = ((g - 1) * n_kb + WG_MAIN - 1) / WG_MAIN;
                if (first_kb >= n_kb || (int)((long long)first_kb * WG_MAIN / n_kb) != g - 1) continue;
                tmem_ld16(taddr + g * TN + cb, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (m < m_total) {
                float* p = dw + (long long)m * c_out + o0 + cb;
                if (o0 + cb + 15 < c_out && (c_out & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        red_add_f32x4(p + j, make_float4(sum[j] * s_ab, sum[j + 1] * s_ab, sum[j + 2] * s_ab, sum[j + 3] * s_ab));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (o0 + cb + j < c_out) atomicAdd(p + j, sum[j] * s_ab);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == kProducerWarps) {
        fence_after();
        tmem_dealloc(tmem_d, tmem_cols);
    }
}

void set_attrs() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(gather_gemm_p16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gather_gemm_p16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(wgrad_p16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes);
    cudaFuncSetAttribute(wgrad_p16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes);
    done = true;
}

}  // namespace

extern "C" {

int64_t hpl_split16_bytes(int64_t n_rows, int64_t channels) { return n_rows * ((channels + 31) / 32) * 128; }

int hpl_split16(const float* x, int64_t ld, int64_t n_rows, int64_t channels, const uint32_t* amax, void* x16, void* stream) {
    HPL_CHECK_ARG(x && amax && x16 && channels > 0 && ld % 4 == 0 && ld >= channels);
    HPL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)x16 & 127) == 0);
    if (n_rows == 0) return 0;
    const int blocks = (int)((channels + 31) / 32);
    const long long work = n_rows * blocks * 8;
    split16_kernel<<<(unsigned)((work + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld, n_rows, (int)channels, blocks, amax,
                                                                               reinterpret_cast<uint8_t*>(x16));
    HPL_RETURN_LAST();
}

int hpl_blur_gemm_p16(const void* in16, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size, int64_t n_out_rows,
                      int64_t c_in, int64_t c_out, const float* w, const float* bias, int act, float* out, int64_t ld_out,
                      int out_channel_major, void* workspace, const uint32_t* in_amax, void* stream) {
    HPL_CHECK_ARG(in16 && w && out && workspace && in_amax && c_in > 0 && c_out > 0 && filter_size > 0);
    HPL_CHECK_ARG(((uintptr_t)in16 & 127) == 0 && ((uintptr_t)workspace & 15) == 0 && ((uintptr_t)w & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    HPL_CHECK_ARG(out_channel_major ? ld_out >= n_out_rows : ld_out >= c_out);
    if (n_out_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    set_attrs();
    const int kb_per_tap = (int)((c_in + TK - 1) / TK);
    const long long n_tiles = (c_out + TN - 1) / TN;
    const long long image_bytes = n_tiles * filter_size * kb_per_tap * 2 * kBHalf;
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* w_amax = reinterpret_cast<uint32_t*>(image + image_bytes);
    const int rc = hpl_absmax(w, filter_size * c_in * c_out, w_amax, stream);
    if (rc != 0) return rc;
    const long long chunks = n_tiles * filter_size * kb_per_tap * (TN * (TK / 8));
    weight_image16p_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(w, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);
    const long long steps = (long long)filter_size * kb_per_tap * (TK / 16);
    const int n_main = steps <= 160 ? 1 : (steps <= 480 ? 3 : 7);
    dim3 grid((unsigned)((n_out_rows + TM - 1) / TM), (unsigned)n_tiles);
    const uint8_t* in8 = reinterpret_cast<const uint8_t*>(in16);
    if (idx64)
        gather_gemm_p16_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(in8, kb_per_tap, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_out, image,
                                                                        bias, act, out, ld_out, out_channel_major, n_main, in_amax, w_amax);
    else
        gather_gemm_p16_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(in8, kb_per_tap, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_out, image,
                                                                         bias, act, out, ld_out, out_channel_major, n_main, in_amax, w_amax);
    HPL_RETURN_LAST();
}

int hpl_blur_wgrad_p16(const void* in16, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size, int64_t n_out_rows,
                       int64_t c_in, int64_t c_out, const void* dz16, float* dw, const uint32_t* in_amax, const uint32_t* dz_amax,
                       void* stream) {
    HPL_CHECK_ARG(in16 && dz16 && dw && in_amax && dz_amax && c_in > 0 && c_out > 0 && filter_size > 0 && c_in % 32 == 0);
    HPL_CHECK_ARG(((uintptr_t)in16 & 127) == 0 && ((uintptr_t)dz16 & 127) == 0 && ((uintptr_t)dw & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    if (n_out_rows == 0) return 0;
    set_attrs();
    const long long m_tiles = (filter_size * c_in + TM - 1) / TM, n_tiles = (c_out + TN - 1) / TN;
    const long long base = m_tiles * n_tiles;
    long long splits = (4LL * num_sms() + base - 1) / base;
    const long long max_splits = (n_out_rows + 8 * TK - 1) / (8 * TK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rows_per_split = (n_out_rows + splits - 1) / splits;
    rows_per_split = (rows_per_split + TK - 1) / TK * TK;
    const long long max_rows = 160LL * WG_MAIN * 16;
    if (rows_per_split > max_rows) rows_per_split = max_rows;
    splits = (n_out_rows + rows_per_split - 1) / rows_per_split;
    HPL_CHECK_ARG(m_tiles <= 65535 && n_tiles <= 65535);
    dim3 grid((unsigned)splits, (unsigned)m_tiles, (unsigned)n_tiles);
    cudaStream_t s = as_stream(stream);
    const uint8_t* in8 = reinterpret_cast<const uint8_t*>(in16);
    const uint8_t* dz8 = reinterpret_cast<const uint8_t*>(dz16);
    const int blocks_in = (int)(c_in / 32), blocks_dz = (int)((c_out + 31) / 32);
    if (idx64)
        wgrad_p16_kernel<true><<<grid, kThreads, kWSmemBytes, s>>>(in8, blocks_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, dz8,
                                                                  blocks_dz, dw, rows_per_split, in_amax, dz_amax);
    else
        wgrad_p16_kernel<false><<<grid, kThreads, kWSmemBytes, s>>>(in8, blocks_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, dz8,
                                                                   blocks_dz, dw, rows_per_split, in_amax, dz_amax);
    HPL_RETURN_LAST();
}

}  // extern "C"
