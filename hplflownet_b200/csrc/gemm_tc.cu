// Gather-GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) with 3xTF32 error
// compensation -- the fast path of the learned lattice convolution (bilateralNN.py:198-221):
//     out[v, :] = act(bias + sum_f in[nbr[f, v], :] . W_f)
//
// Why 3xTF32: the contraction must match the reference's fp32 result to 1e-5 relative with K up to
// 8700; a single TF32 (10-bit mantissa) or BF16 pass cannot.  Each fp32 operand x is split as
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi) (x - hi is exact), and
//     a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi         (fp32 accumulation in TMEM)
// drops only the lo.lo term (~2^-22 relative).  Three tcgen05.mma per K step instead of one.
// The tensor core truncates when it aligns addends to a large accumulator (measured: the error of a
// single accumulator grows linearly with the number of accumulate steps, 3.6e-5 at K = 8700), so
// the two small cross terms go to their own TMEM accumulator and the hi.hi term is spread over
// 1, 3 or 7 accumulators by K range; the epilogue adds them in fp32 registers.
//
// Structure (one CTA = one 128-vertex x 64-output tile, 10 warps, 2 CTAs / SM):
//   warps 0-7  producers: gather the rows named by the neighbour table straight from HBM/L2
//              (LDG.128, 8 rows x 64 B per warp instruction), split hi/lo, and store both into
//              shared memory in the UMMA canonical K-major no-swizzle layout (a quarter warp writes
//              one 128-byte core matrix: conflict-free); 4-deep register prefetch.
//   warp 8     lane 0 issues tcgen05.mma (M=128, N=64, K=8, kind::tf32), accumulator in TMEM,
//              tcgen05.commit releases the stage / signals the epilogue through mbarriers.
//   warp 9     lane 0 streams the pre-split weight image of each K block with cp.async.bulk (TMA
//              bulk copy) completing on the stage's mbarrier.
//   warps 0-3  epilogue: tcgen05.ld the accumulator (one row per thread), bias + (Leaky)ReLU, store
//              vertex-major (float4) or channel-major (coalesced along vertices).
// The F-times gathered copy of the reference (bilateralNN.py:215-217) never exists in HBM.
#include "common.cuh"

namespace {

constexpr int TM = 128;            // vertices per CTA tile (UMMA M)
constexpr int TN = 64;             // output channels per CTA tile (UMMA N)
constexpr int TK = 16;             // K elements per pipeline stage (2 UMMA K-steps of 8)
constexpr int kStages = 4;
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 64;   // + MMA warp + weight-loader warp
constexpr int kPrefetch = 4;       // K blocks of gathered rows kept in flight per producer thread

constexpr int kAHalfBytes = TM * TK * 4;          // 8 KB: hi (or lo) part of the A stage
constexpr int kBHalfBytes = TN * TK * 4;          // 4 KB
constexpr int kStageBytes = 2 * kAHalfBytes + 2 * kBHalfBytes;   // 24 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024;          // + alignment slack

// canonical K-major, no swizzle: 16-byte chunk (row r, k-chunk kc) of a tile with R rows lives at
//   kc * (R * 16) + (r / 8) * 128 + (r % 8) * 16        -> SBO = 128 B, LBO = R * 16 B
constexpr uint32_t kA_LBO = TM * 16, kB_LBO = TN * 16, kSBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)      // suspend-time hint: sleep in hardware instead of re-polling
        : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {   // shared-window store (not a generic ST)
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);              // start address      bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // leading byte off.  bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride byte off.   bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    return d;                                           // base_offset 0, SWIZZLE_NONE
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 64
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
    hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
    lo.x = tf32_hi(x.x - hi.x); lo.y = tf32_hi(x.y - hi.y); lo.z = tf32_hi(x.z - hi.z); lo.w = tf32_hi(x.w - hi.w);
}

// ---------------------------------------------------------------------------------------------
// Weight image: w (F, C, Co) fp32 -> per (N tile, K block) an 8 KB block [hi 4 KB | lo 4 KB] already
// in the shared-memory layout, so one bulk copy per stage brings it in.
__global__ void weight_image_kernel(const float* __restrict__ w, int filter_size, int c_in, int c_out, int kb_per_tap,
                                    float* __restrict__ image) {
    // one thread per 16-byte chunk (n, kc) of one (tile, kblock)
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = TN * (TK / 4);                                   // 256 per K block
    const long long n_kb = (long long)filter_size * kb_per_tap;
    const long long n_tiles = (c_out + TN - 1) / TN;
    if (t >= n_tiles * n_kb * chunks) return;
    const int chunk = (int)(t % chunks);
    const long long blk = t / chunks;                                   // tile * n_kb + kb
    const long long kb = blk % n_kb, tile = blk / n_kb;
    const int f = (int)(kb / kb_per_tap), c0 = (int)(kb % kb_per_tap) * TK;
    const int n = chunk & (TN - 1), kc = chunk / TN;                    // kc in [0, 4)
    const int o = (int)tile * TN + n;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + 4 * kc + i;
        v[i] = (c < c_in && o < c_out) ? __ldg(w + ((long long)f * c_in + c) * c_out + o) : 0.f;
    }
    float4 hi, lo;
    split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
    const long long off = blk * (2 * kBHalfBytes / 4) + (kc * (TN * 16) + (n >> 3) * 128 + (n & 7) * 16) / 4;
    *reinterpret_cast<float4*>(image + off) = hi;
    *reinterpret_cast<float4*>(image + off + kBHalfBytes / 4) = lo;
}

// ---------------------------------------------------------------------------------------------
template <bool I64, bool SCALED>
__global__ void __launch_bounds__(kThreads, 2)
gather_gemm_tc_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows, const void* __restrict__ nbr,
                      int filter_size, long long n_out_rows, int c_in, int c_out, int kb_per_tap,
                      const float* __restrict__ w_image, const float* __restrict__ bias, int act,
                      float* __restrict__ out, long long ld_out, int out_cm, int n_main,
                      const float* __restrict__ row_scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n_tile = blockIdx.y;
    const int n_kb = filter_size * kb_per_tap;
    const uint32_t tmem_cols = (uint32_t)(TN * (n_main + 1));   // accumulator 0: cross terms; 1..n_main: hi.hi

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], kProducerWarps + 1);     // one elected lane per producer warp + the weight loader
            mbar_init(&empty_bar[s], 1);                     // one tcgen05.commit
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) {                            // MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ producers
        const int q = lane >> 3, r8 = lane & 7;              // 16-byte chunk of the 64-byte K block, row in group
        int row[2] = {-1, -1};
        float rscale[2] = {1.f, 1.f};                        // density normalisation folded into the gather
        float4 pre[kPrefetch][2];
        float pre_s[kPrefetch][2];                           // scale of each prefetched chunk (applied when consumed,
                                                             // so the loads stay in flight)

        auto load_rows = [&](int f) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long v = m0 + (warp * 2 + i) * 8 + r8;
                int r = -1;
                if (v < n_out_rows) {
                    r = nbr != nullptr ? load_idx<I64>(nbr, (long long)f * n_out_rows + v) : (int)v;
                    if (r >= n_in_rows) r = -1;
                }
                row[i] = r;
                if constexpr (SCALED) rscale[i] = r >= 0 ? __ldg(row_scale + r) : 1.f;
            }
        };
        auto issue = [&](int kb, float4* dst, float* sdst) {
            // rows of tap f must be current: callers walk kb in order, so refresh on tap change
            const int f = kb / kb_per_tap, c = (kb - f * kb_per_tap) * TK + 4 * q;
            if (kb % kb_per_tap == 0) load_rows(f);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                dst[i] = (row[i] >= 0 && c < c_in) ? __ldg(reinterpret_cast<const float4*>(in + (long long)row[i] * ld_in + c))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (SCALED) sdst[i] = rscale[i];
            }
        };

#pragma unroll
        for (int d = 0; d < kPrefetch; ++d)
            if (d < n_kb) issue(d, pre[d], pre_s[d]);

        for (int kb0 = 0; kb0 < n_kb; kb0 += kPrefetch) {
#pragma unroll
            for (int d = 0; d < kPrefetch; ++d) {
                const int kb = kb0 + d;
                if (kb >= n_kb) break;
                const int stage = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                float4 hi[2], lo[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float4 v = pre[d][i];
                    if constexpr (SCALED) {
                        const float sc = pre_s[d][i];
                        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                    }
                    split4(v, hi[i], lo[i]);
                }
                if (kb + kPrefetch < n_kb) issue(kb + kPrefetch, pre[d], pre_s[d]);   // refill this slot
                if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1);     // one poller per warp
                __syncwarp();
                const uint32_t a_hi = smem_base + stage * kStageBytes;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t off = q * (TM * 16) + (warp * 2 + i) * 128 + r8 * 16;
                    sts128(a_hi + off, hi[i]);
                    sts128(a_hi + kAHalfBytes + off, lo[i]);
                }
                fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);   // 32 same-address arrives would serialise in the LSU
            }
        }
    } else if (warp == kProducerWarps) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int last_g = -1;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int stage = kb % kStages;
                const int g = (int)((long long)kb * n_main / n_kb);
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait(&full_bar[stage], (kb / kStages) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + stage * kStageBytes);
                const uint32_t a_lo = a_hi + kAHalfBytes;
                const uint32_t b_hi = a_hi + 2 * kAHalfBytes;
                const uint32_t b_lo = b_hi + kBHalfBytes;
#pragma unroll
                for (int j = 0; j < TK / 8; ++j) {
                    const uint64_t dah = smem_desc(a_hi + j * 2 * kA_LBO, kA_LBO, kSBO);
                    const uint64_t dal = smem_desc(a_lo + j * 2 * kA_LBO, kA_LBO, kSBO);
                    const uint64_t dbh = smem_desc(b_hi + j * 2 * kB_LBO, kB_LBO, kSBO);
                    const uint64_t dbl = smem_desc(b_lo + j * 2 * kB_LBO, kB_LBO, kSBO);
                    umma_tf32(tmem_d, dal, dbh, (kb | j) != 0);      // cross terms -> accumulator 0
                    umma_tf32(tmem_d, dah, dbl, 1);
                    umma_tf32(tmem_main, dah, dbh, g == last_g);    // first step of a K range overwrites
                    last_g = g;
                }
                umma_commit(&empty_bar[stage]);                      // frees the stage when the MMAs retire
            }
            umma_commit(&accum_bar);
        }
    } else {
        // ------------------------------------------------------------------ weight loader
        if (lane == 0) {
            const float* src = w_image + (long long)n_tile * n_kb * (2 * kBHalfBytes / 4);
            for (int kb = 0; kb < n_kb; ++kb) {
                const int stage = kb % kStages;
                mbar_wait(&empty_bar[stage], ((kb / kStages) & 1) ^ 1);
                const uint32_t dst = smem_u32(smem + stage * kStageBytes + 2 * kAHalfBytes);
                mbar_arrive_expect_tx(&full_bar[stage], 2 * kBHalfBytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(src + (long long)kb * (2 * kBHalfBytes / 4)), "r"(2 * kBHalfBytes), "r"(smem_u32(&full_bar[stage]))
                             : "memory");
            }
        }
    }

    // ---------------------------------------------------------------------- epilogue (warps 0-3)
    if (warp < 4) {
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const long long m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const int o0 = n_tile * TN;
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            float sum[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = 0.f;
            for (int g = 0; g <= n_main; ++g) {                     // cross terms first, then the K ranges
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + g * TN + cb));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (m < n_out_rows) {
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int o = o0 + cb + j;
                    const float b = (bias != nullptr && o < c_out) ? __ldg(bias + o) : 0.f;
                    y[j] = apply_act(sum[j] + b, act);
                }
                if (!out_cm) {
                    float* p = out + m * ld_out + o0 + cb;
                    if (o0 + cb + 15 < c_out && (ld_out & 3) == 0 && ((uintptr_t)out & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(p + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
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
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:   dw[(f,c), o] += sum_v in[nbr[f,v], c] * dz[v, o]
// D (M x N) = A (M x K) . B (K x N) with M = (tap, channel) flattened, N = output channel, K = vertex.
// Both operands are contiguous along M / N in HBM (a gathered lattice row, a dz row), i.e. "MN-major".
// For 32-bit elements the tensor core accepts MN-major operands only in the SWIZZLE_128B_BASE32B
// layout (probed on B200 with tools/umma_probe.cu: every other layout type yields zeros).  An atom is
// 32 MN elements (128 B) x 4 K rows; element (m, k) of a tile lives at
//   (m / 32) * LBO + (k / 4) * SBO + (((k % 4) * 128 + (m % 32) * 4) ^ ((k % 4) << 5))
// i.e. the 16-byte chunk index inside a 128-byte row is XORed with 2 * (k % 4).  A quarter warp
// stores one full 128-byte row (a permutation of its 8 chunks): conflict-free, and the matching
// global read is 4 rows x 128 contiguous bytes per warp instruction.
// One CTA = one 128-row M tile x one 64-column N tile x a contiguous vertex range; the partial
// product leaves through fp32 RED.  Same 3xTF32 split and accumulator spreading as the forward.
constexpr int kPrefetchW = 3;                                   // register budget: 3 operands per stage here
constexpr int WG_MAIN = 3;                                     // hi.hi accumulators per CTA
constexpr uint32_t kMnAtom = 512;                              // 32 MN elements x 4 K rows
constexpr uint32_t kWA_SBO = (TM / 32) * kMnAtom;              // 2048: next 4 vertices of the A tile
constexpr uint32_t kWB_SBO = (TN / 32) * kMnAtom;              // 1024
constexpr uint32_t kInstrDescMN = kInstrDesc | (1u << 15) | (1u << 16);   // A and B MN-major

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t sbo_bytes) {
    return smem_desc(addr, kMnAtom, sbo_bytes) | ((uint64_t)1 << 61);    // layout type 1: SWIZZLE_128B_BASE32B
}

__device__ __forceinline__ void umma_tf32_mn(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(kInstrDescMN), "r"(accumulate)
        : "memory");
}

template <bool I64>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_tc_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows, const void* __restrict__ nbr,
                int filter_size, long long n_out_rows, int c_in, int c_out, const float* __restrict__ dz, long long ld_dz,
                float* __restrict__ dw, long long rows_per_split, const float* __restrict__ row_scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TM, o0 = blockIdx.z * TN;
    const int m_total = filter_size * c_in;
    const long long v_lo = rows_per_split * blockIdx.x;
    const long long v_hi = min(n_out_rows, v_lo + rows_per_split);
    const int n_kb = v_lo < v_hi ? (int)((v_hi - v_lo + TK - 1) / TK) : 0;
    const uint32_t tmem_cols = (uint32_t)(TN * (WG_MAIN + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], kProducerWarps);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < kProducerWarps) {
        // -------------------------------------------------------------- producers (A gathered, B = dz)
        const int kq = lane >> 3, c8 = lane & 7;       // K row inside the group of 4, 16-byte chunk of the 128-byte row
        const uint32_t swz = kq * 128 + ((c8 ^ (2 * kq)) * 16);
        // two A chunks per thread: (K group, MN atom) = warp task; tap / channel fixed for the whole kernel
        int tap[2], ch[2], kk_a[2];
        uint32_t off_a[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int wt = warp * 2 + i;
            const int kgrp = wt & 3, atom = wt >> 2;
            kk_a[i] = kgrp * 4 + kq;
            const int m = m0 + atom * 32 + c8 * 4;
            tap[i] = m < m_total ? m / c_in : -1;
            ch[i] = m < m_total ? m - tap[i] * c_in : 0;
            off_a[i] = kgrp * kWA_SBO + atom * kMnAtom + swz;
        }
        // one B chunk per thread
        const int kgrp_b = warp & 3, atom_b = warp >> 2;
        const int kk_b = kgrp_b * 4 + kq, n_b = atom_b * 32 + c8 * 4;
        const uint32_t off_b = kgrp_b * kWB_SBO + atom_b * kMnAtom + swz;
        const bool b_live = o0 + n_b < c_out;

        float4 pre[kPrefetchW][3];
        float pre_s[kPrefetchW][2];
        auto issue = [&](int kb, float4* dst, float* sdst) {
            const long long vb = v_lo + (long long)kb * TK;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long v = vb + kk_a[i];
                int r = -1;
                if (v < v_hi && tap[i] >= 0) {
                    r = nbr != nullptr ? load_idx<I64>(nbr, (long long)tap[i] * n_out_rows + v) : (int)v;
                    if (r >= n_in_rows) r = -1;
                }
                dst[i] = r >= 0 ? __ldg(reinterpret_cast<const float4*>(in + (long long)r * ld_in + ch[i]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                sdst[i] = (row_scale != nullptr && r >= 0) ? __ldg(row_scale + r) : 1.f;
            }
            const long long v = vb + kk_b;
            dst[2] = (v < v_hi && b_live) ? __ldg(reinterpret_cast<const float4*>(dz + v * ld_dz + o0 + n_b))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        };
#pragma unroll
        for (int d = 0; d < kPrefetchW; ++d)
            if (d < n_kb) issue(d, pre[d], pre_s[d]);

        for (int kb0 = 0; kb0 < n_kb; kb0 += kPrefetchW) {
#pragma unroll
            for (int d = 0; d < kPrefetchW; ++d) {
                const int kb = kb0 + d;
                if (kb >= n_kb) break;
                const int stage = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                float4 hi[3], lo[3];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float4 v = pre[d][i];
                    if (row_scale != nullptr) {
                        const float sc = pre_s[d][i];
                        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                    }
                    split4(v, hi[i], lo[i]);
                }
                split4(pre[d][2], hi[2], lo[2]);
                if (kb + kPrefetchW < n_kb) issue(kb + kPrefetchW, pre[d], pre_s[d]);
                if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1);
                __syncwarp();
                const uint32_t a_hi = smem_base + stage * kStageBytes;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    sts128(a_hi + off_a[i], hi[i]);
                    sts128(a_hi + kAHalfBytes + off_a[i], lo[i]);
                }
                sts128(a_hi + 2 * kAHalfBytes + off_b, hi[2]);
                sts128(a_hi + 2 * kAHalfBytes + kBHalfBytes + off_b, lo[2]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
            }
        }
    } else if (warp == kProducerWarps) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            int last_g = -1;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int stage = kb % kStages;
                const int g = (int)((long long)kb * WG_MAIN / n_kb);
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait(&full_bar[stage], (kb / kStages) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_base + stage * kStageBytes;
                const uint32_t a_lo = a_hi + kAHalfBytes;
                const uint32_t b_hi = a_hi + 2 * kAHalfBytes;
                const uint32_t b_lo = b_hi + kBHalfBytes;
#pragma unroll
                for (int j = 0; j < TK / 8; ++j) {
                    const uint64_t dah = smem_desc_mn(a_hi + j * 2 * kWA_SBO, kWA_SBO);
                    const uint64_t dal = smem_desc_mn(a_lo + j * 2 * kWA_SBO, kWA_SBO);
                    const uint64_t dbh = smem_desc_mn(b_hi + j * 2 * kWB_SBO, kWB_SBO);
                    const uint64_t dbl = smem_desc_mn(b_lo + j * 2 * kWB_SBO, kWB_SBO);
                    umma_tf32_mn(tmem_d, dal, dbh, (kb | j) != 0);
                    umma_tf32_mn(tmem_d, dah, dbl, 1);
                    umma_tf32_mn(tmem_main, dah, dbh, g == last_g);
                    last_g = g;
                }
                umma_commit(&empty_bar[stage]);
            }
            umma_commit(&accum_bar);
        }
    }

    // ------------------------------------------------------------------ epilogue (warps 0-3): RED into dw
    if (warp < 4 && n_kb > 0) {
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const int m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            float sum[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = 0.f;
            for (int g = 0; g <= WG_MAIN; ++g) {
                if (g > 0) {
                    // accumulator g-1 was written iff some kb maps to it: kb*WG_MAIN/n_kb == g-1
                    const int first_kb = ((g - 1) * n_kb + WG_MAIN - 1) / WG_MAIN;
                    if (first_kb >= n_kb || (int)((long long)first_kb * WG_MAIN / n_kb) != g - 1) continue;
                }
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + g * TN + cb));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (m < m_total) {
                float* p = dw + (long long)m * c_out + o0 + cb;
                if (o0 + cb + 15 < c_out && (c_out & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) red_add_f32x4(p + j, make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (o0 + cb + j < c_out) atomicAdd(p + j, sum[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

extern "C" {

int64_t hpl_blur_gemm_tc_workspace(int64_t filter_size, int64_t c_in, int64_t c_out) {
    const int64_t kb_per_tap = (c_in + TK - 1) / TK, n_tiles = (c_out + TN - 1) / TN;
    return n_tiles * filter_size * kb_per_tap * 2 * kBHalfBytes;
}

int hpl_blur_gemm_tc(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size,
                     int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* w, const float* bias, int act,
                     float* out, int64_t ld_out, int out_channel_major, float* workspace, const float* row_scale,
                     void* stream) {
    HPL_CHECK_ARG(in && w && out && workspace && c_in > 0 && c_out > 0 && filter_size > 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0 && ((uintptr_t)workspace & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    HPL_CHECK_ARG(out_channel_major ? ld_out >= n_out_rows : ld_out >= c_out);
    if (n_out_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int kb_per_tap = (int)((c_in + TK - 1) / TK);
    const long long n_tiles = (c_out + TN - 1) / TN;
    const long long chunks = n_tiles * filter_size * kb_per_tap * (TN * (TK / 4));
    weight_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(w, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, workspace);

    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gather_gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        cudaFuncSetAttribute(gather_gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        cudaFuncSetAttribute(gather_gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        cudaFuncSetAttribute(gather_gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr_set = true;
    }
    dim3 grid((unsigned)((n_out_rows + TM - 1) / TM), (unsigned)n_tiles);
    // accumulate steps of the hi.hi term per TMEM accumulator kept <= ~160 (truncation bias ~ steps * 2^-25)
    const long long steps = (long long)filter_size * kb_per_tap * (TK / 8);
    const int n_main = steps <= 160 ? 1 : (steps <= 480 ? 3 : 7);
#define HPL_LAUNCH_TC(I64, SC)                                                                                          \
    gather_gemm_tc_kernel<I64, SC><<<grid, kThreads, kSmemBytes, s>>>(in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, \
                                                                      (int)c_in, (int)c_out, kb_per_tap, workspace, bias, act, \
                                                                      out, ld_out, out_channel_major, n_main, row_scale)
    if (idx64) {
        if (row_scale) HPL_LAUNCH_TC(true, true); else HPL_LAUNCH_TC(true, false);
    } else {
        if (row_scale) HPL_LAUNCH_TC(false, true); else HPL_LAUNCH_TC(false, false);
    }
#undef HPL_LAUNCH_TC
    HPL_RETURN_LAST();
}

int hpl_blur_wgrad_tc(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size,
                      int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* dz, int64_t ld_dz, float* dw, float* db,
                      const float* row_scale, void* stream) {
    HPL_CHECK_ARG(in && dz && dw && c_in > 0 && c_out > 0 && filter_size > 0 && c_in % 4 == 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0);
    HPL_CHECK_ARG(ld_dz % 4 == 0 && ld_dz >= c_out && ((uintptr_t)dz & 15) == 0 && ((uintptr_t)dw & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    if (n_out_rows == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr_set = true;
    }
    const long long m_tiles = (filter_size * c_in + TM - 1) / TM, n_tiles = (c_out + TN - 1) / TN;
    const long long base = m_tiles * n_tiles;
    long long splits = (4LL * num_sms() + base - 1) / base;                 // ~2 waves at 2 CTAs / SM
    const long long max_splits = (n_out_rows + 8 * TK - 1) / (8 * TK);      // at least 8 stages per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rows_per_split = (n_out_rows + splits - 1) / splits;
    rows_per_split = (rows_per_split + TK - 1) / TK * TK;
    // keep the accumulate steps per hi.hi accumulator <= ~160 (see the forward kernel)
    const long long max_rows = 160LL * WG_MAIN * 8;
    if (rows_per_split > max_rows) rows_per_split = max_rows;
    splits = (n_out_rows + rows_per_split - 1) / rows_per_split;
    HPL_CHECK_ARG(m_tiles <= 65535 && n_tiles <= 65535);
    dim3 grid((unsigned)splits, (unsigned)m_tiles, (unsigned)n_tiles);
    cudaStream_t s = as_stream(stream);
    if (idx64)
        wgrad_tc_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in,
                                                                  (int)c_out, dz, ld_dz, dw, rows_per_split, row_scale);
    else
        wgrad_tc_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in,
                                                                   (int)c_out, dz, ld_dz, dw, rows_per_split, row_scale);
    if (db != nullptr) return hpl_column_sums(dz, ld_dz, n_out_rows, c_out, db, stream);
    HPL_RETURN_LAST();
}

}  // extern "C"
