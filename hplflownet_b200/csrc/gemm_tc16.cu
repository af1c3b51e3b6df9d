// Gather-GEMM and weight gradient on tcgen05 with a scaled FP16 hi/lo split ("3xFP16").
//
// Same contraction and the same fp32-level accuracy target as gemm_tc.cu (3xTF32), but every operand
// element travels as two 2-byte halves instead of two 4-byte TF32 words.  Both kernels are bound by
// the SM's shared-memory data pipe (operand staging of a *gathered* A, ncu: ~80 % busy, tensor pipe
// ~20 %), so halving the operand bytes -- and using kind::f16 MMAs with K = 16 -- is what moves them.
//
//   x / s = hi + lo * 2^-11 (+ <= 2^-22 |x / s|),   hi = fp16(x / s),   lo = fp16((x / s - hi) * 2^11)
//   s = 2^(floor(log2 max|x|) - 13): a per-tensor power of two that places the largest element in
//   [2^13, 2^14), so fp16's 5-bit exponent covers 2^-28 of the tensor's range at full relative precision
//   (smaller elements keep an absolute error <= 2^-39 of the maximum).  The maximum is a device scalar
//   produced by hpl_absmax -- no host synchronisation.
//   a.b = s_a s_b [ hi_a.hi_b + 2^-11 (hi_a.lo_b + lo_a.hi_b) ]   (lo.lo dropped: 2^-22 relative)
// fp16 x fp16 products are exact in the fp32 accumulator; the cross terms and the main term use separate
// TMEM accumulators and the main term is spread over 1/3/7 accumulators by K range (truncation bias, see
// gemm_tc.cu).  Layouts: forward operands K-major no-swizzle, weight-gradient operands MN-major no-swizzle
// (valid for 16-bit elements; address maps verified with tools/umma_probe.cu).
//
// Gather mapping (ncu-driven): a 128-bit LDG is serviced one quarter warp at a time, so the 8 lanes of a
// quarter warp must read ONE 128-byte line (one wavefront) -- lanes are (row = lane / 8, 16-byte chunk =
// lane % 8), i.e. 4 gathered rows x 128 contiguous bytes per warp instruction.  The first version had the 8
// lanes of a quarter warp on 8 different rows: 32 wavefronts per instruction and the L1 data pipe at ~80 %.
// Each lane then owns 4 consecutive K (or M) elements = half of a 16-byte operand chunk and stores hi / lo
// with 8-byte STS; the chunk strides (LBO / SBO, free parameters of the UMMA descriptor) are padded by 32
// bytes so that the 4 chunks a half warp touches fall into 4 different bank groups (conflict-free).
//
// Tile widths.  N = 64 (narrow layers, two CTAs of 8 producer warps per SM), N = 128 (Co >= 128), and N = 256 for
// Co >= 256 with >= 8192 rows: ONE CTA of 16 producer warps per SM, so the gathered operand is loaded and split a
// quarter as often per output.  At N = 256 tensor memory (512 columns) holds the cross accumulator plus a single main
// accumulator, which is exact for <= ~180 accumulate steps: the forward kernel splits longer K ranges over
// blockIdx.z (partial tiles summed in the zeroed output by fp32 RED, bias / activation / statistics by
// bias_act_kernel), the weight gradient limits its vertex range per CTA to 2560.
//
// Statistics.  The scale of every operand tensor comes from max|x|; the forward epilogue records max|out| for the
// next layer (RED.MAX per warp), see also rows.cu (act_backward_stats_kernel) -- no separate absmax passes between
// layers.  The weight operand is read strided (any dense permutation of (F, C, Co)), so the reference's conv-layout
// weights and their transposes for the data gradient are consumed in place.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TM = 128, TN = 64;
constexpr int TK = 32;                       // K elements per stage = 2 MMAs of K = 16
constexpr int kStages = 4;
constexpr int kProducerWarps = 8;
constexpr int kThreads = kProducerWarps * 32 + 64;
constexpr int kPrefetch = 2;                 // stages of gathered rows in flight per producer thread (weight gradient)
constexpr int kPrefetchF = 3;                // ... forward / data-gradient kernel (4 float4 per stage)
constexpr uint32_t kA_LBO = TM * 16 + 32;    // K-chunk stride of the A tile, padded: 2080
constexpr uint32_t kB_LBO = TN * 16, kSBO = 128;
constexpr int kAHalf = (TK / 8) * kA_LBO;    // 8320 B: hi (or lo) part of the A stage
constexpr int kBHalf = TN * TK * 2;          // 4 KB
constexpr int kStageBytes = 2 * kAHalf + 2 * kBHalf;
constexpr int kSmemBytes = kStages * kStageBytes + 1024;
constexpr uint32_t kIdescK = instr_desc(0, TM, TN, 0, 0);
constexpr uint32_t kIdescMN = instr_desc(0, TM, TN, 1, 1);
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ void scale_from_amax(uint32_t bits, float& scale, float& inv_scale) {
    int e = (int)((bits >> 23) & 0xff) - 127;
    if (bits == 0) e = 13;
    int se = e - 13;
    se = se < -100 ? -100 : (se > 100 ? 100 : se);
    scale = __uint_as_float((uint32_t)(se + 127) << 23);
    inv_scale = __uint_as_float((uint32_t)(127 - se) << 23);
}

// 4 consecutive fp32 -> 4 hi halves + 4 lo halves (2 x b32 each)
__device__ __forceinline__ void split4h(const float4 a, float inv_s, uint32_t* hi, uint32_t* lo) {
#ifdef HPL_EXP_NOSPLIT                      // timing experiment only (wrong numbers): one conversion, no residual arithmetic
    {
        const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
        hi[0] = lo[0] = *reinterpret_cast<const uint32_t*>(&h0);
        hi[1] = lo[1] = *reinterpret_cast<const uint32_t*>(&h1);
        return;
    }
#endif
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

// 8 consecutive fp32 -> 8 hi halves + 8 lo halves (4 x b32 each)
__device__ __forceinline__ void split8(const float4 a, const float4 b, float inv_s, uint32_t* hi, uint32_t* lo) {
    const float x[8] = {a.x * inv_s, a.y * inv_s, a.z * inv_s, a.w * inv_s, b.x * inv_s, b.y * inv_s, b.z * inv_s, b.w * inv_s};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((x[2 * i] - f.x) * kLoScale, (x[2 * i + 1] - f.y) * kLoScale);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
}

// ------------------------------------------------------------------------------------ absmax
__global__ void absmax_kernel(const float4* __restrict__ x, long long n4, uint32_t* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order as uints
}
__global__ void absmax_tail_kernel(const float* __restrict__ x, long long lo, long long n, uint32_t* __restrict__ out) {
    const long long i = lo + threadIdx.x;
    if (i < n) {
        const float m = fabsf(x[i]);
        if (m > 0.f) atomicMax(out, __float_as_uint(m));
    }
}

// ------------------------------------------------------------------------------ weight image
// w (F, C, Co) fp32 -> per (N tile, K block of 32) [hi 4 KB | lo 4 KB] in the shared-memory layout.
template <int TNv>
__global__ void weight_image16_kernel(const float* __restrict__ w, long long w_sf, long long w_sc, long long w_so,
                                      int filter_size, int c_in, int c_out, int kb_per_tap,
                                      const uint32_t* __restrict__ w_amax, uint8_t* __restrict__ image) {
    constexpr int TN = TNv;
    constexpr int kBHalf = TN * TK * 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int chunks = TN * (TK / 8);                                  // 16-byte chunks per K block
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
        v[i] = (c < c_in && o < c_out) ? __ldg(w + f * w_sf + c * w_sc + o * w_so) : 0.f;
    }
    uint32_t hi[4], lo[4];
    split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), inv_s, hi, lo);
    uint8_t* dst = image + blk * (2 * kBHalf) + kc * (TN * 16) + (n >> 3) * 128 + (n & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + kBHalf) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------ forward / dgrad
// PW producer warps (8: two CTAs per SM; 16: the 256-wide tile, one CTA per SM).  kb_per_split > 0: the K blocks are
// split over blockIdx.z, every CTA accumulates its range and adds its partial tile to `out` with fp32 RED (out zeroed
// by the launcher, bias / activation applied by bias_act_kernel afterwards).
// CL = 2 (256-wide tile only): the two CTAs of a thread-block cluster work on adjacent M tiles of the same (N tile, K
// range) and SHARE the weight stream -- each loads one half of a stage's weight tile and multicasts it into both CTAs
// (cp.async.bulk ... .multicast::cluster); a stage slot is free again when BOTH CTAs' MMAs on it have completed
// (tcgen05.commit ... .multicast::cluster arrives on the `empty` barrier of both).  The 580 -> 1024 layer is bound by
// L2 -> SM traffic (12.5 GB per launch, 8.7 GB of it weight tiles re-read by each of the 244 M tiles).
template <bool I64, int TNv, int kStagesV, int PW, int PF = 3, int CL = 1>
__global__ void __launch_bounds__(PW * 32 + 64, PW == 8 ? (PF == 1 ? 3 : 2) : 1)
gather_gemm_f16_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows, const void* __restrict__ nbr,
                       int filter_size, long long n_out_rows, int c_in, int c_out, int kb_per_tap,
                       const uint8_t* __restrict__ w_image, const float* __restrict__ bias, int act, float* __restrict__ out,
                       long long ld_out, int out_cm, int n_main, const uint32_t* __restrict__ in_amax,
                       const uint32_t* __restrict__ w_amax, uint32_t* __restrict__ out_amax, int kb_per_split) {
    // tile width (64 for narrow layers, 128 for Co >= 128, 256 for Co >= 256: fewer re-gathers of A) and ring depth
    constexpr int TN = TNv, kStages = kStagesV;
    constexpr int kProducerWarps = PW, kRowsPerWarp = TM / PW, NB = kRowsPerWarp / 4;
    constexpr int kPrefetchF = PF;                                       // (shadows the file-level default)
    constexpr bool kGroupSync = PW == 16;
    constexpr int kBHalf = TN * TK * 2;
    constexpr int kStageBytes = 2 * kAHalf + 2 * kBHalf;
    constexpr uint32_t kB_LBO = TN * 16;
    constexpr uint32_t kIdescK = instr_desc(0, TM, TN, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n_tile = blockIdx.y;
    const int n_kb_total = filter_size * kb_per_tap;
    const int kb_lo = kb_per_split > 0 ? (int)blockIdx.z * kb_per_split : 0;
    const int n_kb = kb_per_split > 0 ? min(kb_per_split, n_kb_total - kb_lo) : n_kb_total;      // K blocks of this CTA
    const bool partial = kb_per_split > 0;
    const uint32_t tmem_cols = (uint32_t)(TN * (n_main + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], (PW == 16 ? 1 : kProducerWarps) + 1);   // producer arrivals + the weight stream
            mbar_init(&empty_bar[s], CL);                                   // MMA completion of every CTA that receives the stage's weight tile
        }
        mbar_init(&accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == kProducerWarps) tmem_alloc(&tmem_slot, tmem_cols);
    fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();                                              // the peer's barriers exist before anything is multicast to them
    fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), accum_a = smem_u32(&accum_bar);
    const bool lane0 = lane == 0;
    float s_in, inv_in, s_w, inv_w;
    scale_from_amax(__ldg(in_amax), s_in, inv_in);
    scale_from_amax(__ldg(w_amax), s_w, inv_w);

    if (warp < kProducerWarps) {
        // lane = (row within a group of 4, 16-byte chunk of a 128-byte line): one line per quarter warp
        const int rq = lane >> 3, c16 = lane & 7;
        const float* rowp[NB];                                         // gathered rows of the current tap (+ chunk offset)
        int row_next[NB];                                              // indices of the next tap, loaded one tap early
        float4 pre[kPrefetchF][NB];
        uint32_t off[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int ml = warp * kRowsPerWarp + b * 4 + rq;             // row of the tile
            off[b] = (c16 >> 1) * kA_LBO + (ml >> 3) * 128 + (ml & 7) * 16 + (c16 & 1) * 8;
            rowp[b] = nullptr;
        }
        const long long v_first = m0 + warp * kRowsPerWarp + rq;
        auto fetch_rows = [&](int f) {                                 // issue the index loads of tap f (no use yet)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const long long v = v_first + b * 4;
                int r = -1;
                if (v < n_out_rows && f < filter_size) r = nbr != nullptr ? load_idx<I64>(nbr, (long long)f * n_out_rows + v) : (int)v;
                row_next[b] = r;
            }
        };
        int tap_i = kb_lo / kb_per_tap, kt_i = kb_lo % kb_per_tap, issued = 0;   // issue-side position: tap, K block inside the tap
        bool adopt = true;                                             // the first issue adopts its tap's rows even mid-tap
        fetch_rows(tap_i);
        auto issue = [&](float4* dst) {
            if (kt_i == 0 || adopt) {                                  // adopt this tap's rows, start fetching the next tap's
                adopt = false;
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    rowp[b] = (row_next[b] >= 0 && row_next[b] < n_in_rows) ? in + (long long)row_next[b] * ld_in + 4 * c16 : nullptr;
                fetch_rows(tap_i + 1);
            }
            const int c = kt_i * TK;
            const bool live = c + 4 * c16 < c_in;
#pragma unroll
            for (int b = 0; b < NB; ++b)
#ifdef HPL_EXP_NOLOAD                       // timing experiment only: no gathered loads
                dst[b] = make_float4((float)c, 1.f, 2.f, (float)b);
#else
                dst[b] = (rowp[b] != nullptr && live) ? __ldg(reinterpret_cast<const float4*>(rowp[b] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
            ++issued;
            if (++kt_i == kb_per_tap) { kt_i = 0; ++tap_i; }
        };
#pragma unroll
        for (int d = 0; d < kPrefetchF; ++d)
            if (d < n_kb) issue(pre[d]);

        int stage = 0;
        uint32_t phase = 0;
        for (int kb0 = 0; kb0 < n_kb; kb0 += kPrefetchF) {
#pragma unroll
            for (int d = 0; d < kPrefetchF; ++d) {
                if (kb0 + d >= n_kb) break;
                // Wide tile (16 producer warps, one CTA per SM): one mbarrier wait and one arrival per stage for ALL producer
                // warps, named barrier in between -- the SM serialises mbarrier operations (~40 cycles each; measured +4-5 %
                // on the 580 -> 1024 and 324 -> 256 layers).  With 8 warps and three CTAs per SM the lock step costs more than
                // it saves (cfg2: 0.211 -> 0.229 ms), so those keep per-warp waits / arrivals.
                // (waits are executed by whole warps: a single waiting lane leaves its warp diverged in front of the barrier /
                // the stores, which costs a slow reconvergence and dead-locked a sibling kernel in one build)
                if (kGroupSync) {
                    if (warp == 0) mbar_wait_a(empty_a + 8 * stage, phase ^ 1);
                    asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");
                } else {
                    mbar_wait_a(empty_a + 8 * stage, phase ^ 1);
                }
                const uint32_t a_hi = smem_base + stage * kStageBytes;
#pragma unroll
                for (int b = 0; b < NB; ++b) {                         // convert + store one chunk at a time (few live registers)
                    uint32_t hi[2], lo[2];
                    split4h(pre[d][b], inv_in, hi, lo);
                    sts64(a_hi + off[b], hi[0], hi[1]);
                    sts64(a_hi + kAHalf + off[b], lo[0], lo[1]);
                }
                if (issued < n_kb) issue(pre[d]);                      // refill the slot: 3 stages of loads stay in flight
                fence_proxy_async();
                if (kGroupSync) {
                    asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");
                    if (threadIdx.x == 0) mbar_arrive_a(full_a + 8 * stage);
                } else {
                    __syncwarp();
                    if (lane0) mbar_arrive_a(full_a + 8 * stage);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kProducerWarps) {
        // The issuer warp stays converged and one elected lane issues: under `if (lane == 0)` every tcgen05.mma / commit is
        // wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~80 cycles per instruction on the issuing thread).
        {
            // This warp paces the CTA: no division per K block (g = floor(kb * n_main / n_kb) is tracked
            // incrementally) and descriptors built once per stage as a constant high word + a running address.
            int last_g = -1, stage = 0, g = 0, g_num = 0;
            uint32_t phase = 0;
            constexpr uint64_t kDescA = (uint64_t)((kA_LBO >> 4) & 0x3fff) << 16 | (uint64_t)((kSBO >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
            constexpr uint64_t kDescB = (uint64_t)((kB_LBO >> 4) & 0x3fff) << 16 | (uint64_t)((kSBO >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
            for (int kb = 0; kb < n_kb; ++kb) {
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait_a(full_a + 8 * stage, phase);
                fence_after();
                // (descriptor addresses are offsets inside the CTA's own shared window: in a cluster launch the 32-bit
                // shared address of rank 1 carries the rank above bit 18)
                const uint32_t a_hi = ((smem_base & 0x3ffffu) + stage * kStageBytes) >> 4;  // 16-byte units from here on
                const uint32_t a_lo = a_hi + (kAHalf >> 4), b_hi = a_hi + (2 * kAHalf >> 4), b_lo = b_hi + (kBHalf >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < TK / 16; ++j) {
                        const uint64_t dah = kDescA | (a_hi + j * (2 * kA_LBO >> 4)), dal = kDescA | (a_lo + j * (2 * kA_LBO >> 4));
                        const uint64_t dbh = kDescB | (b_hi + j * (2 * kB_LBO >> 4)), dbl = kDescB | (b_lo + j * (2 * kB_LBO >> 4));
                        umma_f16(tmem_d, dal, dbh, kIdescK, (kb | j) != 0);
                        umma_f16(tmem_d, dah, dbl, kIdescK, 1);
                        umma_f16(tmem_main, dah, dbh, kIdescK, j == 0 ? g == last_g : 1);
                    }
                    if (CL > 1) umma_commit_mc(empty_a + 8 * stage, (uint16_t)((1u << CL) - 1));
                    else umma_commit_a(empty_a + 8 * stage);
                }
                last_g = g;
                if (++stage == kStages) { stage = 0; phase ^= 1; }
                g_num += n_main;
                if (g_num >= n_kb) { g_num -= n_kb; ++g; }
            }
            if (elect_one()) umma_commit_a(accum_a);
        }
    } else {
        if (lane == 0) {
            const uint8_t* src = w_image + ((long long)n_tile * n_kb_total + kb_lo) * (2 * kBHalf);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait_a(empty_a + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx_a(full_a + 8 * stage, 2 * kBHalf);
                if (CL == 2) {                                          // this CTA's half (rank 0: W_hi, rank 1: W_lo) into both CTAs
                    const uint32_t r = cluster_ctarank();
                    bulk_load_mc(smem_base + stage * kStageBytes + 2 * kAHalf + r * kBHalf, src + r * kBHalf, kBHalf, full_a + 8 * stage, 3);
                } else {
                    bulk_load_a(smem_base + stage * kStageBytes + 2 * kAHalf, src, 2 * kBHalf, full_a + 8 * stage);
                }
                src += 2 * kBHalf;
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }

    if (warp < 4) {
        mbar_wait_a(accum_a, 0);
        fence_after();
        const long long m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const int o0 = n_tile * TN;
        const float s_ab = s_in * s_w;
        float y_max = 0.f;                                                  // max |output| of this thread's row
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            float sum[16];
            uint32_t v[16];
            tmem_ld16(taddr + cb, v);                                       // cross terms
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]) * kLoInv;
            for (int g = 1; g <= n_main; ++g) {
                tmem_ld16(taddr + g * TN + cb, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (m < n_out_rows) {
                // bias + activation + max|y| with chunk-uniform branches (per-element switch / bound tests cost more issue
                // slots than the tensor-memory reads)
                float y[16];
                const bool full_chunk = o0 + cb + 16 <= c_out;
                if (bias == nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = sum[j] * s_ab;
                } else if (full_chunk && ((uintptr_t)bias & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + o0 + cb + j));
                        y[j] = fmaf(sum[j], s_ab, b.x);
                        y[j + 1] = fmaf(sum[j + 1], s_ab, b.y);
                        y[j + 2] = fmaf(sum[j + 2], s_ab, b.z);
                        y[j + 3] = fmaf(sum[j + 3], s_ab, b.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = fmaf(sum[j], s_ab, o0 + cb + j < c_out ? __ldg(bias + o0 + cb + j) : 0.f);
                }
                if (act == HPL_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = y[j] > 0.f ? y[j] : 0.f;
                } else if (act == HPL_ACT_LEAKY) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = y[j] > 0.f ? y[j] : HPL_LEAKY_RATE * y[j];
                }
                if (full_chunk) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y_max = fmaxf(y_max, fabsf(y[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (o0 + cb + j < c_out) y_max = fmaxf(y_max, fabsf(y[j]));
                }
                if (partial) {                                              // K split: add this CTA's partial tile (bias / act later)
                    if (!out_cm) {
                        float* p = out + m * ld_out + o0 + cb;
                        if (o0 + cb + 15 < c_out && (ld_out & 3) == 0 && ((uintptr_t)out & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                red_add_f32x4(p + j, make_float4(sum[j] * s_ab, sum[j + 1] * s_ab, sum[j + 2] * s_ab, sum[j + 3] * s_ab));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (o0 + cb + j < c_out) atomicAdd(p + j, sum[j] * s_ab);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (o0 + cb + j < c_out) atomicAdd(out + (long long)(o0 + cb + j) * ld_out + m, sum[j] * s_ab);
                    }
                } else if (!out_cm) {
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
        if (out_amax != nullptr && !partial) {                              // fused hpl_absmax of the output (next layer's scale)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) y_max = fmaxf(y_max, __shfl_xor_sync(0xffffffffu, y_max, o));
            if (lane == 0 && y_max > 0.f) atomicMax(out_amax, __float_as_uint(y_max));
        }
    }
    fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();                                              // no peer arrival / copy may land after this CTA is gone
    if (warp == kProducerWarps) {
        fence_after();
        tmem_dealloc(tmem_d, tmem_cols);
    }
}

// ------------------------------------------------------------------------------ bias / activation after a K split
// out = act(out + bias) in place over (n_rows, channels) vertex-major (ld) or (channels, n_rows) channel-major; max|out|
// of the result -> amax (may be NULL).
__global__ void bias_act_kernel(float* __restrict__ out, long long ld, long long n_rows, int channels,
                                const float* __restrict__ bias, int act, int out_cm, uint32_t* __restrict__ amax) {
    float m = 0.f;
    const long long total = (long long)n_rows * channels;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long idx;
        int o;
        if (out_cm) { o = (int)(t / n_rows); idx = (long long)o * ld + (t - (long long)o * n_rows); }
        else { const long long v = t / channels; o = (int)(t - v * channels); idx = v * ld + o; }
        const float y = apply_act(out[idx] + (bias != nullptr ? __ldg(bias + o) : 0.f), act);
        out[idx] = y;
        m = fmaxf(m, fabsf(y));
    }
    if (amax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
    }
}

// ------------------------------------------------------------------------------ weight gradient
// dw[(f,c), o] += sum_v in[nbr[f,v], c] * dz[v, o];  both operands MN-major no-swizzle fp16:
//   element (m, k) at (k / 8) * LBO + (m / 8) * 128 + (k % 8) * 16 + (m % 8) * 2
constexpr int kWStages = 3;
constexpr uint32_t kW_SBO = 128 + 32;                       // MN-chunk stride, padded (bank spreading)
constexpr uint32_t kWA_LBO = (TM / 8) * kW_SBO;             // 2560: next 8 vertices of the A tile
constexpr int kWAHalf = (TK / 8) * kWA_LBO;                 // 10240
template <int TNv> struct WCfg {
    static constexpr uint32_t kB_LBO = (TNv / 8) * kW_SBO;  // 1280 (N = 64) / 5120 (N = 256)
    static constexpr int kBHalf = (TK / 8) * kB_LBO;        // 5120 / 20480
    static constexpr int kStageBytes = 2 * kWAHalf + 2 * kBHalf;
    static constexpr int kSmemBytes = kWStages * kStageBytes + 1024;
};

// TNv = 64, PW = 8 producer warps, MAINS = 3 (two CTAs per SM) -- or the wide tile for Co >= 256: TNv = 256, PW = 16,
// MAINS = 1 (one CTA per SM; the gathered operand is staged a quarter as often, vertex ranges of <= 2560 per CTA).
template <bool I64, int TNv, int PW, int MAINS>
__global__ void __launch_bounds__(PW * 32 + 64, PW == 8 ? 2 : 1)
wgrad_f16_kernel(const float* __restrict__ in, long long ld_in, long long n_in_rows, const void* __restrict__ nbr,
                 int filter_size, long long n_out_rows, int c_in, int c_out, const float* __restrict__ dz, long long ld_dz,
                 float* __restrict__ dw, long long rows_per_split, const uint32_t* __restrict__ in_amax,
                 const uint32_t* __restrict__ dz_amax) {
    constexpr int TN = TNv, kProducerWarps = PW, WG_MAIN = MAINS;
    using W = WCfg<TNv>;
    constexpr uint32_t kWB_LBO = W::kB_LBO;
    constexpr int kWBHalf = W::kBHalf, kWStageBytes = W::kStageBytes;
    constexpr uint32_t kIdescMN = instr_desc(0, TM, TN, 1, 1);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kWStages], empty_bar[kWStages], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    // blockIdx.x = M tile (fastest): the CTAs that share a vertex range -- and therefore its dz rows and most of its
    // gathered rows -- are launched together and hit L2 (ncu before: 350 MB of DRAM reads for 139 MB of operands)
    const int m0 = blockIdx.x * TM, o0 = blockIdx.z * TN;
    const int m_total = filter_size * c_in;
    const long long v_lo = rows_per_split * blockIdx.y;
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
    float s_in, inv_in, s_dz, inv_dz;
    scale_from_amax(__ldg(in_amax), s_in, inv_in);
    scale_from_amax(__ldg(dz_amax), s_dz, inv_dz);

    if (warp < kProducerWarps) {
        // lane = (vertex within a group of 4, 16-byte chunk of a 128-byte line): one line per quarter warp.
        // Warp w owns vertex quad w % 8 of the stage (vertices 4q..4q+3) and six line tasks per vertex:
        //   warps 0-7 : the 4 segments of 32 M rows of the gathered operand + output segments 0, 1 of dz
        //   warps 8-15: output segments 2..7 of dz (N = 256 only)
        const int rq = lane >> 3, c16 = lane & 7;
        const int wq = warp & 7;
        const bool a_group = warp < 8;
        const int bseg0 = a_group ? 0 : 2;                             // first dz segment of this warp
        const int kk = wq * 4 + rq;                                    // vertex inside the stage
        const uint32_t sm_row = (wq >> 1) * 1u, sm_in = ((wq & 1) * 4 + rq) * 16 + (c16 & 1) * 8;
        const uint32_t base_a = sm_row * kWA_LBO + (c16 >> 1) * kW_SBO + sm_in;     // + seg * 4 * kW_SBO
        const uint32_t base_b = sm_row * kWB_LBO + (c16 >> 1) * kW_SBO + sm_in + bseg0 * 4 * kW_SBO;
        int ch[4];                                                     // channel of this lane's chunk per segment, -1 = beyond M
        unsigned ipos[4];                                              // element offset into nbr + v_lo (32-bit; host checks the range)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int m = m0 + t * 32 + 4 * c16;
            const int tp = (a_group && m < m_total) ? m / c_in : -1;
            ch[t] = tp >= 0 ? m - tp * c_in : -1;
            ipos[t] = (unsigned)((nbr != nullptr && tp >= 0 ? (long long)tp * n_out_rows : 0) + kk);
        }
        unsigned zpos = (unsigned)(kk * ld_dz + o0 + 32 * bseg0 + 4 * c16);   // element offset into dz + v_lo * ld_dz; segment t adds 32
        const int o_first = o0 + 32 * bseg0 + 4 * c16;
        const float* dz_base = dz + v_lo * ld_dz;

        float4 pre[kPrefetch][6];                                      // a_group: 4 gathered + 2 dz lines; else 6 dz lines
        int idx_next[4];                                               // gathered-row indices of the next stage to issue
        const int span = (int)(v_hi - v_lo);                           // vertices of this CTA
        int fetched = 0, issued = 0;                                   // stages whose indices / data have been requested
        auto fetch_idx = [&]() {
            const int base = fetched * TK;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                int r = -1;
                if (base + kk < span && ch[t] >= 0)
                    r = nbr != nullptr ? load_idx<I64>(nbr, v_lo + ipos[t]) : (int)(v_lo + ipos[t]);
                idx_next[t] = r;
                ipos[t] += TK;
            }
            ++fetched;
        };
        if (a_group) fetch_idx();
        auto issue = [&](float4* dst) {
            const int base = issued * TK;
            const bool v_ok = base + kk < span;
            if (a_group) {
                int rr[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) rr[t] = idx_next[t] < n_in_rows ? idx_next[t] : -1;
                fetch_idx();
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    dst[t] = rr[t] >= 0 ? __ldg(reinterpret_cast<const float4*>(in + (long long)rr[t] * ld_in + ch[t]))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    dst[4 + t] = (v_ok && o_first + 32 * t < c_out) ? __ldg(reinterpret_cast<const float4*>(dz_base + zpos + 32 * t))
                                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int t = 0; t < 6; ++t)
                    dst[t] = (v_ok && o_first + 32 * t < c_out) ? __ldg(reinterpret_cast<const float4*>(dz_base + zpos + 32 * t))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            zpos += (unsigned)(TK * ld_dz);
            ++issued;
        };
#pragma unroll
        for (int d = 0; d < kPrefetch; ++d)
            if (d < n_kb) issue(pre[d]);

        int stage = 0;
        uint32_t phase = 0;
        for (int kb0 = 0; kb0 < n_kb; kb0 += kPrefetch) {
#pragma unroll
            for (int d = 0; d < kPrefetch; ++d) {
                if (kb0 + d >= n_kb) break;
                mbar_wait_a(empty_a + 8 * stage, phase ^ 1);                        // (whole warp: see the forward kernel)
                const uint32_t a_hi = smem_base + stage * kWStageBytes;
                const uint32_t b_hi = a_hi + 2 * kWAHalf;
                // convert and store one chunk at a time (keeps the live registers low: this kernel holds
                // 12 gathered float4 per thread), then refill the slot with the loads of a later stage
                if (a_group) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        uint32_t hi[2], lo[2];
                        split4h(pre[d][t], inv_in, hi, lo);
                        sts64(a_hi + base_a + t * 4 * kW_SBO, hi[0], hi[1]);
                        sts64(a_hi + kWAHalf + base_a + t * 4 * kW_SBO, lo[0], lo[1]);
                    }
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        uint32_t hi[2], lo[2];
                        split4h(pre[d][4 + t], inv_dz, hi, lo);
                        sts64(b_hi + base_b + t * 4 * kW_SBO, hi[0], hi[1]);
                        sts64(b_hi + kWBHalf + base_b + t * 4 * kW_SBO, lo[0], lo[1]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 6; ++t) {
                        uint32_t hi[2], lo[2];
                        split4h(pre[d][t], inv_dz, hi, lo);
                        sts64(b_hi + base_b + t * 4 * kW_SBO, hi[0], hi[1]);
                        sts64(b_hi + kWBHalf + base_b + t * 4 * kW_SBO, lo[0], lo[1]);
                    }
                }
                if (issued < n_kb) issue(pre[d]);
                fence_proxy_async();
                __syncwarp();
                if (lane0) mbar_arrive_a(full_a + 8 * stage);
                if (++stage == kWStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kProducerWarps) {
        {                                                               // converged warp, elected lane issues (see the forward kernel)
            int last_g = -1, stage = 0, g = 0, g_num = 0;               // g = floor(kb * WG_MAIN / n_kb), incrementally
            uint32_t phase = 0;
            constexpr uint64_t kDescA = (uint64_t)((kWA_LBO >> 4) & 0x3fff) << 16 | (uint64_t)((kW_SBO >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
            constexpr uint64_t kDescB = (uint64_t)((kWB_LBO >> 4) & 0x3fff) << 16 | (uint64_t)((kW_SBO >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
            for (int kb = 0; kb < n_kb; ++kb) {
                const uint32_t tmem_main = tmem_d + (uint32_t)(TN * (1 + g));
                mbar_wait_a(full_a + 8 * stage, phase);
                fence_after();
                const uint32_t a_hi = (smem_base + stage * kWStageBytes) >> 4;             // 16-byte units from here on
                const uint32_t a_lo = a_hi + (kWAHalf >> 4), b_hi = a_hi + (2 * kWAHalf >> 4), b_lo = b_hi + (kWBHalf >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < TK / 16; ++j) {                  // one MMA = 2 K groups of 8 vertices
                        const uint64_t dah = kDescA | (a_hi + j * (2 * kWA_LBO >> 4)), dal = kDescA | (a_lo + j * (2 * kWA_LBO >> 4));
                        const uint64_t dbh = kDescB | (b_hi + j * (2 * kWB_LBO >> 4)), dbl = kDescB | (b_lo + j * (2 * kWB_LBO >> 4));
                        umma_f16(tmem_d, dal, dbh, kIdescMN, (kb | j) != 0);
                        umma_f16(tmem_d, dah, dbl, kIdescMN, 1);
                        umma_f16(tmem_main, dah, dbh, kIdescMN, j == 0 ? g == last_g : 1);
                    }
                    umma_commit_a(empty_a + 8 * stage);
                }
                last_g = g;
                if (++stage == kWStages) { stage = 0; phase ^= 1; }
                g_num += WG_MAIN;
                if (g_num >= n_kb) { g_num -= n_kb; ++g; }
            }
            if (elect_one()) umma_commit_a(accum_a);
        }
    }

    if (warp < 4 && n_kb > 0) {
        mbar_wait_a(accum_a, 0);
        fence_after();
        const int m = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const float s_ab = s_in * s_dz;
#pragma unroll 1
        for (int cb = 0; cb < TN; cb += 16) {
            if (o0 + cb >= c_out) break;
            float sum[16];
            uint32_t v[16];
            tmem_ld16(taddr + cb, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]) * kLoInv;
            for (int g = 1; g <= WG_MAIN; ++g) {
                const int first_kb = ((g - 1) * n_kb + WG_MAIN - 1) / WG_MAIN;     // accumulator used iff some kb maps to it
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
    cudaFuncSetAttribute(gather_gemm_f16_kernel<true, 64, 3, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kStageBytes + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<false, 64, 3, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kStageBytes + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<true, 64, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<false, 64, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<true, 128, 3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (2 * kAHalf + 2 * 128 * TK * 2) + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<false, 128, 3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (2 * kAHalf + 2 * 128 * TK * 2) + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<true, 256, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 * kAHalf + 2 * 256 * TK * 2) + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<false, 256, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 * kAHalf + 2 * 256 * TK * 2) + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<true, 256, 4, 16, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 * kAHalf + 2 * 256 * TK * 2) + 1024);
    cudaFuncSetAttribute(gather_gemm_f16_kernel<false, 256, 4, 16, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 * kAHalf + 2 * 256 * TK * 2) + 1024);
    cudaFuncSetAttribute(wgrad_f16_kernel<true, 64, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(wgrad_f16_kernel<false, 64, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(wgrad_f16_kernel<true, 256, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<256>::kSmemBytes);
    cudaFuncSetAttribute(wgrad_f16_kernel<false, 256, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<256>::kSmemBytes);
    done = true;
}

}  // namespace

extern "C" {

int hpl_absmax(const float* x, int64_t count, uint32_t* out_bits, void* stream) {
    HPL_CHECK_ARG(out_bits && (x || count == 0) && ((uintptr_t)x & 15) == 0);
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(out_bits, 0, 4, s);
    if (e != cudaSuccess) return (int)e;
    if (count == 0) return 0;
    const long long n4 = count / 4;
    if (n4 > 0) {
        long long blocks = (n4 + 255) / 256;
        const long long cap = 8LL * num_sms();
        if (blocks > cap) blocks = cap;
        absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n4, out_bits);
    }
    if (count % 4) absmax_tail_kernel<<<1, 32, 0, s>>>(x, n4 * 4, count, out_bits);
    HPL_RETURN_LAST();
}

int64_t hpl_blur_gemm_f16_workspace(int64_t filter_size, int64_t c_in, int64_t c_out) {
    const int64_t kb_per_tap = (c_in + TK - 1) / TK, n_cols = (c_out + 255) / 256 * 256;       // covers all tile widths
    return n_cols * filter_size * kb_per_tap * 2 * (TK * 2) + 16;      // image + the weight absmax slot
}

int hpl_blur_gemm_f16(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size,
                      int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* w, const float* bias, int act, float* out,
                      int64_t ld_out, int out_channel_major, void* workspace, const uint32_t* in_amax, void* stream) {
    return hpl_blur_gemm_f16_amax(in, ld_in, n_in_rows, nbr, idx64, filter_size, n_out_rows, c_in, c_out, w, 0, 0, 0, bias, act, out,
                                  ld_out, out_channel_major, workspace, 0, in_amax, nullptr, stream);
}

int hpl_blur_gemm_f16_amax(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size,
                           int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so,
                           const float* bias, int act, float* out, int64_t ld_out, int out_channel_major, void* workspace,
                           int workspace_valid, const uint32_t* in_amax, uint32_t* out_amax, void* stream) {
    if (w_sf == 0 && w_sc == 0 && w_so == 0) { w_sf = c_in * c_out; w_sc = c_out; w_so = 1; }   // contiguous (F, C, Co)
    HPL_CHECK_ARG(in && w && out && workspace && in_amax && c_in > 0 && c_out > 0 && filter_size > 0 && c_in % 4 == 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0 && ((uintptr_t)workspace & 15) == 0);
    HPL_CHECK_ARG(((uintptr_t)w & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    HPL_CHECK_ARG(out_channel_major ? ld_out >= n_out_rows : ld_out >= c_out);
    if (n_out_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    set_attrs();
    const int kb_per_tap = (int)((c_in + TK - 1) / TK);
    // tile width: 64 for narrow layers; 128 for Co >= 128 (the gathered operand is staged half as often); 256 for
    // Co >= 256 with enough vertices to fill the GPU at one CTA per SM (staged a quarter as often; 16 producer warps)
    static int wide_knob = -1;                               // HPL_GEMM_TN256=0 disables the 256-wide tile
    if (wide_knob < 0) { const char* e = getenv("HPL_GEMM_TN256"); wide_knob = e ? atoi(e) : 1; }
    const long long n_kb_total = filter_size * kb_per_tap;
    const bool tn256 = wide_knob != 0 && c_out >= 256 && n_out_rows >= 64LL * TM;
    const bool wide = c_out >= 128;
    const int tn = tn256 ? 256 : (wide ? 128 : 64);
    const long long n_tiles = (c_out + tn - 1) / tn;
    const long long image_bytes = n_tiles * n_kb_total * 2 * (tn * TK * 2);
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* w_amax = reinterpret_cast<uint32_t*>(image + image_bytes);
    if (!workspace_valid) {                                  // (the caller may keep the image of an unchanged weight: same n_out_rows class)
        const long long w_count = filter_size * c_in * c_out;
        const int rc = hpl_absmax(w, w_count, w_amax, stream);
        if (rc != 0) return rc;
        const long long chunks = n_tiles * n_kb_total * (tn * (TK / 8));
        const unsigned img_blocks = (unsigned)((chunks + 255) / 256);
        if (tn256)
            weight_image16_kernel<256><<<img_blocks, 256, 0, s>>>(w, w_sf, w_sc, w_so, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);
        else if (wide)
            weight_image16_kernel<128><<<img_blocks, 256, 0, s>>>(w, w_sf, w_sc, w_so, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);
        else
            weight_image16_kernel<64><<<img_blocks, 256, 0, s>>>(w, w_sf, w_sc, w_so, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);
    }
    const long long steps = n_kb_total * (TK / 16);
    const long long m_tiles = (n_out_rows + TM - 1) / TM;
    if (tn256) {
        // TMEM (512 columns) holds ONE main accumulator at this width, good for <= ~180 accumulate steps: longer K ranges
        // are split over blockIdx.z, partial tiles are summed in `out` by fp32 RED, bias / activation / statistics follow
        // in bias_act_kernel.
        const int max_kb = 180 / (TK / 16);
        const long long splits = (n_kb_total + max_kb - 1) / max_kb;
        const int kb_per_split = splits > 1 ? (int)((n_kb_total + splits - 1) / splits) : 0;
        const long long z = splits > 1 ? (n_kb_total + kb_per_split - 1) / kb_per_split : 1;
        HPL_CHECK_ARG(z <= 65535);
        dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)z);
        constexpr int kSt = 4, kPW = 16;
        const int smem = kSt * (2 * kAHalf + 2 * 256 * TK * 2) + 1024;
        if (z > 1) {
            const long long bytes = 4LL * (out_channel_major ? c_out * ld_out : n_out_rows * ld_out);
            cudaError_t e = cudaMemsetAsync(out, 0, (size_t)bytes, s);
            if (e != cudaSuccess) return (int)e;
        }
        const float* k_bias = z > 1 ? nullptr : bias;
        const int k_act = z > 1 ? HPL_ACT_NONE : act;
        uint32_t* k_amax = z > 1 ? nullptr : out_amax;
        static int cluster_knob = -1;                           // HPL_WIDE_CLUSTER=1: pairs of M tiles share the weight stream (multicast)
        if (cluster_knob < 0) { const char* e = getenv("HPL_WIDE_CLUSTER"); cluster_knob = e ? atoi(e) : 0; }
        if (cluster_knob && m_tiles >= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)((m_tiles + 1) / 2 * 2), (unsigned)n_tiles, (unsigned)z);   // (an odd last M tile gets an idle partner)
            cfg.blockDim = dim3(kPW * 32 + 64);
            cfg.dynamicSmemBytes = (size_t)smem;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            const int fs = (int)filter_size, ci = (int)c_in, co = (int)c_out, one = 1;
            cudaError_t e;
            if (idx64)
                e = cudaLaunchKernelEx(&cfg, gather_gemm_f16_kernel<true, 256, kSt, kPW, 3, 2>, in, ld_in, n_in_rows, nbr, fs, n_out_rows, ci, co,
                                       kb_per_tap, (const uint8_t*)image, k_bias, k_act, out, ld_out, out_channel_major, one, in_amax,
                                       (const uint32_t*)w_amax, k_amax, kb_per_split);
            else
                e = cudaLaunchKernelEx(&cfg, gather_gemm_f16_kernel<false, 256, kSt, kPW, 3, 2>, in, ld_in, n_in_rows, nbr, fs, n_out_rows, ci, co,
                                       kb_per_tap, (const uint8_t*)image, k_bias, k_act, out, ld_out, out_channel_major, one, in_amax,
                                       (const uint32_t*)w_amax, k_amax, kb_per_split);
            if (e != cudaSuccess) return (int)e;
        } else if (idx64)
            gather_gemm_f16_kernel<true, 256, kSt, kPW><<<grid, kPW * 32 + 64, smem, s>>>(
                in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, kb_per_tap, image, k_bias, k_act, out,
                ld_out, out_channel_major, 1, in_amax, w_amax, k_amax, kb_per_split);
        else
            gather_gemm_f16_kernel<false, 256, kSt, kPW><<<grid, kPW * 32 + 64, smem, s>>>(
                in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, kb_per_tap, image, k_bias, k_act, out,
                ld_out, out_channel_major, 1, in_amax, w_amax, k_amax, kb_per_split);
        if (z > 1 && (bias != nullptr || act != HPL_ACT_NONE || out_amax != nullptr)) {
            const long long total = out_channel_major ? c_out * n_out_rows : n_out_rows * c_out;
            long long blocks = (total + 255) / 256;
            if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
            bias_act_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, ld_out, n_out_rows, (int)c_out, bias, act, out_channel_major, out_amax);
        }
        HPL_RETURN_LAST();
    }
    // accumulate steps per hi.hi accumulator <= ~160-190; TMEM holds (n_main + 1) * tn <= 512 columns
    const int n_main = steps <= 160 ? 1 : ((steps <= 480 || wide) ? 3 : 7);
    dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
#define HPL_LAUNCH_F16(I64, TNV, ST)                                                                                         \
    gather_gemm_f16_kernel<I64, TNV, ST, 8><<<grid, kThreads, ST * (2 * kAHalf + 2 * TNV * TK * 2) + 1024, s>>>(             \
        in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, kb_per_tap, image, bias, act, out,   \
        ld_out, out_channel_major, n_main, in_amax, w_amax, out_amax, 0)
    if (wide) {
        if (idx64) HPL_LAUNCH_F16(true, 128, 3); else HPL_LAUNCH_F16(false, 128, 3);
    } else if (m_tiles >= 2LL * num_sms() && n_main == 1) {    // (n_main == 1: three accumulator sets fit tensor memory)
        // Enough tiles to fill every SM three times over: THREE CTAs per SM (3 stages, prefetch depth 1, 64 registers).
        // Measured on cfg2 x 32 clouds: forward 0.228 -> 0.205 ms, data gradient 0.219 -> 0.198 ms -- a third independent
        // pipeline per SM hides the tile heads / tails (18 % of the producers' time at two CTAs per SM) better than a
        // deeper per-thread prefetch does.
        const int smem = 3 * kStageBytes + 1024;
        if (idx64)
            gather_gemm_f16_kernel<true, 64, 3, 8, 1><<<grid, kThreads, smem, s>>>(in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, kb_per_tap, image, bias, act, out, ld_out, out_channel_major, n_main, in_amax, w_amax, out_amax, 0);
        else
            gather_gemm_f16_kernel<false, 64, 3, 8, 1><<<grid, kThreads, smem, s>>>(in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, kb_per_tap, image, bias, act, out, ld_out, out_channel_major, n_main, in_amax, w_amax, out_amax, 0);
    } else {
        if (idx64) HPL_LAUNCH_F16(true, 64, 4); else HPL_LAUNCH_F16(false, 64, 4);
    }
#undef HPL_LAUNCH_F16
    HPL_RETURN_LAST();
}

int hpl_blur_wgrad_f16(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size,
                       int64_t n_out_rows, int64_t c_in, int64_t c_out, const float* dz, int64_t ld_dz, float* dw, float* db,
                       const uint32_t* in_amax, const uint32_t* dz_amax, void* stream) {
    HPL_CHECK_ARG(in && dz && dw && in_amax && dz_amax && c_in > 0 && c_out > 0 && filter_size > 0 && c_in % 4 == 0);
    HPL_CHECK_ARG(ld_in % 4 == 0 && ld_in >= c_in && ((uintptr_t)in & 15) == 0);
    HPL_CHECK_ARG(ld_dz % 4 == 0 && ld_dz >= c_out && ((uintptr_t)dz & 15) == 0 && ((uintptr_t)dw & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || filter_size == 1);
    if (n_out_rows == 0) return 0;
    set_attrs();
    static int wide_knob = -1;                               // HPL_GEMM_TN256=0 disables the 256-wide tile
    if (wide_knob < 0) { const char* e = getenv("HPL_GEMM_TN256"); wide_knob = e ? atoi(e) : 1; }
    const bool tn256 = wide_knob != 0 && c_out >= 256 && n_out_rows >= 64LL * TM;
    const int tn = tn256 ? 256 : TN, n_mains = tn256 ? 1 : 3;
    const long long m_tiles = (filter_size * c_in + TM - 1) / TM, n_tiles = (c_out + tn - 1) / tn;
    const long long base = m_tiles * n_tiles;
    long long splits = ((tn256 ? 2LL : 4LL) * num_sms() + base - 1) / base;
    const long long max_splits = (n_out_rows + 8 * TK - 1) / (8 * TK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rows_per_split = (n_out_rows + splits - 1) / splits;
    rows_per_split = (rows_per_split + TK - 1) / TK * TK;
    const long long max_rows = 160LL * n_mains * 16;                       // <= ~160 accumulate steps (K = 16) per accumulator
    if (rows_per_split > max_rows) rows_per_split = max_rows;
    splits = (n_out_rows + rows_per_split - 1) / rows_per_split;
    HPL_CHECK_ARG(splits <= 65535 && n_tiles <= 65535);
    HPL_CHECK_ARG(filter_size * n_out_rows < (1LL << 31) && (rows_per_split + TK) * ld_dz < (1LL << 31));   // 32-bit in-kernel offsets
    dim3 grid((unsigned)m_tiles, (unsigned)splits, (unsigned)n_tiles);
    cudaStream_t s = as_stream(stream);
#define HPL_LAUNCH_WG(I64, TNV, PWV, MN)                                                                                          \
    wgrad_f16_kernel<I64, TNV, PWV, MN><<<grid, PWV * 32 + 64, WCfg<TNV>::kSmemBytes, s>>>(                                       \
        in, ld_in, n_in_rows, nbr, (int)filter_size, n_out_rows, (int)c_in, (int)c_out, dz, ld_dz, dw, rows_per_split, in_amax, dz_amax)
    if (tn256) {
        if (idx64) HPL_LAUNCH_WG(true, 256, 16, 1); else HPL_LAUNCH_WG(false, 256, 16, 1);
    } else {
        if (idx64) HPL_LAUNCH_WG(true, 64, 8, 3); else HPL_LAUNCH_WG(false, 64, 8, 3);
    }
#undef HPL_LAUNCH_WG
    if (db != nullptr) return hpl_column_sums(dz, ld_dz, n_out_rows, c_out, db, stream);
    HPL_RETURN_LAST();
}

}  // extern "C"
