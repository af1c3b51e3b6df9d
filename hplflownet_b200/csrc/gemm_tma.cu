// Engine 4: "split once, copy many" 3xFP16 gather-GEMM (blur forward / data gradient) as a PERSISTENT kernel whose
// operands are staged by the copy engines -- TMA and cp.async -- instead of by converting producer warps.
//
// ncu on the register-staged kernels (gemm_tc16.cu) showed them paced by instruction issue: eight producer warps LDG
// the gathered rows, split every element into fp16 hi/lo once per tap (15x) and STS them into the UMMA layout; the
// tensor pipe idles at ~20 %.  Here the split happens ONCE per tensor (hpl_h16_split) and shared memory is filled by
//   * cp.async (LDGSTS, 16 bytes, zero-fill) -- a quarter warp moves one 128-byte line (one plane of one gathered row)
//     straight into its SWIZZLE_128B position; completion is reported asynchronously to the stage's mbarrier
//     (cp.async.mbarrier.arrive.noinc), so producers never wait for their own copies;
//   * the Tensor Memory Accelerator: one cp.async.bulk per weight block, 128 x 64 tile loads for the un-gathered 1x1
//     layers, and -- optionally, HPL_TMA_WARPS -- its row-gather mode for the gathered rows,
//       cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4      (UTMALDG.2D.GATHER4 in SASS)
//     which takes FOUR row indices per instruction and writes the four 128-byte rows in the SWIZZLE_128B pattern, i.e.
//     exactly the K-major operand layout tcgen05.mma consumes.  A missing neighbour (-1) is redirected to an all-zero
//     row appended to the image.
//
// CTA = persistent worker, one per SM (~225 KB of shared memory), work item = (128-vertex tile, N tile):
//   warps 0-7   A producers : stage ring of 5 (N = 64) / 4 (N = 128) x 32 KB (hi plane | lo plane, 64 channels);
//                             the tile's whole index block (F x 128 entries) sits in shared memory one item ahead
//   warp 9      B producer  : pre-split, pre-swizzled weight image, one cp.async.bulk per (tap, 64-channel block)
//   warp 14     fence relay : cp.async writes through the generic proxy and completes asynchronously, tcgen05.mma reads
//                             through the async proxy: this warp waits for a full stage, executes fence.proxy.async and
//                             hands the stage to the issuer (keeps the MEMBAR out of the pacing thread)
//   warp 8      MMA issuer  : lane 0, tcgen05.mma.cta_group::1.kind::f16 M128 x N{64,128} x K16, three MMAs per K step
//                             (lo.hi + hi.lo into a cross accumulator, hi.hi into 1/3/7 main accumulators)
//   warps 10-13 epilogue    : tcgen05.ld, scales, bias, activation, stores -- overlapped with the next work item
//                             through double-buffered TMEM accumulators when 512 columns allow
//
// Measured on B200 (cfg2 x 32 clouds, H = 242 429; 64 -> 64 channels; tools/try_tma.py):
//   * TMA gather4 alone: 0.31 ms.  The TMA unit retires about one gather4 (512 bytes) per ~48 cycles per SM whatever the
//     number of issuing warps: ~3 TB/s over the chip, below the ~5.8 TB/s the LSU paths reach on the same rows.
//   * cp.async alone: 0.25 ms -- on par with the register-staged engine 2 (0.25 ms in the same harness), hybrids no
//     better.  Ablations (loads off / MMAs off / both): every variant of the L2 -> SM gather tops out near 5.8-5.9 TB/s
//     (ncu: L2 13 % busy, L1 19 %, tensor pipe 14 %, nothing saturated).  It is not a byte limit either: spatially
//     sorted vertices (overlapping gathers) change nothing for engine 2 and slow this engine down (tools/try_sorted.py).
//     Engine 2 is bound by its producers' instruction stream, this engine by its ONE MMA-issuing thread and ONE ring per
//     SM (~2000 cycles per stage); the next step is two issue pipelines per SM, see DESIGN.md 3.2b.  Pitfalls found on
//     the way: a per-tap dependent index load cost 150 us (hence the shared-memory index block), a 64-bit division per K
//     step in the issuer thread another 40 us.
//   * Since engine 4 needs two extra split passes (21 us each) per layer and direction, engine 2 stays the default.
//
// h16 image of a vertex-major fp32 matrix x (n_rows, C):  (n_rows + 1) rows of [ hi(ld16) | lo(ld16) ] halves,
// ld16 = round8(C), x / s = hi + lo * 2^-11 (s: per-tensor power of two from hpl_absmax, as in gemm_tc16.cu);
// row n_rows is zero.  Two tensor maps (hi plane, lo plane) view it as (n_rows + 1) x C with a 4 * ld16-byte row
// pitch; channels beyond C are zero-filled by the TMA unit (out-of-bounds box columns) or by cp.async (src-size 0).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int TM = 128;
constexpr int TK = 64;                          // channels per stage = one 128-byte swizzle row of halves
constexpr int kAPlane = TM * TK * 2;            // 16 KB: hi (or lo) plane of an A stage
constexpr int kAStage = 2 * kAPlane;            // 32 KB
constexpr int kProducerWarps = 8, kMmaWarp = 8, kBWarp = 9;      // warps 10..13: epilogue, warp 14: proxy-fence relay
constexpr int kRelayWarp = 14;
constexpr int kThreads = 15 * 32;
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
constexpr int kMaxTaps = 16;                    // filter taps the shared-memory index block holds (F = 15 in HPLFlowNet)
constexpr int kIdxBlock = kMaxTaps * TM;        // ints per index block (double-buffered)
constexpr int kIdxPerThread = kIdxBlock / (kProducerWarps * 32);

template <int TN> struct Cfg {
    static constexpr int kBPlane = TN * TK * 2;                     // 8 / 16 KB
    static constexpr int kBStage = 2 * kBPlane;
    static constexpr int kStagesA = TN == 64 ? 5 : 4;
    static constexpr int kStagesB = TN == 64 ? 3 : 2;
    static constexpr int kSmem = kStagesA * kAStage + kStagesB * kBStage + 2 * kIdxBlock * 4 + 1024;
};

__device__ __forceinline__ void scale_from_amax(uint32_t bits, float& scale, float& inv_scale) {
    int e = (int)((bits >> 23) & 0xff) - 127;
    if (bits == 0) e = 13;
    int se = e - 13;
    se = se < -100 ? -100 : (se > 100 ? 100 : se);
    scale = __uint_as_float((uint32_t)(se + 127) << 23);
    inv_scale = __uint_as_float((uint32_t)(127 - se) << 23);
}

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - f.x) * kLoScale, (b - f.y) * kLoScale);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// bounded spin: a barrier that never flips (descriptor / byte-count bug) traps instead of hanging the GPU
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spins > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_tile2d(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar)
        : "memory");
}

// 16-byte LDGSTS; src_bytes = 0 writes zeros (missing rows / channels beyond the row)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed (the pending count
// is not incremented: the barrier is initialised with one expected arrival per copying thread)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major SWIZZLE_128B shared-memory descriptor: 8-row x 128-byte atoms, 1024 bytes apart (SBO); LBO unused
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ------------------------------------------------------------------------------------ h16 split
// one thread = 8 consecutive channels of one row (two float4 in, two uint4 out); rows >= n_rows are zero
__global__ void h16_split_kernel(const float* __restrict__ x, long long ld, long long n_rows, int channels, int ld16,
                                 const uint32_t* __restrict__ amax, uint4* __restrict__ out) {
    const int cpr = ld16 / 8;                                              // 16-byte chunks per plane row
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_rows + 1) * cpr) return;
    const long long row = t / cpr;
    const int c0 = (int)(t % cpr) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row < n_rows) {
        const float* p = x + row * ld + c0;
        if (c0 + 8 <= channels) {                                          // ld % 4 == 0 and c0 % 8 == 0: 16-byte aligned
            const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + i < channels) v[i] = __ldg(p + i);
        }
    }
    float s, inv_s;
    scale_from_amax(__ldg(amax), s, inv_s);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i] * inv_s, v[2 * i + 1] * inv_s, hi[i], lo[i]);
    uint4* dst = out + row * (2 * cpr) + (c0 >> 3);
    dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[cpr] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------ weight image
// w (F, C, Co) fp32 -> per (N tile, tap, 64-channel block): [hi TN x 128 B | lo TN x 128 B], rows = output
// channel n, SWIZZLE_128B: 16-byte chunk kc of row n sits at (n / 8) * 1024 + (n % 8) * 128 + ((kc ^ (n % 8)) * 16)
template <int TN>
__global__ void weight_image_sw128_kernel(const float* __restrict__ w, int filter_size, int c_in, int c_out, int kb_per_tap,
                                          const uint32_t* __restrict__ w_amax, uint8_t* __restrict__ image) {
    constexpr int kBPlane = Cfg<TN>::kBPlane;
    constexpr int chunks = TN * (TK / 8);
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_kb = (long long)filter_size * kb_per_tap;
    const long long n_tiles = (c_out + TN - 1) / TN;
    if (t >= n_tiles * n_kb * chunks) return;
    float s, inv_s;
    scale_from_amax(*w_amax, s, inv_s);
    const int chunk = (int)(t % chunks);
    const long long blk = t / chunks;
    const long long kb = blk % n_kb, tile = blk / n_kb;
    const int f = (int)(kb / kb_per_tap), c0 = (int)(kb % kb_per_tap) * TK;
    const int n = chunk % TN, kc = chunk / TN;                              // n fastest: coalesced reads of w
    const int o = (int)tile * TN + n;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = c0 + 8 * kc + 2 * i + j;
            a[j] = (c < c_in && o < c_out) ? __ldg(w + ((long long)f * c_in + c) * c_out + o) * inv_s : 0.f;
        }
        split2(a[0], a[1], hi[i], lo[i]);
    }
    uint8_t* dst = image + blk * (2 * kBPlane) + (n >> 3) * 1024 + (n & 7) * 128 + ((kc ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + kBPlane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------ forward / dgrad
struct GemmArgs {
    const void* nbr;             // (F, n_out_rows) table or nullptr (1x1 layer: row v reads row v)
    const uint8_t* in16;         // h16 image of the input
    const uint8_t* w_image;
    const float* bias;
    float* out;
    const uint32_t* in_amax;
    const uint32_t* w_amax;
    long long n_in_rows, n_out_rows, ld_out;
    int filter_size, c_in, c_out, kb_per_tap, act, out_cm;
    int n_main, acc_stages, total_steps;
    int n_mtiles, n_ntiles;
    int ld16, tma_warps;
};

template <bool I64, int TN>
__global__ void __launch_bounds__(kThreads, 1)
gather_gemm_tma_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                       const __grid_constant__ CUtensorMap map_t_hi, const __grid_constant__ CUtensorMap map_t_lo,
                       const GemmArgs p) {
    using C = Cfg<TN>;
    constexpr uint32_t kIdesc = instr_desc(0, TM, TN, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * C::kStagesA + 2 * C::kStagesB + 4];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base, b_base = smem_base + C::kStagesA * kAStage;
    const uint32_t idx_base = b_base + C::kStagesB * C::kBStage;            // 2 x (kMaxTaps x 128) ints
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full_a = bar0, empty_a = full_a + 8 * C::kStagesA;
    const uint32_t full_b = empty_a + 8 * C::kStagesA, empty_b = full_b + 8 * C::kStagesB;
    const uint32_t acc_full = empty_b + 8 * C::kStagesB, acc_empty = acc_full + 16;
    const uint32_t ready_a = acc_empty + 16;
    const bool gather = p.nbr != nullptr;
    const int tma_warps = gather ? p.tma_warps : 0;             // producer warps whose rows travel by TMA gather4

    if (threadIdx.x == 0) {
        // full_a: one arrival per cp.async thread (arrive.noinc when its copies land) + one arrive.expect_tx per TMA warp;
        // tile mode (1x1 layers): a single arrive.expect_tx
        const uint32_t full_count = gather ? (uint32_t)((kProducerWarps - tma_warps) * 32 + tma_warps) : 1u;
        for (int s = 0; s < C::kStagesA; ++s) { mbar_init(&bars[s], full_count); mbar_init(&bars[C::kStagesA + s], 1); }
        for (int s = 0; s < C::kStagesB; ++s) {
            mbar_init(&bars[2 * C::kStagesA + s], 1);
            mbar_init(&bars[2 * C::kStagesA + C::kStagesB + s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[2 * C::kStagesA + 2 * C::kStagesB + s], 1);
            mbar_init(&bars[2 * C::kStagesA + 2 * C::kStagesB + 2 + s], 4);
        }
        for (int s = 0; s < C::kStagesA; ++s) mbar_init(&bars[2 * C::kStagesA + 2 * C::kStagesB + 4 + s], 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;

    const int n_kb = p.filter_size * p.kb_per_tap;
    const int n_items = p.n_mtiles * p.n_ntiles;               // work item = (128-vertex tile, N tile)
    const int acc_cols = (p.n_main + 1) * TN;                  // TMEM columns of one accumulator set

    if (warp < kProducerWarps && !gather) {
        // ---------------- A producer, 1x1 layers: two TMA tile loads (128 rows x 64 channels, hi and lo) per stage
        if (warp == 0 && lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int v0 = (item / p.n_ntiles) * TM;
                for (int cb = 0; cb < p.kb_per_tap; ++cb) {
                    wait_bar(empty_a + 8 * stage, phase ^ 1);
                    const uint32_t bar = full_a + 8 * stage, dst = a_base + stage * kAStage;
                    mbar_arrive_expect_tx_a(bar, kAStage);
                    tma_tile2d(dst, &map_t_hi, cb * TK, v0, bar);
                    tma_tile2d(dst + kAPlane, &map_t_lo, cb * TK, v0, bar);
                    if (++stage == C::kStagesA) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < kProducerWarps) {
        // ---------------- A producers, gathered rows.  Warp w owns rows 16w .. 16w+15 of the tile.
        //  * The tile's whole index block (F taps x 128 rows) is staged in shared memory ONE ITEM AHEAD (coalesced loads
        //    issued at the start of an item, stored at its end, double-buffered), so the ~1.5 us latency of the neighbour
        //    table is paid once per kernel instead of once per tap (measured: the per-tap chain alone cost 150 us).
        //  * cp.async warps: a quarter warp copies one 128-byte line (one plane of one row) per instruction into its
        //    SWIZZLE_128B position (chunk ^ (row & 7): the 8 lanes cover all 32 banks) and publishes asynchronously
        //    (cp.async.mbarrier.arrive.noinc): the warp never waits for its own copies, the whole ring stays in flight.
        //  * TMA warps (the first `tma_warps`): lanes 0..3 each own a row quad and issue gather4 for both planes; the TMA
        //    unit retires roughly one gather4 per ~48 cycles, so it carries a minority share next to the LSU path.
        const bool by_tma = warp < tma_warps;
        const int rq = lane >> 3, c8 = lane & 7;
        const int ptid = threadIdx.x;                                           // 0 .. 255 among the producers
        const long long row_bytes = 4LL * p.ld16;
        const int plane_bytes = 2 * p.ld16;
        const int n_idx = p.filter_size * TM;                                   // entries of one index block
        int stage = 0;
        uint32_t phase = 0;
        int pre[kIdxPerThread];                                                 // next item's index block, in flight
        auto load_block = [&](int item) {
            const long long v0 = (long long)(item / p.n_ntiles) * TM;
#pragma unroll
            for (int k = 0; k < kIdxPerThread; ++k) {
                const int e = ptid + k * (kProducerWarps * 32);
                const int f = e >> 7, r = e & (TM - 1);
                int val = -1;
                if (e < n_idx && item < n_items && v0 + r < p.n_out_rows) val = load_idx<I64>(p.nbr, (long long)f * p.n_out_rows + v0 + r);
                pre[k] = val;
            }
        };
        auto store_block = [&](int buf) {
#pragma unroll
            for (int k = 0; k < kIdxPerThread; ++k) {
                const int e = ptid + k * (kProducerWarps * 32);
                if (e < n_idx) {
                    const int r = pre[k];
                    const int val = (r < 0 || r >= p.n_in_rows) ? (int)p.n_in_rows : r;          // -> the zero row
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(idx_base + (uint32_t)(buf * kIdxBlock + e) * 4), "r"(val) : "memory");
                }
            }
        };
        int buf = 0;
        load_block(blockIdx.x);
        store_block(0);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            load_block(item + gridDim.x);                                       // lands while this item is being copied
            const uint32_t idx_cur = idx_base + (uint32_t)(buf * kIdxBlock) * 4;
            for (int f = 0; f < p.filter_size; ++f) {
                int idx[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int local = by_tma ? 4 * (lane & 3) + i : 4 * i + rq;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(idx[i]) : "r"(idx_cur + (uint32_t)(f * TM + warp * 16 + local) * 4) : "memory");
                }
                for (int cb = 0; cb < p.kb_per_tap; ++cb) {
                    const int col = cb * TK + c8 * 8;                        // first channel of this lane's 16-byte chunk
                    const uint32_t src_bytes = col < p.ld16 ? 16u : 0u;      // chunks beyond the row are zero-filled
                    wait_bar(empty_a + 8 * stage, phase ^ 1);
                    const uint32_t bar = full_a + 8 * stage;
                    const uint32_t dst = a_base + stage * kAStage;
                    if (by_tma) {
                        if (lane == 0) mbar_arrive_expect_tx_a(bar, kAStage / kProducerWarps);
                        __syncwarp();
                        if (lane < 4) {
                            const uint32_t d = dst + (warp * 16 + 4 * lane) * 128;
                            tma_gather4(d, &map_g_hi, cb * TK, idx[0], idx[1], idx[2], idx[3], bar);
                            tma_gather4(d + kAPlane, &map_g_lo, cb * TK, idx[0], idx[1], idx[2], idx[3], bar);
                        }
                    } else {
                        const uint8_t* src0 = p.in16 + (src_bytes ? (long long)col * 2 : 0);
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int row = warp * 16 + 4 * b + rq;
                            const uint8_t* src = src0 + (src_bytes ? (long long)idx[b] * row_bytes : 0);
                            const uint32_t d = dst + row * 128 + ((c8 ^ (row & 7)) << 4);
                            cp_async16(d, src, src_bytes);
                            cp_async16(d + kAPlane, src + plane_bytes, src_bytes);
                        }
                        // asynchronous publication: the barrier gets this thread's arrival when its copies have landed
                        cp_async_arrive_noinc(bar);
                    }
                    if (++stage == C::kStagesA) { stage = 0; phase ^= 1; }
                }
            }
            buf ^= 1;
            store_block(buf);                                                   // the other buffer was last read one item ago
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    } else if (warp == kBWarp) {
        // ---------------- B producer: one bulk copy per (tap, channel block)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int n_tile = item % p.n_ntiles;
                const uint8_t* src = p.w_image + (long long)n_tile * n_kb * C::kBStage;
                for (int kb = 0; kb < n_kb; ++kb) {
                    wait_bar(empty_b + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx_a(full_b + 8 * stage, C::kBStage);
                    bulk_load_a(b_base + stage * C::kBStage, src, C::kBStage, full_b + 8 * stage);
                    src += C::kBStage;
                    if (++stage == C::kStagesB) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kRelayWarp) {
        // ---------------- proxy-fence relay.  cp.async writes shared memory through the generic proxy and reports its
        // completion asynchronously, so no producer thread can fence afterwards; tcgen05.mma reads through the async proxy.
        // This warp waits for a stage to be full, executes fence.proxy.async and hands the stage on (ready_a).  Keeping the
        // fence (MEMBAR + FENCE.VIEW.ASYNC in SASS) out of the MMA issuer matters: that single thread is the pacing
        // instruction stream of the CTA (ncu: its ~150 scalar instructions per stage were the floor of the whole kernel).
        if (lane == 0) {
            int sa = 0;
            uint32_t pa = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int kb = 0; kb < n_kb; ++kb) {
                    wait_bar(full_a + 8 * sa, pa);
                    fence_proxy_async();
                    mbar_arrive_a(ready_a + 8 * sa);
                    if (++sa == C::kStagesA) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- MMA issuer: one thread.  Everything per stage is incremental (no division / modulo) and the
        // descriptors are a constant high word plus a running 16-byte-unit address.  (Measured alternatives: running the
        // loop warp-converged with an elected issuer removes the ELECT loops ptxas wraps around every tcgen05 instruction
        // of a lone thread, but the per-stage warp synchronisation costs as much as it saves: 0.256 ms either way on cfg2
        // and 3.37 vs 2.87 ms on the 580 -> 1024 layer.)
        if (lane == 0) {
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
            const int last_ksteps = min(TK / 16, (p.c_in - (p.kb_per_tap - 1) * TK + 15) / 16);
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, pacc = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                wait_bar(acc_empty + 8 * acc, pacc ^ 1);                  // epilogue has drained this accumulator set
                fence_after();
                const uint32_t tmem_cross = tmem_d + (uint32_t)(acc * acc_cols);
                // main accumulator of K step s: g(s) = floor(s * n_main / total_steps), tracked incrementally
                int g = 0, g_num = 0, pg = -1, cb = 0;
                uint32_t first = 0;                                        // 0 on the very first K step of the item
                for (int kb = 0; kb < n_kb; ++kb) {
                    const int ksteps = (cb == p.kb_per_tap - 1) ? last_ksteps : TK / 16;
                    if (++cb == p.kb_per_tap) cb = 0;
                    wait_bar(full_b + 8 * sb, pb);
                    wait_bar(ready_a + 8 * sa, pa);
                    fence_after();
                    const uint32_t a16 = (a_base + sa * kAStage) >> 4, b16 = (b_base + sb * C::kBStage) >> 4;
#pragma unroll
                    for (int j = 0; j < TK / 16; ++j) {
                        if (j < ksteps) {
                            const uint64_t dah = ((uint64_t)kDescHi << 32) | (a16 + 2 * j);
                            const uint64_t dal = ((uint64_t)kDescHi << 32) | (a16 + (kAPlane >> 4) + 2 * j);
                            const uint64_t dbh = ((uint64_t)kDescHi << 32) | (b16 + 2 * j);
                            const uint64_t dbl = ((uint64_t)kDescHi << 32) | (b16 + (C::kBPlane >> 4) + 2 * j);
                            umma_f16(tmem_cross, dal, dbh, kIdesc, first);
                            umma_f16(tmem_cross, dah, dbl, kIdesc, 1);
                            umma_f16(tmem_cross + (uint32_t)((1 + g) * TN), dah, dbh, kIdesc, g == pg);
                            first = 1;
                            pg = g;
                            g_num += p.n_main;
                            if (g_num >= p.total_steps) { g_num -= p.total_steps; ++g; }
                        }
                    }
                    umma_commit_a(empty_a + 8 * sa);
                    umma_commit_a(empty_b + 8 * sb);
                    if (++sa == C::kStagesA) { sa = 0; pa ^= 1; }
                    if (++sb == C::kStagesB) { sb = 0; pb ^= 1; }
                }
                umma_commit_a(acc_full + 8 * acc);
                if (++acc == p.acc_stages) { acc = 0; pacc ^= 1; }
            }
        }
    } else {
        // ---------------- epilogue: warps 10..13 own TMEM lane quarters warp % 4
        const int q = warp & 3;
        float s_in, inv_in, s_w, inv_w;
        scale_from_amax(__ldg(p.in_amax), s_in, inv_in);
        scale_from_amax(__ldg(p.w_amax), s_w, inv_w);
        const float s_ab = s_in * s_w;
        int acc = 0;
        uint32_t pacc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int n_tile = item % p.n_ntiles;
            wait_bar(acc_full + 8 * acc, pacc);
            fence_after();
            const int o0 = n_tile * TN;
            const long long m = (long long)(item / p.n_ntiles) * TM + q * 32 + lane;
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols);
#pragma unroll 1
            for (int cb = 0; cb < TN; cb += 16) {
                if (o0 + cb >= p.c_out) break;
                float sum[16];
                uint32_t v[16];
                tmem_ld16(taddr + cb, v);                                    // cross terms
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]) * kLoInv;
                for (int g = 1; g <= p.n_main; ++g) {
                    tmem_ld16(taddr + g * TN + cb, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
                }
                if (m < p.n_out_rows) {
                    float y[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int o = o0 + cb + j;
                        const float b = (p.bias != nullptr && o < p.c_out) ? __ldg(p.bias + o) : 0.f;
                        y[j] = apply_act(fmaf(sum[j], s_ab, b), p.act);
                    }
                    if (!p.out_cm) {
                        float* dst = p.out + m * p.ld_out + o0 + cb;
                        if (o0 + cb + 15 < p.c_out && (p.ld_out & 3) == 0 && ((uintptr_t)p.out & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (o0 + cb + j < p.c_out) dst[j] = y[j];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (o0 + cb + j < p.c_out) p.out[(long long)(o0 + cb + j) * p.ld_out + m] = y[j];
                    }
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(acc_empty + 8 * acc);
            if (++acc == p.acc_stages) { acc = 0; pacc ^= 1; }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after();
        tmem_dealloc(tmem_d, 512);
    }
}

// ------------------------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// (n_rows + 1) x channels halves, row pitch 4 * ld16 bytes, box = 64 channels x box_rows rows, SWIZZLE_128B
bool make_map(CUtensorMap* map, const void* base, long long rows_total, int channels, int ld16, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)ld16 * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void set_attrs() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(gather_gemm_tma_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::kSmem);
    cudaFuncSetAttribute(gather_gemm_tma_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::kSmem);
    cudaFuncSetAttribute(gather_gemm_tma_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem);
    cudaFuncSetAttribute(gather_gemm_tma_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem);
    done = true;
}

int round8(int64_t c) { return (int)((c + 7) / 8 * 8); }

}  // namespace

extern "C" {

int64_t hpl_h16_bytes(int64_t n_rows, int64_t channels) { return (n_rows + 1) * round8(channels) * 4; }

int hpl_h16_split(const float* x, int64_t ld, int64_t n_rows, int64_t channels, const uint32_t* amax, void* x16, void* stream) {
    HPL_CHECK_ARG((x || n_rows == 0) && amax && x16 && channels > 0 && ld >= channels && ld % 4 == 0);
    HPL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)x16 & 15) == 0);
    const int ld16 = round8(channels);
    const long long work = (n_rows + 1) * (ld16 / 8);
    h16_split_kernel<<<(unsigned)((work + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld, n_rows, (int)channels, ld16, amax,
                                                                                 reinterpret_cast<uint4*>(x16));
    HPL_RETURN_LAST();
}

int64_t hpl_blur_gemm_tma_workspace(int64_t filter_size, int64_t c_in, int64_t c_out) {
    const int64_t kb_per_tap = (c_in + TK - 1) / TK, n_cols = (c_out + 127) / 128 * 128;       // covers both tile widths
    return n_cols * filter_size * kb_per_tap * 2 * (TK * 2) + 16;                              // image + the weight absmax slot
}

int hpl_blur_gemm_tma(const void* in16, int64_t n_in_rows, const void* nbr, int idx64, int64_t filter_size, int64_t n_out_rows,
                      int64_t c_in, int64_t c_out, const float* w, const float* bias, int act, float* out, int64_t ld_out,
                      int out_channel_major, void* workspace, const uint32_t* in_amax, void* stream) {
    HPL_CHECK_ARG(in16 && w && out && workspace && in_amax && c_in > 0 && c_out > 0 && filter_size > 0);
    HPL_CHECK_ARG(((uintptr_t)in16 & 15) == 0 && ((uintptr_t)workspace & 127) == 0 && ((uintptr_t)w & 15) == 0);
    HPL_CHECK_ARG(nbr != nullptr || (filter_size == 1 && n_out_rows <= n_in_rows));
    HPL_CHECK_ARG(filter_size <= kMaxTaps);
    HPL_CHECK_ARG(out_channel_major ? ld_out >= n_out_rows : ld_out >= c_out);
    HPL_CHECK_ARG(n_in_rows + 1 < (1LL << 31) && n_out_rows < (1LL << 31) - TM);
    if (n_out_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    set_attrs();

    const int ld16 = round8(c_in);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(in16);
    CUtensorMap g_hi, g_lo, t_hi, t_lo;
    if (!make_map(&g_hi, base, n_in_rows + 1, (int)c_in, ld16, 1) || !make_map(&g_lo, base + 2 * ld16, n_in_rows + 1, (int)c_in, ld16, 1) ||
        !make_map(&t_hi, base, n_in_rows + 1, (int)c_in, ld16, TM) || !make_map(&t_lo, base + 2 * ld16, n_in_rows + 1, (int)c_in, ld16, TM))
        return HPL_EINVAL;

    const int kb_per_tap = (int)((c_in + TK - 1) / TK);
    const bool wide = c_out >= 128;
    const int tn = wide ? 128 : 64;
    const long long n_ntiles = (c_out + tn - 1) / tn;
    const long long image_bytes = n_ntiles * filter_size * kb_per_tap * 2 * (tn * TK * 2);
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* w_amax = reinterpret_cast<uint32_t*>(image + image_bytes);
    const int rc = hpl_absmax(w, filter_size * c_in * c_out, w_amax, stream);
    if (rc != 0) return rc;
    const long long chunks = n_ntiles * filter_size * kb_per_tap * (tn * (TK / 8));
    if (wide)
        weight_image_sw128_kernel<128><<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(w, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);
    else
        weight_image_sw128_kernel<64><<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(w, (int)filter_size, (int)c_in, (int)c_out, kb_per_tap, w_amax, image);

    GemmArgs p;
    p.nbr = nbr; p.w_image = image; p.bias = bias; p.out = out; p.in_amax = in_amax; p.w_amax = w_amax;
    p.n_in_rows = n_in_rows; p.n_out_rows = n_out_rows; p.ld_out = ld_out;
    p.filter_size = (int)filter_size; p.c_in = (int)c_in; p.c_out = (int)c_out; p.kb_per_tap = kb_per_tap;
    p.act = act; p.out_cm = out_channel_major;
    p.total_steps = (int)(filter_size * ((c_in + 15) / 16));
    // accumulate steps per hi.hi accumulator <= ~160-190 (truncation bias of the TMEM accumulate, see gemm_tc.cu)
    p.n_main = p.total_steps <= 160 ? 1 : ((p.total_steps <= 480 || wide) ? 3 : 7);
    const int acc_cols = (p.n_main + 1) * tn;
    p.n_mtiles = (int)((n_out_rows + TM - 1) / TM);
    p.n_ntiles = (int)n_ntiles;
    p.acc_stages = 2 * acc_cols <= 512 ? 2 : 1;               // double-buffered accumulators when TMEM (512 columns) allows
    const long long n_items = (long long)p.n_mtiles * n_ntiles;
    const unsigned grid = (unsigned)(n_items < num_sms() ? n_items : num_sms());
    p.in16 = base;
    p.ld16 = ld16;
    // share of the gathered rows moved by TMA gather4 (in producer warps of 16 rows; the rest goes through cp.async)
    // (HPL_TMA_WARPS = 0..8, default 0: measured on B200 the TMA unit retires ~one gather4 per 48 cycles per SM, ~3 TB/s
    // over the chip, against ~5.8 TB/s for the LSU path, and the hybrid did not beat cp.async alone)
    static int tma_knob = -1;
    if (tma_knob < 0) {
        const char* e = getenv("HPL_TMA_WARPS");
        tma_knob = e ? atoi(e) : 0;
        if (tma_knob < 0 || tma_knob > kProducerWarps) tma_knob = 0;
    }
    p.tma_warps = tma_knob;
#define HPL_LAUNCH_TMA(I64, TNV) \
    gather_gemm_tma_kernel<I64, TNV><<<grid, kThreads, Cfg<TNV>::kSmem, s>>>(g_hi, g_lo, t_hi, t_lo, p)
    if (wide) {
        if (idx64) HPL_LAUNCH_TMA(true, 128); else HPL_LAUNCH_TMA(false, 128);
    } else {
        if (idx64) HPL_LAUNCH_TMA(true, 64); else HPL_LAUNCH_TMA(false, 64);
    }
#undef HPL_LAUNCH_TMA
    HPL_RETURN_LAST();
}

}  // extern "C"
