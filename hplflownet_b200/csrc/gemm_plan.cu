// Engine 5: the lattice convolution (blur forward / data gradient, models/bilateralNN.py:198-221) and its weight gradient
// as persistent tcgen05 kernels that load every DISTINCT neighbour row of a 128-vertex tile ONCE.
//
// Engines 2 / 4 gather F = 15 rows per vertex straight from L2 (1.2 GB of L2 -> SM traffic per cfg2 x 32 launch, the
// measured floor of those designs).  Here a per-lattice tile plan (plan.cu) lists, for every tile of 128 spatially
// coherent vertices, the ~320 distinct rows its 15 x 128 table entries reference; conv5_kernel
//   1. stages those rows once in shared memory (loader warps, cp.async, one 128-byte line per quarter warp,
//      double-buffered: the next phase's rows arrive while the current phase computes),
//   2. builds the UMMA A operand of every tap from the staged rows -- in TENSOR MEMORY (tcgen05.st by the thread that owns
//      the tile row; default) or, for wide K / odd widths, in shared memory (LDS.128 / STS.128, conflict-free on both sides),
//   3. runs 3xFP16 as TWO MMAs per K step instead of three: the weight tile holds [W_hi ; W_lo] as 128 N rows, so
//      x_hi . [W_hi | W_lo] is one M128 N128 K16 instruction (main term in columns 0-63, cross term in 64-127) and
//      x_lo . W_hi (N = 64) accumulates onto the cross columns.
// Operands arrive pre-split ("h16b" image: per row and 32-channel block, 32 fp16 hi | 32 fp16 lo = one 128-byte line;
// x / s = hi + lo * 2^-11, s a per-tensor power of two, see gemm_tc16.cu); producers of lattice rows write that image
// directly (rows.cu), so nothing is converted here.
//
// Work decomposition of conv5_kernel: CTA = one SM, contiguous range of tiles; phase = (tile, 32-channel block); stage =
// one or two taps of a phase.  Warps: 0-15 four copy groups, 16 (17) MMA issuer(s), 18-19 row loaders, 20-23 epilogue
// (tcgen05.ld, bias, activation, fp32 rows in the reference's vertex order, max|out| statistic), overlapped with the next
// tile through double-buffered TMEM accumulators.  Details at the kernel; measurements in DESIGN.md 3.2c.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

extern "C" int64_t hpl_plan_offset(int64_t n_rows, int which);
extern "C" int64_t hpl_plan_tiles(int64_t n_rows);

namespace {

using namespace tc;

constexpr int TM = 128;
constexpr int kTaps = 16;                        // taps the plan's index block holds (F = 15)
constexpr int kUmax = 464;                       // distinct rows per tile (plan.cu)
constexpr int kURow = 128;                       // staged bytes per row and phase: 32 ch hi | 32 ch lo
constexpr int kUBuf = (kUmax + 1) * kURow;       // + the zero row (slot kUmax)
constexpr uint32_t kA_LBO = TM * 16 + 16;        // 2064: K-chunk stride of an A plane (129 x 16 B: odd -> conflict-free)
constexpr int kAPlane = 4 * kA_LBO;              // 8256: hi (or lo) plane of one tap, 32 channels
constexpr int kATap = 2 * kAPlane;               // 16512
constexpr uint32_t kB_LBO = 128 * 16;            // weight tile: 128 N rows ([W_hi ; W_lo]) x 32 K
constexpr int kBTap = 4 * kB_LBO;                // 8192
constexpr int kC5Slots = 4;                      // forward kernel: ring of one-tap stages (shared-memory A path) = number of copy groups
constexpr int kC5SlotsT = 8;                     // ring on the tensor-memory A path (A slots in TMEM, 8 weight tiles in shared memory)
constexpr int kC5OutStage = 128 * (64 * 4 + 16); // tensor-memory A path: staged output tile (272-byte rows)
constexpr int kIdxBuf = kTaps * TM * 2;          // 4096 B of uint16 slots
constexpr int kCopyWarps = 8, kCopyThreads = kCopyWarps * 32;      // weight-gradient kernel: one copy group
constexpr int kMmaWarp = 8;
constexpr int kThreads = 14 * 32;
// forward kernel: FOUR copy groups of 4 warps; group g builds the stages with G % 4 == g, so one group's fixed latencies
// (barrier wake-up, named barriers, fences) overlap the other groups' work
constexpr int kC5CopyWarps = 16;
constexpr int kC5GroupWarps = kC5CopyWarps / 4;                     // four copy groups of four warps: group g fills ring slot g
constexpr int kC5MmaWarp = 16, kC5LoadWarp = 18, kC5LoadWarps = 2;  // warps 16, 17: MMA issuers; 18, 19: row loaders; 20-23: epilogue
constexpr int kC5LoadThreads = kC5LoadWarps * 32;
constexpr int kC5Threads = 24 * 32;
constexpr int kC5RingBytes = kC5Slots * (kATap + kBTap) > kC5OutStage + kC5SlotsT * kBTap ? kC5Slots * (kATap + kBTap)
                                                                                          : kC5OutStage + kC5SlotsT * kBTap;
constexpr int kC5UniqBuf = 2048;                 // a tile's list of distinct rows (kUmax ints), staged one tile ahead
constexpr int kSmem = 2 * kUBuf + kC5RingBytes + 2 * kIdxBuf + kC5UniqBuf + 1024;
constexpr int kUIters = (kUmax * 8 + kCopyThreads - 1) / kCopyThreads;      // 15 row-chunk copies per thread and phase
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
static_assert(kSmem <= 232448, "shared memory budget");

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

// bounded wait: a barrier that never flips (descriptor / byte-count bug) traps after ~4 s instead of hanging the GPU.
// The suspend-time hint lets the thread sleep in hardware until the phase completes.
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t tries = 0; !done; ++tries) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x100000u)
            : "memory");
        if (!done && tries >= 64) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}


// Un-rotate the eight 16-byte chunks a copy thread read in lane-rotated order (v[i] = chunk (i + lane) % 8): a barrel of
// three conditional rotations, 96 SEL.  (Selecting arithmetically on the FMA pipe -- a * m + b * (1 - m), two IMADs -- for
// one of the stages was slower, 0.0961 -> 0.0996 ms: the copy warps are bound by issue slots, not by the ALU pipe.)
__device__ __forceinline__ void unrotate8(uint4* v, int lane) {
#pragma unroll
    for (int b = 1; b < 8; b <<= 1) {
        uint4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = (lane & b) ? v[(i - b) & 7] : v[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = t[i];
    }
}

// tcgen05.mma with the A operand in tensor memory (M = 128: lane = row, every 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 registers of every lane -> 32 consecutive tensor-memory columns of the warp's lane quarter
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint4* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
        "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w),
        "r"(v[4].x), "r"(v[4].y), "r"(v[4].z), "r"(v[4].w), "r"(v[5].x), "r"(v[5].y), "r"(v[5].z), "r"(v[5].w),
        "r"(v[6].x), "r"(v[6].y), "r"(v[6].z), "r"(v[6].w), "r"(v[7].x), "r"(v[7].y), "r"(v[7].z), "r"(v[7].w)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// one row of the output tile: shared memory -> global memory through the bulk-copy engine (no LSU store transactions)
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------- h16b image
// x (n_rows, ld) fp32 vertex-major -> (n_rows, CB) lines of [32 hi | 32 lo] halves, CB = ceil(C / 32); channels beyond C
// are zero.  norm != nullptr: x[v, :] is first multiplied by 1 / (norm[v] + 1e-5) (density normalisation,
// bilateralNN.py:185-186, fused into the split).  One thread = 8 channels of one row.
__global__ void h16b_split_kernel(const float* __restrict__ x, long long ld, long long n_rows, int channels, int cb_count,
                                  const float* __restrict__ norm, const uint32_t* __restrict__ amax, uint4* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cpr = cb_count * 4;                                         // 8-channel groups per row
    if (t >= n_rows * cpr) return;
    const long long row = t / cpr;
    const int g = (int)(t - row * cpr);
    const int c0 = g * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    const float* p = x + row * ld + c0;
    if (c0 + 8 <= channels) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c0 + i < channels) v[i] = __ldg(p + i);
    }
    float s, inv_s;
    scale_from_amax(__ldg(amax), s, inv_s);
    if (norm != nullptr) {
        const float r = 1.0f / (__ldg(norm + row) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= r;                             // same rounding as normalize_rows_kernel
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i] * inv_s, v[2 * i + 1] * inv_s, hi[i], lo[i]);
    // line (row, cb = g / 4): 8 chunks of 16 B -- hi chunks 0..3, lo chunks 4..7
    uint4* dst = out + (row * cb_count + (g >> 2)) * 8 + (g & 3);
    dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[4] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// The same split with the passes around it folded in (one trip over the rows instead of three or four):
//   forward : x = splat accumulators, norm = weight sums -> normalised image, inv_out[v] = 1 / (norm[v] + 1e-5) (for the
//             backward of the splat), norm_amax_out <- max norm (bound of the slice-backward magnitudes, see below);
//   backward: x = dz, y = the layer's saved output -> x *= act'(y), image of the result, colsum += column sums (bias
//             gradient); the operand scale comes from *amax_a (exact maximum) or from the BOUND *amax_a x *amax_b
//             (|dz[v]| <= max|g| x sum of barycentric weights at v for a slice backward) so that no separate max pass is
//             needed; amax_out <- the value the image was scaled with (what the contraction kernels read).
//   dispose : 0 keep x, 1 write the act'-scaled values back (a consumer still needs fp32 rows), 2 zero x (the buffer goes
//             back to the zero pool: the next splat accumulates into it without a memset).
__device__ __forceinline__ uint32_t bound_bits(const uint32_t* a, const uint32_t* b) {
    const uint32_t ba = __ldg(a);
    if (b == nullptr) return ba;
    const uint32_t bb = __ldg(b);
    if (ba == 0 || bb == 0) return 0;
    int e = (int)((ba >> 23) & 0xff) + (int)((bb >> 23) & 0xff) - 254 + 2;           // a < 2^(ea+1), b < 2^(eb+1): a b < 2^(ea+eb+2)
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
    return (uint32_t)(e + 127) << 23;
}

// CSR = true: the rows are not read from x but GATHERED -- v = sum over the vertex's splat contributions
// w(e) * src[point(e), :], csr_ptr / csr_ent = the contributions sorted by vertex (plans.py: splat plan; entry = point |
// remainder << 30, weight = bary[remainder, point]).  This is the splat (bilateralNN.py:150-182) as a deterministic gather:
// no atomics, no accumulator, every lattice row written once, straight into the operand image.  norm_self: normalise by the
// row's own weight sum (computed in the same loop).
struct CsrArgs {
    const float* src;            // (n_points, ld_src) point-major source rows
    long long ld_src;
    const float* bary;           // (4, n_points)
    long long n_points;
    const int* ptr;              // (n_rows + 1)
    const int* ent;
    int norm_self;
};

template <bool CSR>
__global__ void __launch_bounds__(256)
h16b_split_fused_kernel(float* __restrict__ x, long long ld, long long n_rows, int channels, int cb_count,
                        const float* __restrict__ norm, float* __restrict__ inv_out, uint32_t* __restrict__ norm_amax_out,
                        const float* __restrict__ y, long long ld_y, float slope, const uint32_t* __restrict__ amax_a,
                        const uint32_t* __restrict__ amax_b, uint32_t* __restrict__ amax_out, float* __restrict__ colsum,
                        int dispose, uint4* __restrict__ out, const CsrArgs csr) {
    __shared__ float part[256][9];
    const int cpr = cb_count * 4;                                         // 8-channel groups per row
    const int rows_per_iter = 256 / cpr;
    const int r_in = threadIdx.x / cpr, g = threadIdx.x - r_in * cpr;
    const bool worker = r_in < rows_per_iter;
    const int c0 = g * 8;
    const uint32_t bits = bound_bits(amax_a, amax_b);
    if (amax_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *amax_out = bits;
    float s, inv_s;
    scale_from_amax(bits, s, inv_s);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float n_max = 0.f;
    for (long long row = (long long)blockIdx.x * rows_per_iter + r_in; worker && row < n_rows; row += (long long)gridDim.x * rows_per_iter) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        float* p = CSR ? nullptr : x + row * ld + c0;
        const bool full8 = c0 + 8 <= channels;
        float w_self = 0.f;
        if (CSR) {
            // the 8 threads of a row read one contiguous source row per contribution (8 x 32 B); contributions in the plan's
            // fixed order -> the sum is reproducible
            const int e1 = __ldg(csr.ptr + row + 1);
            for (int e = __ldg(csr.ptr + row); e < e1; ++e) {
                const int ent = __ldg(csr.ent + e);
                const long long pt = ent & 0x3fffffff;
                const float w = __ldg(csr.bary + (long long)((unsigned)ent >> 30) * csr.n_points + pt);
                const float* q = csr.src + pt * csr.ld_src + c0;
                if (full8) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(q)), b = __ldg(reinterpret_cast<const float4*>(q + 4));
                    v[0] = fmaf(w, a.x, v[0]); v[1] = fmaf(w, a.y, v[1]); v[2] = fmaf(w, a.z, v[2]); v[3] = fmaf(w, a.w, v[3]);
                    v[4] = fmaf(w, b.x, v[4]); v[5] = fmaf(w, b.y, v[5]); v[6] = fmaf(w, b.z, v[6]); v[7] = fmaf(w, b.w, v[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c0 + i < channels) v[i] = fmaf(w, __ldg(q + i), v[i]);
                }
                w_self += w;
            }
        } else if (full8) {
            const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + i < channels) v[i] = p[i];
        }
        if (norm != nullptr || (CSR && csr.norm_self)) {
            const float w = (CSR && csr.norm_self) ? w_self : __ldg(norm + row);
            const float r = 1.0f / (w + 1e-5f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= r;                         // same rounding as normalize_rows_kernel
            if (g == 0) {
                if (inv_out != nullptr) inv_out[row] = r;
                n_max = fmaxf(n_max, w);
            }
        }
        if (y != nullptr) {
            const float* py = y + row * ld_y + c0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + i < channels) v[i] *= __ldg(py + i) > 0.f ? 1.f : slope;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
        if (!CSR && dispose != 0) {
            const float z = dispose == 2 ? 0.f : 1.f;
            if (full8) {
                *reinterpret_cast<float4*>(p) = make_float4(v[0] * z, v[1] * z, v[2] * z, v[3] * z);
                *reinterpret_cast<float4*>(p + 4) = make_float4(v[4] * z, v[5] * z, v[6] * z, v[7] * z);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c0 + i < channels) p[i] = v[i] * z;
            }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(v[2 * i] * inv_s, v[2 * i + 1] * inv_s, hi[i], lo[i]);
        uint4* dst = out + (row * cb_count + (g >> 2)) * 8 + (g & 3);
        dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[4] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (norm_amax_out != nullptr) {                                       // (block-uniform condition)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_max = fmaxf(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
        if ((threadIdx.x & 31) == 0 && n_max > 0.f && __float_as_uint(n_max) > *reinterpret_cast<volatile uint32_t*>(norm_amax_out))
            atomicMax(norm_amax_out, __float_as_uint(n_max));
    }
    if (colsum != nullptr) {                                              // (block-uniform condition)
#pragma unroll
        for (int i = 0; i < 8; ++i) part[threadIdx.x][i] = worker ? acc[i] : 0.f;
        __syncthreads();
        if (threadIdx.x < cpr * 8) {                                      // one thread per channel slot of the (padded) row
            const int gg = threadIdx.x >> 3, i = threadIdx.x & 7;
            float t = 0.f;
            for (int r = 0; r < rows_per_iter; ++r) t += part[r * cpr + gg][i];
            const int c = gg * 8 + i;
            if (c < channels) atomicAdd(colsum + c, t);
        }
    }
}

// ---------------------------------------------------------------------------------------- weight image
// w[f, c, o] (strided) -> per (cb, stage = tap pair): two 8 KB tiles; tile rows n = 0..127 are [W_hi(o = n) ; W_lo(o = n - 64)],
// K = 32 channels of block cb, K-major no-swizzle: chunk kc at kc * 2048 + (n / 8) * 128 + (n % 8) * 16.
// tap_map (may be NULL): image tap g holds source tap tap_map[g] (the data gradient uses the mirrored tap).
__global__ void weight_image5_kernel(const float* __restrict__ w, long long w_sf, long long w_sc, long long w_so, int filter_size,
                                     int c_in, int c_out, int cb_count, const int* __restrict__ tap_map,
                                     const uint32_t* __restrict__ w_amax, uint8_t* __restrict__ image) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)cb_count * kTaps * 64 * 4;          // (cb, tap, o, kc)
    if (t >= total) return;
    const int o = (int)(t & 63);
    const int kc = (int)((t >> 6) & 3);
    const int tap = (int)((t >> 8) & (kTaps - 1));
    const int cb = (int)(t >> 12);
    float s, inv_s;
    scale_from_amax(*w_amax, s, inv_s);
    const int f = tap < filter_size ? (tap_map != nullptr ? tap_map[tap] : tap) : -1;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = cb * 32 + kc * 8 + 2 * i + j;
            a[j] = (f >= 0 && c < c_in && o < c_out) ? __ldg(w + f * w_sf + c * w_sc + o * w_so) * inv_s : 0.f;
        }
        split2(a[0], a[1], hi[i], lo[i]);
    }
    uint8_t* tile = image + ((long long)cb * kTaps + tap) * kBTap;
    uint8_t* dst = tile + kc * kB_LBO + (o >> 3) * 128 + (o & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + 64 * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);          // rows 64..127
}

// ---------------------------------------------------------------------------------------- the kernel
struct Conv5Args {
    const uint8_t* in16;          // h16b image of the input rows
    const uint8_t* w_image;
    const int* tile_rows;         // plan
    const int* n_uniq;
    const int* uniq;
    const unsigned short* local;
    const float* bias;
    float* out;
    const uint32_t* in_amax;
    const uint32_t* w_amax;
    uint32_t* out_amax;
    long long ld_out;
    int n_tiles, cb_count, filter_size, c_out, act, n_main, steps_total;
    long long* trace;             // timing experiments: clock64 stamps of CTA 0 (HPL_CONV5_TRACE), 8 per stage
    int dbg;                      // timing experiments (HPL_CONV5_DBG): 1 no A copies, 2 no MMAs, 4 no W loads, 8 no U loads, 16 no stores
};

// F = taps processed (15: HPLFlowNet's r = 1 neighbourhood; 16: any other count, padded with zero taps)
//
// Pipeline.  The stages of a CTA are numbered G = 0, 1, 2, ... across phases and tiles; copy group g (4 warps) builds the
// stages with G % 4 == g, the ring has NS slots (slot = G % NS).  Measured on earlier structures: a slot's
// copy -> MMA -> copy chain carries ~1750 cycles of hand-over latency per round trip (so the ring must be deep enough to
// hide it), the SM serialises mbarrier operations (so there is ONE wait and ONE arrival per group and stage, executed by
// the group's whole first warp / one thread, with named barriers inside the group), and a warp must never wait with a
// single lane in front of a named barrier (see the copy loop).
// AT = false (shared-memory A path; wide K with n_main > 1, odd c_out): a stage is one tap (A: 16.1 KB hi | lo planes in
// the K-major no-swizzle UMMA layout, W: 8 KB), NS = 4, two MMA issuers on alternate stages with their own accumulators.
// AT = true (default): the A operand lives in TENSOR MEMORY.  Every copy thread owns one row of the tile, reads its staged
// line from shared memory (chunk order rotated by the lane so that a quarter warp hits 8 different bank groups,
// un-rotated in registers) and writes it with tcgen05.st; the MMA then fetches only the weight tile from shared memory
// (71 -> 39 KB of shared-memory traffic per tap).  The shared memory the A stages no longer need stages the output tile,
// which leaves through the bulk-copy engine (one 256-byte row per instruction) instead of 32-lines-per-instruction LSU
// stores.  One issuer; tensor memory: accumulators 2 x 128 columns (double-buffered), A slots NS x 32 * TPS columns.
// DBG: the instantiation with the experiment hooks (HPL_CONV5_DBG ablation bits, HPL_CONV5_TRACE clock stamps); the
// production instantiation carries neither their tests nor their address arithmetic.
// TPS: taps per stage (tensor-memory path only: 2).  The per-stage costs that do not scale with the data -- the issuer's
// barrier poll / fence / commit (~400 cycles per stage on one thread, which paced the one-tap version), the groups'
// empty wait and two named barriers -- are then paid once per TWO taps; a phase of F = 15 taps is 7 double stages and one
// single stage.
template <int F, bool AT, bool DBG, int TPS>
__global__ void __launch_bounds__(kC5Threads, 1) conv5_kernel(const Conv5Args p) {
    static_assert(TPS == 1 || (AT && TPS == 2), "two-tap stages exist on the tensor-memory path only");
    const int dbg = DBG ? p.dbg : 0;
    extern __shared__ uint8_t smem_raw[];
    constexpr int SP = (F + TPS - 1) / TPS;                              // stages per phase
    constexpr int NS = AT ? kC5SlotsT / TPS : kC5Slots;                  // ring slots (group g fills slots g, g + 4, ...)
    constexpr uint32_t kBSlot = kBTap * TPS;                             // weight tiles of one stage
    constexpr uint32_t kASlotCols = 32 * TPS;                            // TMEM columns of one A slot
    __shared__ __align__(8) uint64_t bars[2 * kC5SlotsT + 4 + 4];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[64];                           // bias (zero beyond c_out / without one)

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t u_base = smem_base;                                   // 2 x kUBuf
    const uint32_t a_base = u_base + 2 * kUBuf;                          // kC5Slots x kATap; AT: the output staging tile instead
    const uint32_t b_base = a_base + (AT ? kC5OutStage : kC5Slots * kATap);   // NS x kBTap
    const uint32_t idx_base = a_base + kC5RingBytes;                     // 2 x kIdxBuf
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full = bar0, empty = full + 8 * NS, ufull = empty + 8 * NS, ufree = ufull + 16;
    const uint32_t acc_full = ufree + 16, acc_empty = acc_full + 16;

    const int per = (p.n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per;
    const int t_end = min(p.n_tiles, t_begin + per);
    const int n_my = max(0, t_end - t_begin);
    const int CB = p.cb_count;
    const int n_phases = n_my * CB;
    const int n_stages = n_phases * SP;
    // MMA issuers: two on the shared-memory A path (alternate stages, own accumulators); ONE on the tensor-memory path, whose
    // accumulators (128 columns per tile) are then double-buffered next to the A slots, so the epilogue of a tile
    // overlaps the next tile's stages.
    constexpr int NI = AT ? 1 : 2;
    const int acc_cols = NI * p.n_main * 128;                            // TMEM columns of one tile: issuers x n_main x [main | cross]
    const int acc_stages = (2 * acc_cols + (AT ? NS * (int)kASlotCols : 0) <= 512) ? 2 : 1;
    constexpr uint32_t kTmemA = 256;                                     // AT: first column of the A slots (32 columns each)
    constexpr uint32_t kOutPitch = 64 * 4 + 16;                          // AT: staged output row (272 B: odd multiple of 16 -> conflict-free)

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[NS + s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[2 * NS + s], 1);                            // ufull: the phase's rows have landed (one arrival for the loaders)
            mbar_init(&bars[2 * NS + 2 + s], kC5Slots);            // ufree: every copy group is done reading the buffer
            mbar_init(&bars[2 * NS + 4 + s], NI);                  // acc_full: every MMA issuer has committed the tile
            mbar_init(&bars[2 * NS + 6 + s], 4);
        }
        fence_mbar_init();
    }
    if (warp == kC5MmaWarp) tmem_alloc(&tmem_slot, 512);
    // zero rows of both U buffers (slot kUmax): never overwritten (a tile stages at most kUmax rows)
    if (threadIdx.x < 16) {
        const uint32_t z = u_base + (threadIdx.x >> 3) * kUBuf + kUmax * kURow + (threadIdx.x & 7) * 16;
        sts128(z, 0, 0, 0, 0);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int o = threadIdx.x - 64;
        bias_s[o] = (p.bias != nullptr && o < p.c_out) ? __ldg(p.bias + o) : 0.f;
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp < kC5CopyWarps) {
        // ------------------------------------------------------------------ copy warps
        // Their instruction stream paces the kernel (a first version with run-time tap / modulo tests ran ~500 dependent
        // instructions per warp and stage), so addresses are per-thread constants plus a few adds: per stage 4 slot loads,
        // 4 LDS.128, 4 STS.128 per thread; the staged rows are read BEFORE the wait for the stage buffer.
        const int tid = threadIdx.x;
        const int grp = tid >> 7;                                        // copy group = ring slot it fills (stages G % 4 == grp)
        const int gt = tid & 127;
        const bool lead_warp = (warp & (kC5GroupWarps - 1)) == 0;        // the group's first warp (warp-uniform)
        const int c8 = gt & 7;                                           // 16-byte chunk of a 128-byte line
        const int rg = gt >> 3;                                          // 0..15: row within a group of 16
        const uint32_t src_off = c8 * 16;
        const uint32_t idx_off = rg * 2;                                 // + (tap * 128 + i * 16) * 2
        // A position of (row = i * 16 + rg, chunk c8): + i * 256
        const uint32_t dst_off = (c8 >> 2) * kAPlane + (c8 & 3) * kA_LBO + (rg >> 3) * 128 + (rg & 7) * 16;
        const uint32_t abd = a_base + grp * kATap + dst_off;

        int ph = -1, k = 0, cb = 0;                                       // current phase and its (tile, channel block)
        int next_ph_tap0 = 0;                                            // stage number where the next phase starts
        uint32_t ub = 0, ibs = 0;
        for (int G = grp; G < n_stages; G += kC5Slots) {
            if (G >= next_ph_tap0) {                                     // this group's first stage of a new phase
                // (the group's last bar.sync of the previous stage came after every one of its reads of that phase's rows)
                if (ph >= 0) {
                    if (gt == 0) mbar_arrive_a(ufree + 8 * (ph & 1));    // this group is done reading the previous phase's rows
                    if (++cb == CB) { cb = 0; ++k; }
                }
                ++ph;
                next_ph_tap0 += SP;
                if (lead_warp) wait_bar(ufull + 8 * (ph & 1), (ph >> 1) & 1);
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kC5GroupWarps * 32) : "memory");
                ub = u_base + (ph & 1) * kUBuf + src_off;
                ibs = idx_base + (k & 1) * kIdxBuf + idx_off;
            }
            const int tap = (G - (next_ph_tap0 - SP)) * TPS;                 // (first) tap of the stage
            const bool two = TPS == 2 && tap + 1 < F;                        // F odd: the phase's last stage holds one tap
            const uint32_t w_bytes = two ? 2 * kBTap : kBTap;
            const int slot_g = G & (NS - 1);                                 // (NS == 4: always grp)
            const uint32_t full_g = full + 8 * slot_g, empty_g = empty + 8 * slot_g;
            const uint32_t par_g = (uint32_t)(G / NS) & 1;
            const bool tr = DBG && p.trace != nullptr && blockIdx.x == 0 && gt == 0 && G < 256;
            long long* trp = p.trace + G * 8;
            if (tr) trp[0] = clock64();
            uint32_t slot[8];
            uint4 v[8];
            if (AT) {
                // thread = tile row gt: its staged row, 8 chunks in lane-rotated order (conflict-free), then un-rotated
                if (!(dbg & 1)) {
                    const uint32_t sl = lds_u16(idx_base + (k & 1) * kIdxBuf + (tap * TM + gt) * 2);
                    const uint32_t src = u_base + (ph & 1) * kUBuf + sl * kURow;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = lds128(src + (((i + lane) & 7) << 4));
                    unrotate8(v, lane);                                  // v[i] held chunk (i + lane) % 8
                }
            } else if (!(dbg & 1)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) slot[i] = lds_u16(ibs + tap * (TM * 2) + i * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = lds128(ub + slot[i] * kURow);
            }
            // One mbarrier wait and one arrival per GROUP and stage (named barriers inside the group): the SM's mbarrier
            // unit serialises its operations (~40 cycles each, measured: an empty pipeline with one wait + one arrival per
            // WARP ran at ~500 cycles per stage), so they are kept off the per-warp path.
            if (tr) trp[1] = clock64() + (AT ? (v[0].x & 0) : 0);          // (the row reads have landed)
            // The waiting is done by the group's whole first warp -- a warp-uniform branch.  (A single waiting lane leaves
            // its warp diverged in front of the named barrier; that variant dead-locked in one build and, where it ran, cost
            // ~1000 cycles per wait in the reconvergence.)
            if (lead_warp) {
                wait_bar(empty_g, par_g ^ 1);
                // the stage's weight tile (8 KB, one bulk copy) is requested by the group itself as soon as the slot is free
                // (its complete_tx may land before the expect_tx below: the transaction count may go negative meanwhile)
                if (!(dbg & 4) && elect_one())
                    bulk_load_a(b_base + slot_g * kBSlot, p.w_image + ((long long)cb * kTaps + tap) * kBTap, w_bytes, full_g);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kC5GroupWarps * 32) : "memory");
            if (tr) trp[2] = clock64();
            if (AT) {
                if (!(dbg & 1)) {
                    fence_after();                                       // (the slot's previous MMAs, observed through `empty`)
                    const uint32_t ta = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + kTmemA + slot_g * kASlotCols;
                    tmem_st32(ta, v);
                    if (two) {                                           // second tap: read, un-rotate, store (the registers are free again)
                        const uint32_t sl = lds_u16(idx_base + (k & 1) * kIdxBuf + ((tap + 1) * TM + gt) * 2);
                        const uint32_t src = u_base + (ph & 1) * kUBuf + sl * kURow;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = lds128(src + (((i + lane) & 7) << 4));
                        unrotate8(v, lane);
                        tmem_st32(ta + 32, v);
                    }
                }
                fence_before();
            } else {
                if (!(dbg & 1)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) sts128(abd + i * 256, v[i].x, v[i].y, v[i].z, v[i].w);
                }
                if (!(dbg & 32)) fence_proxy_async();
            }
            if (tr) trp[3] = clock64();
            asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kC5GroupWarps * 32) : "memory");
            if (gt == 0) {
                if (dbg & 4) mbar_arrive_a(full_g);
                else mbar_arrive_expect_tx_a(full_g, w_bytes);
            }
            if (tr) trp[4] = clock64();
        }
    } else if (warp >= kC5LoadWarp && warp < kC5LoadWarp + kC5LoadWarps) {
        // ------------------------------------------------------------------ row loaders
        // Stage the distinct rows (and the index block) of the phases, one phase ahead of the copy groups.  They are separate
        // warps because fence.proxy.async -- which every copy thread executes once per stage -- waits for the thread's
        // outstanding cp.async as well: with the loads issued by the copy threads themselves, half of their stages
        // stalled ~1500 cycles (an L2 round trip) in that fence.
        const int lt = threadIdx.x - kC5LoadWarp * 32;                    // 0 .. 63
        const int c8 = lt & 7, r0 = lt >> 3;                              // chunk of the line; rows r0, r0 + 8, ...
        const long long row_bytes = (long long)CB * kURow;
        // The tile's list of distinct rows is read from SHARED memory: it is fetched (cp.async) while the previous tile's
        // last phase loads.  Read with __ldg at the point of use, the cold list cost one DRAM round trip per pass of 64
        // rows -- 8 serialised round trips at every tile boundary, ~3500 idle cycles per tile for the whole CTA.
        const uint32_t uq_base = idx_base + 2 * kIdxBuf;
        auto fetch_list = [&](int t) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.uniq + (long long)t * kUmax);
            for (int c = lt; c < kUmax * 4 / 16; c += kC5LoadThreads) cp_async16(uq_base + c * 16, src + c * 16);
        };
        int n_cur = 0, n_next = 0;
        if (n_phases > 0) {
            fetch_list(t_begin);
            n_cur = min(__ldg(p.n_uniq + t_begin), kUmax);
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync %0, %1;" ::"n"(kC5Slots + 1), "n"(kC5LoadThreads) : "memory");
        }
        int k = 0, cb = 0;
        for (int ph = 0; ph < n_phases; ++ph) {
            // the buffer held phase ph - 2: every copy warp has left it
            if (ph >= 2) {
                if (warp == kC5LoadWarp) wait_bar(ufree + 8 * (ph & 1), ((ph - 2) >> 1) & 1);   // (the whole warp: see the copy groups)
                asm volatile("bar.sync %0, %1;" ::"n"(kC5Slots + 1), "n"(kC5LoadThreads) : "memory");
            }
            const int t = t_begin + k;
            const uint32_t ub = u_base + (ph & 1) * kUBuf + c8 * 16;
            if (cb == 0) {                                               // the tile's index block: 256 chunks of 16 B
#pragma unroll
                for (int i = 0; i < 256 / kC5LoadThreads; ++i)
                    cp_async16(idx_base + (k & 1) * kIdxBuf + (lt + i * kC5LoadThreads) * 16,
                               reinterpret_cast<const uint8_t*>(p.local) + ((long long)t * (kTaps * TM)) * 2 + (lt + i * kC5LoadThreads) * 16);
            }
            const int n = n_cur;
            const uint8_t* src = p.in16 + cb * kURow + c8 * 16;
            if (!(dbg & 8)) {
                for (int j0 = r0; j0 < n; j0 += 8 * 8) {                 // 8 rows in flight per thread
                    int rows[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) rows[i] = j0 + 8 * i < n ? lds_s32(uq_base + (j0 + 8 * i) * 4) : -1;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (rows[i] >= 0) cp_async16(ub + (j0 + 8 * i) * kURow, src + (long long)rows[i] * row_bytes);
                }
            }
            if (cb == CB - 1 && k + 1 < n_my) {                          // the tile's last phase: fetch the next tile's list
                asm volatile("bar.sync %0, %1;" ::"n"(kC5Slots + 1), "n"(kC5LoadThreads) : "memory");   // (everyone has read this one)
                fetch_list(t + 1);
                n_next = min(__ldg(p.n_uniq + t + 1), kUmax);
            }
            // one mbarrier arrival per phase (not one asynchronous arrival per thread): wait for the own copies, meet, arrive
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync %0, %1;" ::"n"(kC5Slots + 1), "n"(kC5LoadThreads) : "memory");
            if (lt == 0) mbar_arrive_a(ufull + 8 * (ph & 1));
            if (++cb == CB) { cb = 0; ++k; n_cur = n_next; }
        }
    } else if (warp == kC5MmaWarp || (NI == 2 && warp == kC5MmaWarp + 1)) {
        // ------------------------------------------------------------------ MMA issuers
        // The tensor core needs 64 / 32 cycles per K step of the N = 128 / N = 64 instruction (tools/umma_rate.cu: a tight
        // issue loop reaches exactly that), 192 cycles per stage; what an issuer adds is its own instruction stream -- the
        // barrier poll, the fence, descriptor arithmetic -- so that stream is kept short and warp-uniform.  On the
        // shared-memory A path two issuers take alternate stages (issuer i: G % 2 == i); each accumulates into its own
        // TMEM columns (the order of tcgen05.mma between threads is not defined, and separate accumulators keep the
        // result deterministic); the epilogue adds them.
        // The issuer warp stays converged and one elected lane issues (see elect_one in tc_common.cuh).
        {
            const int me = warp - kC5MmaWarp;
            constexpr uint32_t kIdescMain = instr_desc(0, TM, 128, 0, 0);
            constexpr uint32_t kIdescLo = instr_desc(0, TM, 64, 0, 0);
            constexpr uint64_t kDescA = (uint64_t)((kA_LBO >> 4) & 0x3fff) << 16 | (uint64_t)(128 >> 4) << 32 | (uint64_t)1 << 46;
            constexpr uint64_t kDescB = (uint64_t)((kB_LBO >> 4) & 0x3fff) << 16 | (uint64_t)(128 >> 4) << 32 | (uint64_t)1 << 46;
            const int per_tile = CB * SP;
            const int n_main = AT ? 1 : p.n_main;
            int acc = 0;
            uint32_t pacc = 0;
            for (int k = 0; k < n_my; ++k) {
                wait_bar(acc_empty + 8 * acc, pacc ^ 1);
                fence_after();
                const uint32_t d0 = tmem_d + (uint32_t)(acc * acc_cols + me * n_main * 128);
                const int G0 = k * per_tile;
                int pg = -1;
                for (int s = NI == 2 ? ((G0 + me) & 1) : 0; s < per_tile; s += NI) {   // two issuers: stages with (G0 + s) % 2 == me
                    const int G = G0 + s;
                    const int slot_i = G & (NS - 1);
                    const bool tr = DBG && p.trace != nullptr && blockIdx.x == 0 && G < 256 && lane == 0;
                    if (tr) p.trace[G * 8 + 5] = clock64();
                    // (every lane polls: a lane-0 poll followed by __syncwarp() cost ~1000 cycles per stage)
                    wait_bar(full + 8 * slot_i, (uint32_t)(G / NS) & 1);
                    fence_after();
                    if (tr) p.trace[G * 8 + 6] = clock64();
                    const uint32_t a16 = (a_base + slot_i * kATap) >> 4, b16 = (b_base + slot_i * kBSlot) >> 4;
                    const int n_taps = (TPS == 2 && !(F % 2 == 1 && s % SP == SP - 1)) ? 2 : 1;
                    const int g0 = n_main == 1 ? 0 : ((2 * s) * n_main) / p.steps_total;         // main accumulator of the K steps
                    const int g1 = n_main == 1 ? 0 : ((2 * s + 1) * n_main) / p.steps_total;
                    if (elect_one()) {
                        if (!(dbg & 2)) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int g = j == 0 ? g0 : g1;
                                const uint32_t ah = a16 + ((j * 2 * kA_LBO) >> 4);
                                const uint32_t bb = b16 + ((j * 2 * kB_LBO) >> 4);
                                const uint32_t dg = d0 + (uint32_t)(g * 128);
                                if (AT) {
                                    const uint32_t at = tmem_d + kTmemA + (uint32_t)(slot_i * kASlotCols + 8 * j);
                                    umma_f16_ts(dg, at, kDescB | bb, kIdescMain, g == pg);                 // x_hi . [W_hi | W_lo]
                                    umma_f16_ts(dg + 64, at + 16, kDescB | bb, kIdescLo, 1);               // x_lo . W_hi
                                    if (n_taps == 2) {                                                     // the stage's second tap
                                        const uint32_t bb2 = bb + (kBTap >> 4);
                                        umma_f16_ts(dg, at + 32, kDescB | bb2, kIdescMain, 1);
                                        umma_f16_ts(dg + 64, at + 32 + 16, kDescB | bb2, kIdescLo, 1);
                                    }
                                } else {
                                    umma_f16(dg, kDescA | ah, kDescB | bb, kIdescMain, g == pg);           // x_hi . [W_hi | W_lo]
                                    umma_f16(dg + 64, kDescA | (ah + (kAPlane >> 4)), kDescB | bb, kIdescLo, 1);   // x_lo . W_hi
                                }
                                pg = g;
                            }
                        }
                        umma_commit_a(empty + 8 * slot_i);
                    }
                    pg = g1;
                    if (tr) p.trace[G * 8 + 7] = clock64();
                }
                if (elect_one()) umma_commit_a(acc_full + 8 * acc);
                if (++acc == acc_stages) { acc = 0; pacc ^= 1; }
            }
        }
    } else if (warp >= kC5LoadWarp + kC5LoadWarps) {
        // ------------------------------------------------------------------ epilogue (TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        float s_in, inv_in, s_w, inv_w;
        scale_from_amax(__ldg(p.in_amax), s_in, inv_in);
        scale_from_amax(__ldg(p.w_amax), s_w, inv_w);
        const float s_ab = s_in * s_w;
        int acc = 0;
        uint32_t pacc = 0;
        float y_max = 0.f;
        for (int k = 0; k < n_my; ++k) {
            const int t = t_begin + k;
            const int row = __ldg(p.tile_rows + (long long)t * TM + q * 32 + lane);
            const bool tre = DBG && p.trace != nullptr && blockIdx.x == 0 && q == 0 && lane == 0 && k < 64;
            if (tre) p.trace[2048 + 4 * k] = clock64();
            wait_bar(acc_full + 8 * acc, pacc);                              // (every lane polls: no diverged warp in front of the .aligned loads)
            __syncwarp();
            fence_after();
            if (tre) p.trace[2048 + 4 * k + 1] = clock64();
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols);
            // the staging row of this thread (AT): the previous tile's bulk store must have read it
            if (AT) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            const uint32_t stage_row = a_base + (uint32_t)(q * 32 + lane) * kOutPitch;
            const int n_sets = NI * p.n_main;                                // (issuer, main accumulator) -> 128 columns [main | cross]
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                if (c0 >= p.c_out) break;
                float sum[16];
                if (AT) {                                                    // one accumulator set: both loads in flight, one wait
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(taddr + 64 + c0, v0);                   // cross
                    tmem_ld16_nowait(taddr + c0, v1);                        // main
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] = fmaf(__uint_as_float(v0[j]), kLoInv, __uint_as_float(v1[j]));
                } else {
                    uint32_t v[16];
                    tmem_ld16(taddr + 64 + c0, v);                           // cross terms
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]);
                    for (int g = 1; g < n_sets; ++g) {
                        tmem_ld16(taddr + g * 128 + 64 + c0, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] *= kLoInv;
                    for (int g = 0; g < n_sets; ++g) {
                        tmem_ld16(taddr + g * 128 + c0, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
                    }
                }
                if (dbg & 16) continue;
                // bias + activation + max|y|: chunk-uniform branches only (a per-element switch / bound test version spent
                // ~5.5k cycles per tile here -- longer than the tensor-memory reads)
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                    y[j] = fmaf(sum[j], s_ab, b.x);
                    y[j + 1] = fmaf(sum[j + 1], s_ab, b.y);
                    y[j + 2] = fmaf(sum[j + 2], s_ab, b.z);
                    y[j + 3] = fmaf(sum[j + 3], s_ab, b.w);
                }
                if (p.act == HPL_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = y[j] > 0.f ? y[j] : 0.f;
                } else if (p.act == HPL_ACT_LEAKY) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = y[j] > 0.f ? y[j] : HPL_LEAKY_RATE * y[j];
                }
                if (row >= 0) {
                    float m = 0.f;
                    if (c0 + 16 <= p.c_out) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) m = fmaxf(m, fabsf(y[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < p.c_out) m = fmaxf(m, fabsf(y[j]));
                    }
                    y_max = fmaxf(y_max, m);
                }
                if (AT) {                                                    // stage the row; it leaves by bulk copy after the release
                    if (dbg & 64) continue;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        sts128(stage_row + (c0 + j) * 4, __float_as_uint(y[j]), __float_as_uint(y[j + 1]), __float_as_uint(y[j + 2]),
                               __float_as_uint(y[j + 3]));
                } else if (row >= 0) {
                    float* dst = p.out + (long long)row * p.ld_out + c0;
                    if (c0 + 15 < p.c_out) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < p.c_out) dst[j] = y[j];
                    }
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(acc_empty + 8 * acc);
            if (tre) p.trace[2048 + 4 * k + 3] = clock64();
            if (AT && !(dbg & (16 | 128))) {                                   // the accumulators are free again; the row leaves asynchronously
                fence_proxy_async();
                if (row >= 0)
                    bulk_store(p.out + (long long)row * p.ld_out, a_base + (uint32_t)(q * 32 + lane) * kOutPitch, (uint32_t)p.c_out * 4u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (tre) p.trace[2048 + 4 * k + 2] = clock64();
            if (++acc == acc_stages) { acc = 0; pacc ^= 1; }
        }
        if (p.out_amax != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) y_max = fmaxf(y_max, __shfl_xor_sync(0xffffffffu, y_max, o));
            if (lane == 0 && y_max > 0.f) atomicMax(p.out_amax, __float_as_uint(y_max));
        }
        if (AT) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last rows have left shared memory
    }
    fence_before();
    __syncthreads();
    if (warp == kC5MmaWarp) {
        fence_after();
        tmem_dealloc(tmem_d, 512);
    }
}

// ======================================================================================== weight gradient
// dw[f, c, o] += sum_v x[nbr[f, v], c] * dz[v, o]   (autograd of models/bilateralNN.py:219) on the same tile plan.
//
// CTA = (32-channel block cb of x, contiguous range of vertex tiles); per tile the distinct x rows are staged once (as in
// conv5_kernel) and the tile's dz rows are loaded as the B operand.  M = (tap, channel): an M tile is 4 taps x 32
// channels, so the 16 (padded) taps of one channel block are 4 M tiles x 128 TMEM columns = all of tensor memory: every
// accumulator stays resident over the CTA's whole vertex range (flushed with fp32 RED every kFlushTiles tiles, which
// keeps the hi.hi accumulator below ~160 accumulate steps).  K = vertices: both operands are MN-major no-swizzle
//   element (m, k) at (m / 8) * SBO + (k / 8) * 128 + (k % 8) * 16 + (m % 8) * 2,   SBO = 8 * 128 + 16 (bank spreading)
// so a staged row chunk (8 channels of one vertex, 16 bytes) is again one LDS.128 / STS.128.
//   MMA 1: x_hi (M128) . [dz_hi | dz_lo] (N128) -> columns [main | cross];   MMA 2: x_lo . dz_hi (N64) -> cross.
// Stage = (K half of 64 vertices, M tile): A hi / lo planes of 16.3 KB each, ring of 2; dz halves: ring of 2.
constexpr uint32_t kW_SBO = 8 * 128 + 16;            // 1040: stride between 8-wide M (or N) groups, K half of 64 vertices
constexpr int kWPlane = 16 * kW_SBO + 64;            // 16704: one A plane (+64: the lo plane lands 4 bank groups further)
constexpr int kWAStage = 2 * kWPlane;
constexpr int kWBHalf = 16 * kW_SBO;                 // 16640: [dz_hi | dz_lo] of 64 vertices
constexpr int kWSmem = 2 * kUBuf + 2 * kWAStage + 2 * kWBHalf + 2 * kIdxBuf + 2 * TM * 4 + 1024;
constexpr int kFlushTiles = 16;                      // 16 tiles x 8 K steps = 128 accumulate steps per accumulator
static_assert(kWSmem <= 232448, "shared memory budget (weight gradient)");

struct Wgrad5Args {
    const uint8_t* x16;           // h16b image of the layer input (gathered operand)
    const uint8_t* dz16;          // h16b image of dz (n_rows, c_out)
    const int* tile_rows;
    const int* n_uniq;
    const int* uniq;
    const unsigned short* local;
    float* dw;                    // (F, C, Co) fp32, accumulated with RED
    const uint32_t* x_amax;
    const uint32_t* dz_amax;
    int n_tiles, cb_count, cbo_count, filter_size, c_in, c_out, tiles_per_cta;
    int dbg;                      // timing experiments (HPL_WGRAD5_DBG): 1 no fence.proxy.async, 4 no MMAs, 8 no operand stores, 16 no row reads, 32 no slot reads
};

__global__ void __launch_bounds__(kThreads, 1) wgrad5_kernel(const Wgrad5Args p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 + 2 + 2 + 2 + 2];
    __shared__ uint32_t tmem_slot;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t u_base = smem_base;
    const uint32_t a_base = u_base + 2 * kUBuf;
    const uint32_t b_base = a_base + 2 * kWAStage;
    const uint32_t idx_base = b_base + 2 * kWBHalf;
    const uint32_t rows_base = idx_base + 2 * kIdxBuf;                           // 2 x 128 output rows (tile_rows)
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full = bar0, empty = full + 16, ufull = empty + 16, bfull = ufull + 16, accb = bfull + 16;   // accb: [0] full, [1] empty

    const int cb = blockIdx.x % p.cb_count;
    const int range = blockIdx.x / p.cb_count;
    const int t_begin = range * p.tiles_per_cta;
    const int t_end = min(p.n_tiles, t_begin + p.tiles_per_cta);
    const int n_my = max(0, t_end - t_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[s], kCopyWarps);
            mbar_init(&bars[2 + s], 1);
            mbar_init(&bars[4 + s], kCopyThreads);
            mbar_init(&bars[6 + s], kCopyThreads);
        }
        mbar_init(&bars[8], 1);
        mbar_init(&bars[9], 4);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x < 16) {
        const uint32_t z = u_base + (threadIdx.x >> 3) * kUBuf + kUmax * kURow + (threadIdx.x & 7) * 16;
        sts128(z, 0, 0, 0, 0);
    }
    // the dz buffers: groups a narrow dz (c_out <= 32) never writes must not hold NaN patterns
    for (int i = threadIdx.x; i < 2 * kWBHalf / 16; i += kThreads) sts128(b_base + i * 16, 0, 0, 0, 0);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp < kCopyWarps) {
        // ------------------------------------------------------------------ copy warps
        const int tid = threadIdx.x;
        const int c8 = tid & 7, rg = tid >> 3;
        const long long row_bytes = (long long)p.cb_count * kURow;
        const long long dz_row_bytes = (long long)p.cbo_count * kURow;
        const uint32_t src_off = c8 * 16;
        const uint32_t idx_off = rg * 2;
        // A position of (row = q * 32 + rg within the K half, chunk c8) for tap tl: + tl * 4 * kW_SBO + q * 512
        const uint32_t dst_off = (c8 >> 2) * kWPlane + (c8 & 3) * kW_SBO + (rg >> 3) * 128 + (rg & 7) * 16;
        const uint8_t* in_thread = p.x16 + cb * kURow + src_off;

        int urow[kUIters];
        auto fetch_rows = [&](int k) {
            const int t = t_begin + k;
            const int n = __ldg(p.n_uniq + t);
            const int* up = p.uniq + (long long)t * kUmax + (tid >> 3);
#pragma unroll
            for (int i = 0; i < kUIters; ++i) urow[i] = (tid >> 3) + i * (kCopyThreads / 8) < n ? __ldg(up + i * (kCopyThreads / 8)) : -1;
        };
        auto load_idx_block = [&](int k) {                               // slots + the tile's output rows
            cp_async16(idx_base + (k & 1) * kIdxBuf + tid * 16,
                       reinterpret_cast<const uint8_t*>(p.local) + ((long long)(t_begin + k) * (kTaps * TM)) * 2 + tid * 16);
            if (tid < TM / 4)
                cp_async16(rows_base + (k & 1) * (TM * 4) + tid * 16,
                           reinterpret_cast<const uint8_t*>(p.tile_rows) + ((long long)(t_begin + k) * TM) * 4 + tid * 16);
        };
        auto load_rows = [&](uint32_t ub, int i0) {
            const uint32_t dst = ub + (tid >> 3) * kURow + src_off;
#pragma unroll
            for (int i = i0; i < i0 + 4; ++i)
                if (i < kUIters && urow[i] >= 0) cp_async16(dst + i * (kCopyThreads / 8) * kURow, in_thread + (long long)urow[i] * row_bytes);
        };
        // dz rows of K half `h` (0 / 1) of tile k into B buffer `bb`: thread = (chunk j of a 64-byte plane segment, row pair
        // kk / kk + 4): the 8 lanes of a quarter warp hit 8 different bank groups.  (The tile's row block must have landed.)
        auto load_dz_half = [&](int k, int h, uint32_t bb) {
            const int j = tid & 3, r4 = (tid >> 2) & 1;
            const int shift = p.cbo_count == 2 ? 2 : 1;                           // (plane, block) pairs per row = 2 * cbo = 1 << shift
            const int n_items = 32 << shift;                                      // (k8, kk < 4) x pairs
            const uint32_t rb = rows_base + (k & 1) * (TM * 4) + h * 64 * 4;
            for (int e = tid >> 3; e < n_items; e += kCopyThreads / 8) {
                const int pair = e & ((1 << shift) - 1), g = e >> shift;          // g = k8 * 4 + kk
                const int plane = pair & 1, b = pair >> 1;
                const int row = (g >> 2) * 8 + (g & 3) + 4 * r4;                  // within the half
                int v;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(rb + row * 4) : "memory");
                const int n8 = plane * 8 + b * 4 + j;
                const uint32_t dst = bb + n8 * kW_SBO + (row >> 3) * 128 + (row & 7) * 16;
                if (v >= 0) cp_async16(dst, p.dz16 + (long long)v * dz_row_bytes + b * kURow + plane * 64 + j * 16);
                else sts128(dst, 0, 0, 0, 0);
            }
        };

        if (n_my > 0) {
            fetch_rows(0);
            load_idx_block(0);
#pragma unroll
            for (int i0 = 0; i0 < kUIters; i0 += 4) load_rows(u_base, i0);
            cp_async_arrive_noinc(ufull);
            wait_bar(ufull, 0);                                                  // (the row block; waiting twice on a phase is fine)
            load_dz_half(0, 0, b_base);
            cp_async_arrive_noinc(bfull);
            load_dz_half(0, 1, b_base + kWBHalf);
            cp_async_arrive_noinc(bfull + 8);
            if (n_my > 1) fetch_rows(1);
        }
        int stage = 0;
        uint32_t phase_bit = 0;
        for (int k = 0; k < n_my; ++k) {
            // (urow holds the row list of tile k + 1: it was fetched at stage 4 of tile k - 1.  Read at the top of the tile
            // that uses it, the cold list -- two dependent DRAM round trips, the count and then the rows -- stalled every copy
            // warp at stage 0 of every tile.)
            const bool has_next = k + 1 < n_my;
            wait_bar(ufull + 8 * (k & 1), (k >> 1) & 1);
            const uint32_t ub = u_base + (k & 1) * kUBuf + src_off;
            const uint32_t ubn = u_base + ((k + 1) & 1) * kUBuf;
            const uint32_t ibs = idx_base + (k & 1) * kIdxBuf + idx_off;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const int s = h * 4 + mt;                                     // stage number inside the tile
                    const uint32_t abd = a_base + stage * kWAStage + dst_off;
                    uint32_t slot[8];
                    uint4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) slot[i] = (p.dbg & 32) ? (uint32_t)i : lds_u16(ibs + ((4 * mt + (i >> 1)) * TM + h * 64 + (i & 1) * 32) * 2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (p.dbg & 16) ? make_uint4(slot[i], 0, 0, 0) : lds128(ub + slot[i] * kURow);
                    // MMAs of stage s - 2 are complete once the stage buffer is free.  s == 1: the previous tile's last stage
                    // (h = 1, mt = 3) is done -> the dz half-1 buffer may be refilled; s == 5: half 0 of this tile is done.
                    wait_bar(empty + 8 * stage, phase_bit ^ 1);                  // (every lane polls: a lane-0 wait leaves the warp diverged)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (!(p.dbg & 8)) sts128(abd + (i >> 1) * (4 * kW_SBO) + (i & 1) * 512, v[i].x, v[i].y, v[i].z, v[i].w);
                    if (has_next) {
                        if (s < 4) {                                              // next tile's x rows
                            if (s == 0) load_idx_block(k + 1);
                            load_rows(ubn, 4 * s);
                            if (s == 3) cp_async_arrive_noinc(ufull + 8 * ((k + 1) & 1));
                        }
                        if (s == 4 && k + 2 < n_my) fetch_rows(k + 2);           // (urow is free again: its last use was stage 3)
                        if (s == 5) {                                             // stage 3 (last user of dz half 0) has completed
                            wait_bar(ufull + 8 * ((k + 1) & 1), ((k + 1) >> 1) & 1);  // next tile's row block (issued at stage 0)
                            load_dz_half(k + 1, 0, b_base);
                            cp_async_arrive_noinc(bfull);
                        }
                    }
                    if (s == 1 && k > 0) {                        // previous tile's stage 7 (last user of half 1) has completed
                        load_dz_half(k, 1, b_base + kWBHalf);
                        cp_async_arrive_noinc(bfull + 8);
                    }
                    if (mt == 0) wait_bar(bfull + 8 * h, k & 1);                  // this half's dz rows have landed
                    if (!(p.dbg & 1)) fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(full + 8 * stage);
                    if (++stage == 2) { stage = 0; phase_bit ^= 1; }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kCopyThreads) : "memory");
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
        {
            constexpr uint32_t kIdescMain = instr_desc(0, TM, 128, 1, 1);
            constexpr uint32_t kIdescLo = instr_desc(0, TM, 64, 1, 1);
            constexpr uint64_t kDesc = (uint64_t)(128 >> 4) << 16 | (uint64_t)((kW_SBO >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
            int stage = 0, since_flush = 0;
            uint32_t phase_bit = 0, pacc = 0;
            for (int k = 0; k < n_my; ++k) {
                if (since_flush == 0 && k > 0) {                                  // accumulators were flushed: wait until they are drained
                    wait_bar(accb + 8, pacc);
                    pacc ^= 1;
                    fence_after();
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t b16 = (b_base + h * kWBHalf) >> 4;
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        wait_bar(full + 8 * stage, phase_bit);
                        fence_after();
                        const uint32_t a16 = (a_base + stage * kWAStage) >> 4;
                        const uint32_t d = tmem_d + (uint32_t)(mt * 128);
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (p.dbg & 4) break;
                                const uint32_t acc = (since_flush | h | j) != 0;
                                umma_f16(d, kDesc | (a16 + j * 16), kDesc | (b16 + j * 16), kIdescMain, acc);
                                umma_f16(d + 64, kDesc | (a16 + (kWPlane >> 4) + j * 16), kDesc | (b16 + j * 16), kIdescLo, 1);
                            }
                            umma_commit_a(empty + 8 * stage);
                        }
                        if (++stage == 2) { stage = 0; phase_bit ^= 1; }
                    }
                }
                if (++since_flush == kFlushTiles || k == n_my - 1) {
                    if (elect_one()) umma_commit_a(accb);
                    since_flush = 0;
                }
            }
        }
    } else if (warp >= 10) {
        // ------------------------------------------------------------------ flush: TMEM -> RED into dw
        const int q = warp & 3;
        float s_x, inv_x, s_z, inv_z;
        scale_from_amax(__ldg(p.x_amax), s_x, inv_x);
        scale_from_amax(__ldg(p.dz_amax), s_z, inv_z);
        const float s_ab = s_x * s_z;
        const int n_flush = (n_my + kFlushTiles - 1) / kFlushTiles;
        const int m = q * 32 + lane;                                              // M row of the tile: tap_l * 32 + channel
        const int ch = cb * 32 + (m & 31);
        for (int fl = 0; fl < n_flush; ++fl) {
            wait_bar(accb, fl & 1);
            __syncwarp();
            fence_after();
#pragma unroll 1
            for (int mt = 0; mt < 4; ++mt) {
                const int f = 4 * mt + (m >> 5);
                const bool live = f < p.filter_size && ch < p.c_in;
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 128);
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    if (c0 >= p.c_out) break;
                    uint32_t vm[16], vc[16];
                    tmem_ld16(taddr + c0, vm);
                    tmem_ld16(taddr + 64 + c0, vc);
                    if (live) {
                        float* dst = p.dw + ((long long)f * p.c_in + ch) * p.c_out + c0;
                        float y[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) y[j] = fmaf(__uint_as_float(vc[j]), kLoInv, __uint_as_float(vm[j])) * s_ab;
                        if (c0 + 15 < p.c_out && (p.c_out & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) red_add_f32x4(dst + j, make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < p.c_out) atomicAdd(dst + j, y[j]);
                        }
                    }
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(accb + 8);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after();
        tmem_dealloc(tmem_d, 512);
    }
}

void set_attrs() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(conv5_kernel<15, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<16, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<15, true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<16, true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<15, false, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<15, true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<15, true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<16, true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(conv5_kernel<15, true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(wgrad5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem);
    done = true;
}

int cb_of(int64_t c) { return (int)((c + 31) / 32); }

}  // namespace

extern "C" {

int64_t hpl_h16b_bytes(int64_t n_rows, int64_t channels) { return n_rows * cb_of(channels) * kURow; }

int hpl_h16b_split(const float* x, int64_t ld, int64_t n_rows, int64_t channels, const float* norm, const uint32_t* amax,
                   void* x16, void* stream) {
    HPL_CHECK_ARG((x || n_rows == 0) && amax && x16 && channels > 0 && ld >= channels && ld % 4 == 0);
    HPL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)x16 & 15) == 0);
    if (n_rows == 0) return 0;
    const int cb = cb_of(channels);
    const long long work = n_rows * cb * 4;
    h16b_split_kernel<<<(unsigned)((work + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld, n_rows, (int)channels, cb, norm, amax,
                                                                                  reinterpret_cast<uint4*>(x16));
    HPL_RETURN_LAST();
}

int hpl_h16b_split_ex(float* x, int64_t ld, int64_t n_rows, int64_t channels, const float* norm, float* inv_out,
                      uint32_t* norm_amax_out, const float* y, int64_t ld_y, int act, const uint32_t* amax_a,
                      const uint32_t* amax_b, uint32_t* amax_out, float* colsum, int dispose, void* x16, void* stream) {
    HPL_CHECK_ARG((x || n_rows == 0) && amax_a && x16 && channels > 0 && ld >= channels && ld % 4 == 0);
    HPL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)x16 & 15) == 0 && dispose >= 0 && dispose <= 2);
    HPL_CHECK_ARG(act == HPL_ACT_NONE || (y != nullptr && ld_y >= channels));
    HPL_CHECK_ARG(cb_of(channels) * 4 <= 256);
    if (n_rows == 0) return 0;
    const int cb = cb_of(channels);
    const int rows_per_iter = 256 / (cb * 4);
    long long blocks = (n_rows + rows_per_iter - 1) / rows_per_iter;
    const long long cap = 8LL * num_sms();                                // grid-stride: few atomics per column sum
    if (blocks > cap) blocks = cap;
    const float slope = act == HPL_ACT_LEAKY ? HPL_LEAKY_RATE : 0.f;
    h16b_split_fused_kernel<false><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        x, ld, n_rows, (int)channels, cb, norm, inv_out, norm_amax_out, act == HPL_ACT_NONE ? nullptr : y, ld_y, slope, amax_a, amax_b,
        amax_out, colsum, dispose, reinterpret_cast<uint4*>(x16), CsrArgs{});
    HPL_RETURN_LAST();
}

/* The splat (bilateralNN.py:150-182) and the slice backward (:226-232 autograd) as a deterministic GATHER fused with the operand split:
 *   row v of the image <- split( post( sum_{e in [csr_ptr[v], csr_ptr[v+1])} bary[r(e), pt(e)] * src[pt(e), :] ) )
 * csr_ent[e] = pt | r << 30 (contributions sorted by vertex: the splat plan), src (n_points, ld_src) point-major.
 * post: normalize != 0 -> divide by (the row's weight sum + 1e-5), inv_out[v] <- that reciprocal, norm_amax_out <- max weight sum;
 *       y / act -> multiply by act'(y[v, :]); colsum += column sums; scale from *amax_a (x *amax_b), amax_out as hpl_h16b_split_ex. */
int hpl_h16b_splat_csr(const float* src, int64_t ld_src, const float* bary, int64_t n_points, const int32_t* csr_ptr,
                       const int32_t* csr_ent, int64_t n_rows, int64_t channels, int normalize, float* inv_out,
                       uint32_t* norm_amax_out, const float* y, int64_t ld_y, int act, const uint32_t* amax_a, const uint32_t* amax_b,
                       uint32_t* amax_out, float* colsum, void* x16, void* stream) {
    HPL_CHECK_ARG((src && bary && csr_ptr && csr_ent) || n_rows == 0);
    HPL_CHECK_ARG(amax_a && x16 && channels > 0 && ld_src >= channels && ld_src % 4 == 0 && n_points < (1LL << 30));
    HPL_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)x16 & 15) == 0);
    HPL_CHECK_ARG(act == HPL_ACT_NONE || (y != nullptr && ld_y >= channels));
    HPL_CHECK_ARG(cb_of(channels) * 4 <= 256);
    if (n_rows == 0) return 0;
    const int cb = cb_of(channels);
    const int rows_per_iter = 256 / (cb * 4);
    long long blocks = (n_rows + rows_per_iter - 1) / rows_per_iter;
    const long long cap = 16LL * num_sms();                               // grid-stride: few atomics per column sum
    if (blocks > cap) blocks = cap;
    const float slope = act == HPL_ACT_LEAKY ? HPL_LEAKY_RATE : 0.f;
    CsrArgs c{src, ld_src, bary, n_points, csr_ptr, csr_ent, normalize};
    h16b_split_fused_kernel<true><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        nullptr, 0, n_rows, (int)channels, cb, nullptr, inv_out, norm_amax_out, act == HPL_ACT_NONE ? nullptr : y, ld_y, slope, amax_a,
        amax_b, amax_out, colsum, 0, reinterpret_cast<uint4*>(x16), c);
    HPL_RETURN_LAST();
}

int64_t hpl_conv5_workspace(int64_t c_in) { return (int64_t)cb_of(c_in) * kTaps * kBTap + 16; }
int hpl_conv5_supported(int64_t filter_size, int64_t c_in, int64_t c_out) {
    const int64_t steps = filter_size * cb_of(c_in) * 2;
    return filter_size >= 1 && filter_size <= kTaps && c_out >= 1 && c_out <= 64 && c_in >= 1 && steps <= 4 * 160;
}

/* out[row, :] = act(bias + sum_f x[nbr[f, row]] . w[f]) for every row of the plan's table; x16 = h16b image of x.
 * w: element (f, c, o) at w + f * w_sf + c * w_sc + o * w_so.  tap_map (device, F ints) or NULL.
 * workspace: hpl_conv5_workspace(c_in) bytes, 128-byte aligned; workspace_valid != 0: its weight image is current. */
/* Weight image of hpl_conv5 built ahead of the call (e.g. on a side stream while the splat runs): max|w| + the tile image
 * into `workspace`; hpl_conv5 is then called with workspace_valid = 1. */
int hpl_conv5_weights(const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so, int64_t filter_size, int64_t c_in, int64_t c_out,
                      const int32_t* tap_map, void* workspace, void* stream) {
    HPL_CHECK_ARG(w && workspace && ((uintptr_t)workspace & 127) == 0);
    HPL_CHECK_ARG(hpl_conv5_supported(filter_size, c_in, c_out));
    const int cb = cb_of(c_in);
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* w_amax = reinterpret_cast<uint32_t*>(image + (long long)cb * kTaps * kBTap);
    const int rc = hpl_absmax(w, filter_size * c_in * c_out, w_amax, stream);
    if (rc != 0) return rc;
    const long long total = (long long)cb * kTaps * 64 * 4;
    weight_image5_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(w, w_sf, w_sc, w_so, (int)filter_size, (int)c_in,
                                                                                         (int)c_out, cb, tap_map, w_amax, image);
    HPL_RETURN_LAST();
}

int hpl_conv5(const void* x16, const void* plan, int64_t n_out_rows, int64_t filter_size, int64_t c_in, int64_t c_out,
              const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so, const int32_t* tap_map, const float* bias, int act,
              float* out, int64_t ld_out, void* workspace, int workspace_valid, const uint32_t* in_amax, uint32_t* out_amax,
              void* stream) {
    HPL_CHECK_ARG(x16 && plan && w && out && workspace && in_amax);
    HPL_CHECK_ARG(hpl_conv5_supported(filter_size, c_in, c_out));
    HPL_CHECK_ARG(((uintptr_t)x16 & 15) == 0 && ((uintptr_t)workspace & 127) == 0 && ((uintptr_t)plan & 255) == 0);
    HPL_CHECK_ARG(ld_out >= c_out && ld_out % 4 == 0 && ((uintptr_t)out & 15) == 0);
    if (n_out_rows == 0) return 0;
    cudaStream_t s = as_stream(stream);
    set_attrs();
    const int cb = cb_of(c_in);
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* w_amax = reinterpret_cast<uint32_t*>(image + (long long)cb * kTaps * kBTap);
    if (!workspace_valid) {
        // max|w| over the (possibly strided) weight: the buffer behind a dense permutation holds exactly the F*C*Co values
        const int rc = hpl_conv5_weights(w, w_sf, w_sc, w_so, filter_size, c_in, c_out, tap_map, workspace, stream);
        if (rc != 0) return rc;
    }
    const uint8_t* pb = reinterpret_cast<const uint8_t*>(plan);
    Conv5Args a;
    a.in16 = reinterpret_cast<const uint8_t*>(x16);
    a.w_image = image;
    a.tile_rows = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 0));
    a.n_uniq = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 1));
    a.uniq = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 2));
    a.local = reinterpret_cast<const unsigned short*>(pb + hpl_plan_offset(n_out_rows, 3));
    a.bias = bias; a.out = out; a.in_amax = in_amax; a.w_amax = w_amax; a.out_amax = out_amax;
    a.ld_out = ld_out;
    a.n_tiles = (int)hpl_plan_tiles(n_out_rows);
    a.cb_count = cb; a.filter_size = (int)filter_size; a.c_out = (int)c_out; a.act = act;
    a.steps_total = (int)((filter_size == 15 ? 15 : 16) * cb * 2);
    a.n_main = (a.steps_total + 319) / 320;                  // <= ~160 accumulate steps per accumulator, two issuers share the K steps
    { const char* e = getenv("HPL_CONV5_DBG"); a.dbg = e ? atoi(e) : 0; }
    a.trace = nullptr;
    { const char* e = getenv("HPL_CONV5_TRACE"); if (e) a.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0)); }
    const unsigned grid = (unsigned)(a.n_tiles < num_sms() ? a.n_tiles : num_sms());
    // A operand: HPL_CONV5_TMEM=0 shared memory; 1 tensor memory, one tap per stage; 2 (default) tensor memory, two taps per stage
    static int at_knob = -1;
    if (at_knob < 0) { const char* e = getenv("HPL_CONV5_TMEM"); at_knob = e ? atoi(e) : 2; }
    const bool rows_ok = ld_out % 4 == 0 && c_out % 4 == 0;   // (bulk row stores: 16-byte multiples)
    const bool at = at_knob && a.n_main == 1 && rows_ok;
    const bool hooks = (a.dbg != 0 || a.trace != nullptr) && filter_size == 15;   // experiment hooks (tools/try_conv5_*.py)
    if (at && at_knob >= 2) {
        if (hooks) conv5_kernel<15, true, true, 2><<<grid, kC5Threads, kSmem, s>>>(a);
        else if (filter_size == 15) conv5_kernel<15, true, false, 2><<<grid, kC5Threads, kSmem, s>>>(a);
        else conv5_kernel<16, true, false, 2><<<grid, kC5Threads, kSmem, s>>>(a);
    } else if (hooks) {
        if (at) conv5_kernel<15, true, true, 1><<<grid, kC5Threads, kSmem, s>>>(a);
        else conv5_kernel<15, false, true, 1><<<grid, kC5Threads, kSmem, s>>>(a);
    } else if (at) {
        if (filter_size == 15) conv5_kernel<15, true, false, 1><<<grid, kC5Threads, kSmem, s>>>(a);
        else conv5_kernel<16, true, false, 1><<<grid, kC5Threads, kSmem, s>>>(a);
    } else {
        if (filter_size == 15) conv5_kernel<15, false, false, 1><<<grid, kC5Threads, kSmem, s>>>(a);
        else conv5_kernel<16, false, false, 1><<<grid, kC5Threads, kSmem, s>>>(a);
    }
    HPL_RETURN_LAST();
}

/* dw[f, c, o] += sum_v x[nbr[f, v], c] * dz[v, o] over the plan's table (dw (F, C, Co) fp32, zeroed by the caller).
 * x16 / dz16: h16b images of x (n_in_rows, c_in) and dz (n_out_rows, c_out); c_out <= 64. */
int hpl_wgrad5(const void* x16, const void* dz16, const void* plan, int64_t n_out_rows, int64_t filter_size, int64_t c_in,
               int64_t c_out, float* dw, const uint32_t* x_amax, const uint32_t* dz_amax, void* stream) {
    HPL_CHECK_ARG(x16 && dz16 && plan && dw && x_amax && dz_amax);
    HPL_CHECK_ARG(filter_size >= 1 && filter_size <= kTaps && c_in >= 1 && c_out >= 1 && c_out <= 64);
    HPL_CHECK_ARG(((uintptr_t)x16 & 15) == 0 && ((uintptr_t)dz16 & 15) == 0 && ((uintptr_t)plan & 255) == 0 && ((uintptr_t)dw & 15) == 0);
    if (n_out_rows == 0) return 0;
    set_attrs();
    const uint8_t* pb = reinterpret_cast<const uint8_t*>(plan);
    Wgrad5Args a;
    a.x16 = reinterpret_cast<const uint8_t*>(x16);
    a.dz16 = reinterpret_cast<const uint8_t*>(dz16);
    a.tile_rows = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 0));
    a.n_uniq = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 1));
    a.uniq = reinterpret_cast<const int*>(pb + hpl_plan_offset(n_out_rows, 2));
    a.local = reinterpret_cast<const unsigned short*>(pb + hpl_plan_offset(n_out_rows, 3));
    a.dw = dw; a.x_amax = x_amax; a.dz_amax = dz_amax;
    a.n_tiles = (int)hpl_plan_tiles(n_out_rows);
    a.cb_count = cb_of(c_in); a.cbo_count = cb_of(c_out);
    a.filter_size = (int)filter_size; a.c_in = (int)c_in; a.c_out = (int)c_out;
    { const char* e = getenv("HPL_WGRAD5_DBG"); a.dbg = e ? atoi(e) : 0; }
    int ranges = num_sms() / a.cb_count;
    if (ranges < 1) ranges = 1;
    if (ranges > a.n_tiles) ranges = a.n_tiles;
    a.tiles_per_cta = (a.n_tiles + ranges - 1) / ranges;
    ranges = (a.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
    wgrad5_kernel<<<(unsigned)(ranges * a.cb_count), kThreads, kWSmem, as_stream(stream)>>>(a);
    HPL_RETURN_LAST();
}

}  // extern "C"
