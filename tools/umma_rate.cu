// GPU experiment: steady-state cycles per tcgen05.mma (kind::f16, M = 128, K = 16, one CTA per SM) for the operand
// sources / layouts the contraction kernels use.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hplflownet_b200/csrc -I include tools/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace tc;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d),
        "r"(a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

// mode 0: SS N=128   1: TS N=128   2: TS N=64   3: conv5 mix (TS: N128 -> d, N64 -> d + 64, x2)   4: SS N=64
// 5: TS N=128 alternating two accumulators   6: TS N=256   7: TS N=8   8: TS N=32   9: two warps issue TS N=128 (own accumulators)
// 10: two warps issue TS N=64
template <int mode>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, uint32_t lbo_b, uint32_t sbo_b, uint32_t lbo_a, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = uniform_warp_idx();
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;   // 1.0 halves
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_slot;
    if (warp == 0 || (mode >= 9 && warp == 1)) {
        const uint32_t a_addr = base, b_addr = base + 16 * 1024;
        const uint64_t da = smem_desc(a_addr, lbo_a, 128), db = smem_desc(b_addr, lbo_b, sbo_b);
        constexpr uint32_t i128 = instr_desc(0, 128, 128, 0, 0), i64 = instr_desc(0, 128, 64, 0, 0), i256 = instr_desc(0, 128, 256, 0, 0);
        constexpr uint32_t i8 = instr_desc(0, 128, 8, 0, 0), i32 = instr_desc(0, 128, 32, 0, 0);
        uint64_t* mybar = &bar[warp];
        long long t0 = 0, t1 = 0, t2 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            if (elect_one()) {
#pragma unroll 8
                for (int i = 0; i < n_mma; ++i) {
                    if (mode == 0) umma_f16(tm, da, db, i128, 1);
                    else if (mode == 1) umma_ts(tm, tm + 256, db, i128, 1);
                    else if (mode == 2) umma_ts(tm, tm + 256, db, i64, 1);
                    else if (mode == 3) { if (i & 1) umma_ts(tm + 64, tm + 256 + 16, db, i64, 1); else umma_ts(tm, tm + 256, db, i128, 1); }
                    else if (mode == 4) umma_f16(tm, da, db, i64, 1);
                    else if (mode == 5) umma_ts(tm + (i & 1) * 128, tm + 256, db, i128, 1);
                    else if (mode == 6) umma_ts(tm, tm + 256, db, i256, 1);
                    else if (mode == 7) umma_ts(tm, tm + 256, db, i8, 1);
                    else if (mode == 8) umma_ts(tm, tm + 256, db, i32, 1);
                    else if (mode == 9) umma_ts(tm + warp * 128, tm + 256 + warp * 32, db, i128, 1);
                    else umma_ts(tm + warp * 128, tm + 256 + warp * 32, db, i64, 1);
                }
                umma_commit(mybar);
            }
            t1 = clock64();
            mbar_wait(mybar, rep & 1);
            t2 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) { fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    const char* names[] = {"SS N128", "TS N128", "TS N64", "TS conv5 mix (N128 + N64)", "SS N64", "TS N128 two accumulators", "TS N256", "TS N8", "TS N32",
                           "2 warps x TS N128 (per warp)", "2 warps x TS N64 (per warp)"};
    const int n = 512;
    auto run = [&](auto kern, int mode) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        kern<<<148, 128, 64 * 1024>>>(n, 2048, 128, 2064, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        printf("%-28s: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", names[mode], (double)out[0] / n, (double)out[1] / n);
    };
    run(rate_kernel<0>, 0); run(rate_kernel<1>, 1); run(rate_kernel<2>, 2); run(rate_kernel<3>, 3); run(rate_kernel<4>, 4); run(rate_kernel<5>, 5);
    run(rate_kernel<6>, 6); run(rate_kernel<7>, 7); run(rate_kernel<8>, 8); run(rate_kernel<9>, 9); run(rate_kernel<10>, 10);
    return 0;
}
