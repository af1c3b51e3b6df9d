// Probe: how does tcgen05.mma (kind::tf32, no swizzle) address an MN-major A operand?
// smem word i holds float(i); B (K-major, known-good layout) = 8x8 identity in its first 8 rows, so
// D[m][n] (n < 8) = the smem word index the hardware read for A element (m, k = n).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

__global__ void probe(uint32_t lbo, uint32_t sbo, int a_mn, uint32_t layout, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* A = reinterpret_cast<float*>(smem);              // 16 KB region, word i = i
    float* B = reinterpret_cast<float*>(smem + 32768);      // K-major N=64 x K=8: (k/4)*1024 + (n/8)*128 + (n%8)*16 + (k%4)*4
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) A[i] = (float)(i % 2048);   // exact in tf32 (< 2^11)
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) B[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 8) {
        const int n = threadIdx.x, k = threadIdx.x;
        B[((k / 4) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4] = 1.0f;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        if (a_mn) idesc |= (1u << 15);
        const uint64_t da = smem_desc(smem_u32(A), lbo, sbo, layout);
        const uint64_t db = smem_desc(smem_u32(B), 1024, 128);
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x < 128) {
        asm volatile(
            "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) out[threadIdx.x * 8 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

#include <cuda_fp16.h>
__global__ void probe16(uint32_t lbo, uint32_t sbo, int a_mn, uint32_t layout, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __half* A = reinterpret_cast<__half*>(smem);              // halfword i = i % 2048
    __half* B = reinterpret_cast<__half*>(smem + 32768);      // K-major N=64 x K=16: (k/8)*1024 + (n/8)*128 + (n%8)*16 + (k%8)*2 bytes
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) A[i] = __float2half((float)(i % 2048));
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) B[i] = __float2half(0.f);
    __syncthreads();
    if (threadIdx.x < 16) {
        const int n = threadIdx.x, k = threadIdx.x;
        B[((k / 8) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2] = __float2half(1.0f);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32
        if (a_mn) idesc |= (1u << 15);
        const uint64_t da = smem_desc(smem_u32(A), lbo, sbo, layout);
        const uint64_t db = smem_desc(smem_u32(B), 1024, 128);
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x < 128) {
        asm volatile(
            "{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 128 * 8 * 4);
    float h[1024];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    // {LBO, SBO, a_mn, layout_type}: 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B
    const uint32_t cfgs[][4] = {{2048, 128, 0, 0}, {512, 2048, 1, 1}, {2048, 512, 1, 1}, {1024, 4096, 1, 2}, {4096, 1024, 1, 2},
                                {256, 1024, 1, 6}, {512, 2048, 1, 4}};
    for (auto& c : cfgs) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 40960>>>(c[0], c[1], (int)c[2], c[3], d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("LBO=%u SBO=%u a_mn=%u layout=%u err=%d\n", c[0], c[1], c[2], c[3], (int)e);
        for (int m : {0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 31, 32, 33, 64, 127}) {
            printf("  m=%3d: word idx for k=0..7:", m);
            for (int k = 0; k < 8; ++k) printf(" %5.0f", h[m * 8 + k]);
            printf("\n");
        }
    }
    printf("\n==== fp16 (kind::f16), halfword index read for A element (m, k), k = 0..15 ====\n");
    float* d2; cudaMalloc(&d2, 128 * 16 * 4);
    static float h2[2048];
    cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    // K-major no-swizzle sanity, then MN-major: no swizzle, 128B, 64B, 32B
    const uint32_t c16[][4] = {{2048, 128, 0, 0}, {2048, 128, 1, 0}, {128, 2048, 1, 0}, {1024, 4096, 1, 2}, {4096, 1024, 1, 2},
                               {512, 2048, 1, 4}, {256, 1024, 1, 6}};
    for (auto& c : c16) {
        cudaMemset(d2, 0, sizeof(h2));
        probe16<<<1, 128, 40960>>>(c[0], c[1], (int)c[2], c[3], d2);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
        printf("LBO=%u SBO=%u a_mn=%u layout=%u err=%d\n", c[0], c[1], c[2], c[3], (int)e);
        for (int m : {0, 1, 2, 7, 8, 9, 16, 63, 64, 65, 127}) {
            printf("  m=%3d:", m);
            for (int k = 0; k < 16; ++k) printf(" %5.0f", h2[m * 16 + k]);
            printf("\n");
        }
    }
    return 0;
}
