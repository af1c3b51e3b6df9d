// Shared device/host helpers for the sm_100a kernels of the BCL path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hplflownet_b200.h"

#define HPL_LEAKY_RATE 0.1f  // models/module_utils.py:6

#define HPL_CHECK_ARG(cond) \
    do {                    \
        if (!(cond)) return HPL_EINVAL; \
    } while (0)

#define HPL_RETURN_LAST()                      \
    do {                                       \
        cudaError_t e__ = cudaGetLastError();  \
        return e__ == cudaSuccess ? 0 : (int)e__; \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Index tables cross the ABI as int64 (reference format) or int32 (native).
template <bool I64>
struct IdxT;
template <>
struct IdxT<true> { using type = long long; };
template <>
struct IdxT<false> { using type = int; };

template <bool I64>
__device__ __forceinline__ int load_idx(const void* base, long long i) {
    using T = typename IdxT<I64>::type;
    return (int)__ldg(reinterpret_cast<const T*>(base) + i);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == HPL_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == HPL_ACT_LEAKY) return v > 0.f ? v : HPL_LEAKY_RATE * v;
    return v;
}

// 16-byte fp32 reduction to global memory (RED.E.ADD.F32x4 on sm_90+).
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

static inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
