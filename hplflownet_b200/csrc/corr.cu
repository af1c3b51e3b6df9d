// Patch-correlation gather / scatter of BilateralCorrelationFlex (models/bnn_flow.py:170-205).
//
// The reference materialises combined_input (B, 2C+C', F, P, H1) -- 172.8 KB per vertex at
// C = C' = 64, F = P = 15 (bnn_flow.py:199) -- and runs Conv3d(1,P,1) over it; the cloud-1 half is
// identical for all F displacements (:192).  The first conv layer is linear before its
// activation, so it factors per source vertex u and patch slot p:
//     T1[u, p, :] = W_a[:, :, p] . S1[u, :]        T2[u, p, :] = W_b[:, :, p] . S2[u, :]
// (two dense GEMMs, hpl_blur_gemm with nbr = NULL), and the pre-activation of output (v, f) is
//     b + sum_p T1[i1[p,v], p, :] + sum_p T2[i2[f,p,v], p, :]
// -- a pure gather-sum of 128-byte vectors, done here.  FLOPs per vertex drop ~16x and nothing of
// size F*P*C is ever stored.
#include "common.cuh"

namespace {

// z[(v*F + f), :] = act(bias + sum_p t1[i1[p,v], p*O : (p+1)*O] + sum_p t2[i2[f,p,v], p*O : ...])
template <bool I64>
__global__ void corr_gather_kernel(const float* __restrict__ t1, long long ld1, const void* __restrict__ i1,
                                   const float* __restrict__ t2, long long ld2, const void* __restrict__ i2,
                                   const float* __restrict__ bias, int act, float* __restrict__ z, long long ldz,
                                   int quads, int width, int patch, int filt, long long h1, long long n1, long long n2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = t / quads;            // v * F + f
    const int q = (int)(t - row * quads);
    if (row >= h1 * filt) return;
    const long long v = row / filt;
    const int f = (int)(row - v * filt);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) acc = __ldg(reinterpret_cast<const float4*>(bias) + q);
    for (int p = 0; p < patch; ++p) {
        const int u1 = load_idx<I64>(i1, (long long)p * h1 + v);
        const int u2 = load_idx<I64>(i2, ((long long)f * patch + p) * h1 + v);
        if (u1 >= 0 && u1 < n1) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(t1 + (long long)u1 * ld1 + p * width) + q);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
        if (u2 >= 0 && u2 < n2) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(t2 + (long long)u2 * ld2 + p * width) + q);
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
    }
    acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act);
    acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act);
    *(reinterpret_cast<float4*>(z + row * ldz) + q) = acc;
}

// backward of the gather: dt2[i2[f,p,v], p, :] += dz[(v*F+f), :];  dt1[i1[p,v], p, :] += sum_f dz[(v*F+f), :]
template <bool I64>
__global__ void corr_scatter_kernel(const float* __restrict__ dz, long long ldz, const void* __restrict__ i1,
                                    const void* __restrict__ i2, float* __restrict__ dt1, long long ld1,
                                    float* __restrict__ dt2, long long ld2, int quads, int width, int patch,
                                    int filt, long long h1, long long n1, long long n2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long v = t / quads;
    const int q = (int)(t - v * quads);
    if (v >= h1) return;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = 0; f < filt; ++f) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dz + (v * filt + f) * ldz) + q);
        sum.x += g.x; sum.y += g.y; sum.z += g.z; sum.w += g.w;
        for (int p = 0; p < patch; ++p) {
            const int u2 = load_idx<I64>(i2, ((long long)f * patch + p) * h1 + v);
            if (u2 >= 0 && u2 < n2) red_add_f32x4(dt2 + (long long)u2 * ld2 + p * width + 4 * q, g);
        }
    }
    for (int p = 0; p < patch; ++p) {
        const int u1 = load_idx<I64>(i1, (long long)p * h1 + v);
        if (u1 >= 0 && u1 < n1) red_add_f32x4(dt1 + (long long)u1 * ld1 + p * width + 4 * q, sum);
    }
}

}  // namespace

extern "C" {

int hpl_corr_gather(const float* t1, int64_t ld1, const void* i1, const float* t2, int64_t ld2, const void* i2,
                    int idx64, const float* bias, int act, float* z, int64_t ldz, int64_t width, int64_t patch,
                    int64_t filt, int64_t h1, int64_t n1, int64_t n2, void* stream) {
    HPL_CHECK_ARG(t1 && t2 && i1 && i2 && z && width > 0 && width % 4 == 0 && patch > 0 && filt > 0);
    HPL_CHECK_ARG(ld1 >= patch * width && ld2 >= patch * width && ldz >= width);
    HPL_CHECK_ARG(ld1 % 4 == 0 && ld2 % 4 == 0 && ldz % 4 == 0);
    HPL_CHECK_ARG((((uintptr_t)t1 | (uintptr_t)t2 | (uintptr_t)z | (uintptr_t)bias) & 15) == 0);
    if (h1 == 0) return 0;
    const int quads = (int)(width / 4);
    const long long work = h1 * filt * quads;
    const unsigned grid = (unsigned)((work + 255) / 256);
    if (idx64)
        corr_gather_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(t1, ld1, i1, t2, ld2, i2, bias, act, z, ldz, quads, (int)width, (int)patch, (int)filt, h1, n1, n2);
    else
        corr_gather_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(t1, ld1, i1, t2, ld2, i2, bias, act, z, ldz, quads, (int)width, (int)patch, (int)filt, h1, n1, n2);
    HPL_RETURN_LAST();
}

int hpl_corr_scatter(const float* dz, int64_t ldz, const void* i1, const void* i2, int idx64, float* dt1,
                     int64_t ld1, float* dt2, int64_t ld2, int64_t width, int64_t patch, int64_t filt, int64_t h1,
                     int64_t n1, int64_t n2, void* stream) {
    HPL_CHECK_ARG(dz && i1 && i2 && dt1 && dt2 && width > 0 && width % 4 == 0 && patch > 0 && filt > 0);
    HPL_CHECK_ARG(ld1 >= patch * width && ld2 >= patch * width && ldz >= width);
    HPL_CHECK_ARG(ld1 % 4 == 0 && ld2 % 4 == 0 && ldz % 4 == 0);
    HPL_CHECK_ARG((((uintptr_t)dt1 | (uintptr_t)dt2 | (uintptr_t)dz) & 15) == 0);
    if (h1 == 0) return 0;
    const int quads = (int)(width / 4);
    const long long work = h1 * quads;
    const unsigned grid = (unsigned)((work + 127) / 128);
    if (idx64)
        corr_scatter_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(dz, ldz, i1, i2, dt1, ld1, dt2, ld2, quads, (int)width, (int)patch, (int)filt, h1, n1, n2);
    else
        corr_scatter_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(dz, ldz, i1, i2, dt1, ld1, dt2, ld2, quads, (int)width, (int)patch, (int)filt, h1, n1, n2);
    HPL_RETURN_LAST();
}

}  // extern "C"
