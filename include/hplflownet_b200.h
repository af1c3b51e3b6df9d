/*
 * hplflownet_b200.h -- C ABI of the B200-native bilateral-convolution-layer path.
 *
 * Drop-in boundary for the hot path of laoreja/HPLFlowNet:
 *   index half : transforms/transforms.py:264-485 (GenerateDataUnsymmetric) and its only
 *                native dependency, the khash int64 map (models/khash_int2int.h:8-33,
 *                bound through cffi in models/build_khash_cffi.py:15-22 and called from
 *                Numba in transforms/transforms.py:20-24,170-261);
 *   value half : models/bilateralNN.py:9-238 (SparseSum, BilateralConvFlex) and
 *                models/bnn_flow.py:96-210 (BilateralCorrelationFlex), which the
 *                reference runs as stock PyTorch ops.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success, a cudaError_t value (>0) when the CUDA runtime
 *     reports an error, or a negative HPL_E* code for argument errors; nothing is
 *     synchronised -- work is enqueued on `stream`;
 *   - B = 1 as in the reference (README.md:57); several clouds are processed in one call by
 *     concatenating their points/vertices (tables carry global row numbers);
 *   - lattice values live VERTEX-MAJOR in HBM: a (rows, ld) fp32 matrix, one contiguous row
 *     of `ld >= C` floats per lattice vertex, ld % 4 == 0, 16-byte aligned.  There is no
 *     physical "null vertex" row: a table entry of -1 (bilateralNN.py:158-159) reads as zeros;
 *   - point features cross the boundary CHANNEL-MAJOR (C, N) exactly as the reference modules
 *     take them (bilateralNN.py:128-134);
 *   - index tables are int64 (idx64 != 0; what the reference hands over) or int32.
 */
#ifndef HPLFLOWNET_B200_H
#define HPLFLOWNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPL_EINVAL (-1)   /* bad argument (null pointer, misaligned ld, unsupported size) */
#define HPL_ENOSPC (-2)   /* caller-provided capacity too small */

#define HPL_ACT_NONE 0
#define HPL_ACT_RELU 1
#define HPL_ACT_LEAKY 2   /* LeakyReLU(0.1), models/module_utils.py:6 */

/* Library / build identification. */
int hpl_version(void);
/* Compiled SM architecture (100 for sm_100a). */
int hpl_sm_arch(void);

/* ---------------------------------------------------------------- value half */

/* SPLAT (scatter half of SparseSum + splat, bilateralNN.py:9-30,150-166; also the backward of
 * SLICE, :226-232):   rows[off[r,n], c] += bary[r,n] * x[c,n]      r < 4
 * and, when wsum != NULL,   wsum[off[r,n]] += bary[r,n]              (:168-182).
 * x (C, N) channel-major; bary (4, N); off (4, N) in [0, n_rows) -- entries outside are dropped (the reference
 * raises an index error); rows (n_rows, ld) and wsum (n_rows) must be zeroed by the caller (hpl_fill_zero).
 * in_amax (device scalar, may be NULL, zeroed by the caller): receives the bit pattern of max|x| (the operand-scale
 * bound of the splatted, normalised rows: a convex combination never exceeds it).  Accumulation order is not
 * deterministic (fp32 RED). */
int hpl_scatter_rows(const float* x, const float* bary, const void* off, int idx64,
                     int64_t n_points, int64_t channels, float* rows, int64_t ld, int64_t n_rows,
                     float* wsum, uint32_t* in_amax, void* stream);

/* Density normalisation (bilateralNN.py:185-186):  inv[v] = 1/(wsum[v] + 1e-5),
 * rows[v, :] *= inv[v].  inv may alias wsum.  rows == NULL: only the reciprocal is computed (the
 * tensor-core contraction applies it while gathering, see row_scale). */
int hpl_normalize_rows(float* rows, int64_t ld, int64_t n_rows, int64_t channels,
                       const float* wsum, float* inv, void* stream);

/* SLICE (bilateralNN.py:226-236; also the backward of SPLAT, :33-40):
 *   y[c,n] = sum_r bary[r,n] * scale[off[r,n]] * rows[off[r,n], c]  (+ bias[c])
 * scale (H) and bias (C) may be NULL. y (C, N) channel-major.  off entries outside [0, n_rows) read zeros. */
int hpl_gather_rows(const float* rows, int64_t ld, const float* bary, const void* off, int idx64,
                    const float* scale, const float* bias, int64_t n_points, int64_t channels,
                    int64_t n_rows, float* y, void* stream);

/* BLUR / learned convolution over lattice neighbours (bilateralNN.py:198-221 with the conv
 * built at :94-113), as a gather-GEMM with fused bias + activation:
 *   out[v, o] = act( bias[o] + sum_{f<F} sum_{c<C} in[nbr[f,v], c] * w[f, c, o] )
 * in (n_in_rows, ld_in) vertex-major; nbr (F, n_out_rows) in [-1, n_in_rows), -1 reads zeros;
 * nbr == NULL means F == 1 and row v reads row v (the 1x1 layers, :99-100).
 * w (F, C, Co) fp32 -- the reference's (Co, C, F, 1) weight permuted (2,1,0).
 * out (n_out_rows, ld_out) vertex-major, or, when out_channel_major != 0, (Co, ld_out)
 * channel-major with ld_out >= n_out_rows (the module's (B, Co, H) result, :221).
 * The same entry point is the data-gradient of the layer when called with the transposed
 * table (hpl_transpose_table) and w permuted to (F, Co, C).
 * precision: 0 = fp32 FMA on CUDA cores (parity anchor). */
int hpl_blur_gemm(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                  int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                  const float* w, const float* bias, int act, float* out, int64_t ld_out,
                  int out_channel_major, int precision, void* stream);

/* Weight gradient of the layer above:
 *   dw[f, c, o] += sum_v in[nbr[f,v], c] * dz[v, o]        db[o] += sum_v dz[v, o]
 * dw (F, C, Co) and db (Co, may be NULL) must be zeroed by the caller; partial sums over
 * vertex ranges are combined with fp32 RED. */
int hpl_blur_wgrad(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                   int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                   const float* dz, int64_t ld_dz, float* dw, float* db, void* stream);

/* "3xFP16" tensor-core variants of the two contractions (csrc/gemm_tc16.cu, tcgen05 with fp32 accumulation in
 * tensor memory): same contracts, fp32-level accuracy (three MMAs per K step: hi.hi + hi.lo + lo.hi).
 * Operands are split hi + lo*2^-11 in FP16 after scaling by a per-tensor power of two derived from
 * max|x|, which the caller provides as a DEVICE scalar holding the fp32 bit pattern of max|x|:
 *   hpl_absmax(x, count, out_bits)      out_bits <- bits(max_i |x[i]|), x 16-byte aligned
 * (for a vertex-major matrix pass the whole (rows * ld) buffer; pad columns are zero).
 * hpl_blur_gemm_f16: c_in % 4 == 0; workspace = hpl_blur_gemm_f16_workspace(F, C, Co) bytes.
 * hpl_blur_wgrad_f16: c_in % 4 == 0. */
int hpl_absmax(const float* x, int64_t count, uint32_t* out_bits, void* stream);
int64_t hpl_blur_gemm_f16_workspace(int64_t filter_size, int64_t c_in, int64_t c_out);
int hpl_blur_gemm_f16(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                      int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                      const float* w, const float* bias, int act, float* out, int64_t ld_out,
                      int out_channel_major, void* workspace, const uint32_t* in_amax, void* stream);
int hpl_blur_wgrad_f16(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                       int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                       const float* dz, int64_t ld_dz, float* dw, float* db, const uint32_t* in_amax,
                       const uint32_t* dz_amax, void* stream);

/* Fused statistics (one pass instead of three over the same rows).  `amax` slots are device scalars holding the
 * bit pattern of max|x| (as hpl_absmax writes it); the caller zeroes them, the kernels RED.MAX into them.
 *   hpl_blur_gemm_f16_amax: hpl_blur_gemm_f16 whose epilogue also records max|out| (post-activation) in `out_amax`
 *     -- the 3xFP16 scale of the next layer's input (NULL = off) -- and whose weight operand is STRIDED: element
 *     (f, c, o) of the logical (F, C, Co) weight sits at w[f * w_sf + c * w_sc + o * w_so] inside one dense buffer of
 *     F * C * Co floats starting at w, so the reference's (Co, C, F, 1) conv weight (strides 1, F, C * F) and its
 *     transpose for the data gradient are consumed in place, without permuted copies.  All three 0 = contiguous.
 *     workspace_valid != 0: `workspace` still holds the image this function built for the SAME weight values, shape,
 *     strides and the same tile-width class (n_out_rows >= 8192 or not) -- the absmax + image kernels are skipped
 *     (the binding keeps one workspace per parameter and tracks its version counter).
 *   hpl_normalize_rows_amax / hpl_cm_to_rows_amax: the same for the producers of a stack's first input.
 *   hpl_act_backward_stats: dz *= act'(y) (act NONE: dz untouched, y may be NULL), max|dz| -> amax, sum_v dz[v, :] ->
 *     colsum (+=, the convolution's bias gradient); amax / colsum may be NULL. */
int hpl_blur_gemm_f16_amax(const float* in, int64_t ld_in, int64_t n_in_rows, const void* nbr, int idx64,
                           int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                           const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so, const float* bias, int act,
                           float* out, int64_t ld_out, int out_channel_major, void* workspace,
                           int workspace_valid, const uint32_t* in_amax, uint32_t* out_amax, void* stream);
int hpl_normalize_rows_amax(float* rows, int64_t ld, int64_t n_rows, int64_t channels, const float* wsum,
                            float* inv, uint32_t* amax, void* stream);
int hpl_cm_to_rows_amax(const float* cm, int64_t ld_cm, int64_t n, int64_t channels, float* rows, int64_t ld,
                        uint32_t* amax, void* stream);
int hpl_act_backward_stats(float* dz, int64_t ld_dz, const float* y, int64_t ld_y, int64_t n_rows,
                           int64_t channels, int act, uint32_t* amax, float* colsum, void* stream);

/* TMA-gathered variant (csrc/gemm_tma.cu): same contraction as hpl_blur_gemm_f16 (replaces the advanced-index
 * gather + Conv2d of models/bilateralNN.py:198-221), but the gathered operand is moved by the Tensor Memory
 * Accelerator (cp.async.bulk.tensor ... tile::gather4, four lattice rows per instruction, written straight into
 * the SWIZZLE_128B operand layout of tcgen05.mma) by a persistent one-CTA-per-SM kernel with no producer warps.
 *   hpl_h16_split(x, ld, n_rows, C, amax, x16): x16 = (n_rows + 1) rows of [hi(ld16) | lo(ld16)] fp16,
 *     ld16 = round8(C); x / s = hi + lo * 2^-11 with s the power of two derived from `amax` (hpl_absmax); the last
 *     row is zero (missing neighbours are redirected to it).  hpl_h16_bytes gives the size; x16 16-byte aligned.
 *   hpl_blur_gemm_tma: `in16` is that image of the (n_in_rows, c_in) input; everything else as hpl_blur_gemm_f16;
 *     workspace of hpl_blur_gemm_tma_workspace(F, C, Co) bytes, 128-byte aligned. */
int64_t hpl_h16_bytes(int64_t n_rows, int64_t channels);
int hpl_h16_split(const float* x, int64_t ld, int64_t n_rows, int64_t channels, const uint32_t* amax,
                  void* x16, void* stream);
int64_t hpl_blur_gemm_tma_workspace(int64_t filter_size, int64_t c_in, int64_t c_out);
int hpl_blur_gemm_tma(const void* in16, int64_t n_in_rows, const void* nbr, int idx64,
                      int64_t filter_size, int64_t n_out_rows, int64_t c_in, int64_t c_out,
                      const float* w, const float* bias, int act, float* out, int64_t ld_out,
                      int out_channel_major, void* workspace, const uint32_t* in_amax, void* stream);

/* ---- Tile plans + engine 5 (csrc/plan.cu, csrc/gemm_plan.cu): the same contraction as hpl_blur_gemm_f16
 * (models/bilateralNN.py:198-221), but every DISTINCT neighbour row of a 128-vertex tile is loaded once.
 *
 * A plan is a per-table precomputation (one allocation of hpl_plan_bytes(n_rows) bytes, 256-byte aligned; array
 * `which` = 0 tile_rows (n_tiles,128) i32, 1 n_uniq (n_tiles) i32, 2 uniq (n_tiles, umax) i32, 3 local
 * (n_tiles,16,128) u16 starts at hpl_plan_offset(n_rows, which)).  Tiles are runs of 128 entries of `order` (a
 * permutation of the table's columns, NULL = identity); hpl_plan_order derives a spatially coherent order from the
 * table alone by propagating lattice coordinates (coord(nbr[f,v]) = coord(v) + offsets[f], offsets (F,4) int32 =
 * transforms.py:112-130) and sorting along a Morton curve.  stats (4 x int32, device): [0] max distinct rows of a
 * tile, [1] tiles above hpl_plan_umax() (such a plan must not be used: take hpl_blur_gemm_f16), [2] sum of distinct
 * rows over tiles, [3] entries violating nbr[tap_mirror[f], nbr[f,v]] == v (tap_mirror: device, F int32, the tap with the
 * opposite offset; NULL or n_in_rows != n_rows: -1 = not checked).  0 means the data gradient may run on the same plan
 * with mirrored taps instead of a transposed table. */
int64_t hpl_plan_tiles(int64_t n_rows);
int64_t hpl_plan_umax(void);
int64_t hpl_plan_offset(int64_t n_rows, int which);
int64_t hpl_plan_bytes(int64_t n_rows);
int hpl_plan_build(const void* nbr, int idx64, int64_t filter_size, int64_t n_rows, int64_t n_in_rows,
                   const int32_t* order, const int32_t* tap_mirror, void* plan, int32_t* stats, void* stream);
int64_t hpl_plan_order_workspace(int64_t n_rows);
int hpl_plan_order(const void* nbr, int idx64, int64_t filter_size, int64_t n_rows, const int32_t* offsets,
                   int iterations, void* workspace, int32_t* order, int32_t* changed_out, void* stream);

/* "h16b" operand image: per row and 32-channel block one 128-byte line [32 fp16 hi | 32 fp16 lo],
 * x / s = hi + lo * 2^-11, s the power of two derived from *amax (any upper bound of max|x| within 2^10 works).
 * norm != NULL: x[v,:] is first multiplied by 1/(norm[v] + 1e-5) (bilateralNN.py:185-186 fused into the split). */
int64_t hpl_h16b_bytes(int64_t n_rows, int64_t channels);
int hpl_h16b_split(const float* x, int64_t ld, int64_t n_rows, int64_t channels, const float* norm,
                   const uint32_t* amax, void* x16, void* stream);

/* The same split with its neighbouring passes folded in (one trip over the rows).  All optional pointers may be NULL.
 *   norm / inv_out / norm_amax_out : density normalisation as above; inv_out[v] <- 1/(norm[v]+1e-5) (not aliasing norm);
 *                                    norm_amax_out (zeroed slot) <- bit pattern of max norm;
 *   y, ld_y, act                   : x[v,c] *= (y[v,c] > 0 ? 1 : slope(act)) first (activation backward, module_utils.py:34);
 *   amax_a, amax_b, amax_out       : operand scale from *amax_a, or from the bound *amax_a x *amax_b when amax_b != NULL
 *                                    (slice backward: |dz[v]| <= max|g| x sum of barycentric weights at v);
 *                                    amax_out <- the value used, for the contraction kernels that read the image;
 *   colsum                         : += column sums of the (act'-scaled) rows (bias gradient), zeroed by the caller;
 *   dispose                        : 0 keep x, 1 write the act'-scaled rows back, 2 zero x (reuse as a splat accumulator). */
int hpl_h16b_split_ex(float* x, int64_t ld, int64_t n_rows, int64_t channels, const float* norm, float* inv_out,
                      uint32_t* norm_amax_out, const float* y, int64_t ld_y, int act, const uint32_t* amax_a,
                      const uint32_t* amax_b, uint32_t* amax_out, float* colsum, int dispose, void* x16,
                      void* stream);

/* The splat (bilateralNN.py:150-182) -- and the backward of the slice (:226-232) -- as a deterministic GATHER fused with the
 * operand split: no atomics, no fp32 accumulator, every lattice row is written once, straight into the h16b image.
 *   row v <- split( post( sum_{e in [csr_ptr[v], csr_ptr[v+1])} bary[r(e), pt(e)] * src[pt(e), :] ) )
 * csr_ptr (n_rows + 1), csr_ent[e] = pt | r << 30: the splat contributions sorted by vertex (fixed order -> reproducible
 * sums); src (n_points, ld_src): the point features / upstream gradient, point-major (hpl_cm_to_rows).
 * post: normalize != 0 -> divide by (the row's weight sum + 1e-5), inv_out[v] <- that reciprocal, norm_amax_out <- max weight
 * sum; y / act -> activation backward; colsum, amax_a / amax_b / amax_out as in hpl_h16b_split_ex. */
int hpl_h16b_splat_csr(const float* src, int64_t ld_src, const float* bary, int64_t n_points, const int32_t* csr_ptr,
                       const int32_t* csr_ent, int64_t n_rows, int64_t channels, int normalize, float* inv_out,
                       uint32_t* norm_amax_out, const float* y, int64_t ld_y, int act, const uint32_t* amax_a,
                       const uint32_t* amax_b, uint32_t* amax_out, float* colsum, void* x16, void* stream);

/* The weight part of hpl_conv5 alone: max|w| and the tile image of w into `workspace` (hpl_conv5_workspace bytes).  Lets a
 * caller build it ahead of time -- on another stream, while the splat runs -- and call hpl_conv5 with workspace_valid = 1. */
int hpl_conv5_weights(const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so, int64_t filter_size, int64_t c_in,
                      int64_t c_out, const int32_t* tap_map, void* workspace, void* stream);

/* out[row,:] = act(bias + sum_f x[nbr[f,row]] . w[f]) over the plan's table; x16 = h16b image of x (n_in_rows, c_in).
 * w element (f,c,o) at w + f*w_sf + c*w_sc + o*w_so; tap_map (device, F int32, may be NULL): kernel tap g uses
 * w[tap_map[g]] (data gradient: mirrored tap, transposed weight).  workspace: hpl_conv5_workspace(c_in) bytes,
 * 128-byte aligned; workspace_valid != 0: the weight image inside is current.  Supported: hpl_conv5_supported(). */
int64_t hpl_conv5_workspace(int64_t c_in);
int hpl_conv5_supported(int64_t filter_size, int64_t c_in, int64_t c_out);
int hpl_conv5(const void* x16, const void* plan, int64_t n_out_rows, int64_t filter_size, int64_t c_in,
              int64_t c_out, const float* w, int64_t w_sf, int64_t w_sc, int64_t w_so, const int32_t* tap_map,
              const float* bias, int act, float* out, int64_t ld_out, void* workspace, int workspace_valid,
              const uint32_t* in_amax, uint32_t* out_amax, void* stream);

/* Weight gradient on the same plan (autograd of bilateralNN.py:219):
 *   dw[f, c, o] += sum_v x[nbr[f, v], c] * dz[v, o],  dw (F, C, Co) fp32 zeroed by the caller; x16 / dz16 = h16b images
 * of x (n_in_rows, c_in) and dz (n_out_rows, c_out), c_out <= 64. */
int hpl_wgrad5(const void* x16, const void* dz16, const void* plan, int64_t n_out_rows, int64_t filter_size,
               int64_t c_in, int64_t c_out, float* dw, const uint32_t* x_amax, const uint32_t* dz_amax, void* stream);

/* sums[c] += sum_v rows[v, c]  (conv bias gradients). rows (n_rows, ld) vertex-major. */
int hpl_column_sums(const float* rows, int64_t ld, int64_t n_rows, int64_t channels, float* sums,
                    void* stream);

/* Activation backward, in place:  dz[v, c] *= (y[v, c] > 0 ? 1 : slope(act)).
 * LeakyReLU/ReLU keep the sign, so the saved output is enough (module_utils.py:34). */
int hpl_act_backward(float* dz, int64_t ld_dz, const float* y, int64_t ld_y, int64_t n_rows,
                     int64_t channels, int act, void* stream);

/* tbl_t[f, tbl[f, v]] = v for every tbl[f, v] >= 0; tbl_t (F, n_src_rows) int32 must be
 * pre-filled with -1 (hpl_fill_i32).  tbl (F, n_rows).  Used to turn the scatter in the
 * data-gradient of BLUR into a gather.  *collisions (device int32, may be NULL) counts
 * entries that found their slot taken (a non-injective table; never produced by the
 * lattice builder). */
int hpl_transpose_table(const void* tbl, int idx64, int64_t filter_size, int64_t n_rows,
                        int32_t* tbl_t, int64_t n_src_rows, int32_t* collisions, void* stream);

/* Layout changes at the module boundary (bilateralNN.py:190-196, :221-224):
 *   cm (C, ld_cm) channel-major  <->  rows (n, ld) vertex-major. */
int hpl_cm_to_rows(const float* cm, int64_t ld_cm, int64_t n, int64_t channels, float* rows,
                   int64_t ld, void* stream);
int hpl_rows_to_cm(const float* rows, int64_t ld, int64_t n, int64_t channels, float* cm,
                   int64_t ld_cm, void* stream);

/* sums[c] += sum_n x[c, n]   (bias gradient of SLICE, bilateralNN.py:235-236). x (C, N). */
int hpl_channel_sums(const float* x, int64_t channels, int64_t n, float* sums, void* stream);

/* Patch-correlation stage of BilateralCorrelationFlex (bnn_flow.py:189-202).  The first
 * Conv3d (1,P,1) layer is applied per SOURCE vertex and patch slot with two dense GEMMs
 * (hpl_blur_gemm, nbr = NULL) giving t1 (H1, P*width) and t2 (H2, P*width); this call then forms
 *   z[v*F + f, :] = act(bias + sum_p t1[i1[p,v], p*width:(p+1)*width]
 *                            + sum_p t2[i2[f,p,v], p*width:(p+1)*width])
 * (n1 / n2: rows of t1 / t2; table entries outside [0, n) read zeros)
 * i1 (P, H1) = pc1_corr_indices, i2 (F, P, H1) = pc2_corr_indices, -1 reads zeros.
 * z (H1*F, ldz): row v*F+f, so the same memory is the (H1, F*ldz) operand of the displacement
 * filter (:205).  width % 4 == 0. */
int hpl_corr_gather(const float* t1, int64_t ld1, const void* i1, const float* t2, int64_t ld2,
                    const void* i2, int idx64, const float* bias, int act, float* z, int64_t ldz,
                    int64_t width, int64_t patch, int64_t filt, int64_t h1, int64_t n1, int64_t n2,
                    void* stream);

/* Backward of hpl_corr_gather w.r.t. t1 and t2 (fp32 RED into zeroed dt1 / dt2):
 *   dt2[i2[f,p,v], p, :] += dz[v*F+f, :]      dt1[i1[p,v], p, :] += sum_f dz[v*F+f, :] */
int hpl_corr_scatter(const float* dz, int64_t ldz, const void* i1, const void* i2, int idx64,
                     float* dt1, int64_t ld1, float* dt2, int64_t ld2, int64_t width, int64_t patch,
                     int64_t filt, int64_t h1, int64_t n1, int64_t n2, void* stream);

/* ---------------------------------------------------------------- index half
 * GPU replacement of GenerateDataUnsymmetric + build_unsymmetric + khash
 * (transforms/transforms.py:133-261,264-485; models/khash_int2int.h:8-33).  All results are
 * bit-exact with the reference (vertex ids = first-occurrence order of a point-outer /
 * remainder-inner scan, transforms.py:179-192).  One call sequence per scale:
 *   hpl_lattice_init_range(range)                       once per cloud pair
 *   hpl_lattice_points(cloud 1), hpl_lattice_points(cloud 2)       fold both key ranges (:384-385)
 *   hpl_lattice_insert(cloud 1), hpl_lattice_insert(cloud 2)       hash build, ids, lattice_offset
 *   hpl_lattice_neighbors / hpl_lattice_corr_table                  blur and correlation tables
 *   hpl_lattice_next_points                                         input points of the next scale
 * The khash table (void* handle, get/set per key through cffi) becomes three flat device arrays
 * owned by the caller: table_keys (cap uint64), table_first (cap int32), table_ids (cap int32). */

/* key_minmax[0..3] = INT_MAX, [4..7] = INT_MIN. */
int hpl_lattice_init_range(int32_t* key_minmax, void* stream);

/* get_keys_and_barycentric (transforms.py:300-353) for pc * scale (:377).  pc (3, N) fp32.
 * Outputs: bary (4, N), el_minus_gr (4, N) fp32; greedy (N, 4) int32 = rounded remainder-0
 * point; rankpack (N) = the 4 ranks, one byte each (together they encode the (4, N, 4) int64
 * key tensor of the reference); key_minmax is folded with this cloud's per-coordinate range. */
int hpl_lattice_points(const float* pc, int64_t n_points, float scale, float* bary, float* el_minus_gr,
                       int32_t* greedy, uint32_t* rankpack, int32_t* key_minmax, void* stream);

/* Hash-table capacity (power of two >= 8 N) and scan workspace length (int32) for N points. */
int64_t hpl_lattice_table_capacity(int64_t n_points);
int64_t hpl_lattice_scan_blocks(int64_t n_points);

/* build_unsymmetric hot loop 1 (transforms.py:179-207), khash_get/set (khash_int2int.h:17-33):
 * insert the packed key (key2int, :70-86) of every (point, remainder), number the distinct
 * keys in first-occurrence order, write lattice_offset (4, N) (int64 if idx64 else int32),
 * vertex_coords (>= 4N capacity, 4) int32 = key of each vertex in id order (the reference's
 * last_pc, :188-189) and *n_vertices (device int32) = H (:387-391).
 * slot_of (4N int32) and scan_ws (hpl_lattice_scan_blocks ints) are workspace. */
int hpl_lattice_insert(const int32_t* greedy, const uint32_t* rankpack, int64_t n_points,
                       const int32_t* key_minmax, uint64_t* table_keys, int32_t* table_first,
                       int32_t* table_ids, int64_t table_cap, int32_t* slot_of, int32_t* scan_ws,
                       void* lattice_offset, int idx64, int32_t* vertex_coords, int32_t* n_vertices,
                       void* stream);

/* Hot loops 2 and 4 (transforms.py:209-221,243-255): out[f, h] = id of key_h + offsets[f] in
 * the given table, or -1.  offsets (F, 4) int32 device.  out (F, ld). h_cap = rows to fill. */
int hpl_lattice_neighbors(const int32_t* vertex_coords, const int32_t* n_vertices, int64_t h_cap,
                          const int32_t* key_minmax, const uint64_t* table_keys,
                          const int32_t* table_ids, int64_t table_cap, const int32_t* offsets,
                          int64_t filter_size, void* out, int idx64, int64_t ld, void* stream);

/* Hot loop 3 (transforms.py:223-241): out[f, p, h] = id IN TABLE 2 of
 * key1_h + corr_offsets[p] + filter_offsets[f], or -1.  out (F*P, ld). */
int hpl_lattice_corr_table(const int32_t* vertex_coords1, const int32_t* n_vertices1, int64_t h_cap,
                           const int32_t* key_minmax, const uint64_t* table_keys2,
                           const int32_t* table_ids2, int64_t table_cap2, const int32_t* corr_offsets,
                           int64_t corr_size, const int32_t* filter_offsets, int64_t filter_size,
                           void* out, int idx64, int64_t ld, void* stream);

/* transforms.py:461-467: out (3, H) = E^T . (vertex_coords / divisor), divisor =
 * fp32(expected_std * scale); true division, k-ordered FMA chain. */
int hpl_lattice_next_points(const int32_t* vertex_coords, int64_t n_vertices, float divisor,
                            float* out, void* stream);

int hpl_fill_zero(void* ptr, int64_t bytes, void* stream);
int hpl_fill_i32(int32_t* ptr, int64_t count, int32_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPLFLOWNET_B200_H */
