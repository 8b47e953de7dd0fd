/*
 * oracle.c -- CPU restatement of the reference's algorithms for the CNN hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under anakin_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / CPU baseline.
 *
 * Every function cites the reference file:line it restates.  The float functions
 * follow the reference's own naive test oracles (test/saber/...), which the
 * reference's tests treat as ground truth at 1e-3; the int8 functions follow the
 * x86 Saber int8 path (saber/funcs/impl/x86/...).  Parity pin: tests/test_cpu_oracle.py
 * compares these restatements with the reference's own code compiled from /root/reference
 * into oracle/_ref/ (make ref): conv_func_helper.h (conv fp32 / int8, int8 pooling),
 * tensor_cmp_host, and the op tests' oracle templates (fp32 pooling, fc, softmax, eltwise,
 * activation) -- bit-exact on everything they share.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle.c -o liboracle.so -lm
 * (-ffp-contract=off: the reference epilogues are separate mul/add except where an
 *  explicit FMA is named below.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* dtype codes = reference DataType (saber/saber_types.h:205-222) */
enum { DT_FLOAT = 1, DT_INT8 = 3, DT_UINT8 = 7 };

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ shapes */
/* Conv output size: reference saber/funcs/funcs_utils.h:41-51 (floor). */
ORACLE_API int oracle_conv_out_size(int in, int pad, int dil, int k, int stride) {
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

/* Pooling output size: reference saber/funcs/pooling.h:69-132.
 * ceil-mode unless floor_as_conv; then drop the last window if it starts in the
 * padding (only when the op is padded). */
ORACLE_API void oracle_pool_out_size(int in_h, int in_w, int window_h, int window_w, int pad_h,
                                     int pad_w, int stride_h, int stride_w, int global_pooling,
                                     int floor_as_conv, int* out_h, int* out_w) {
    int oh, ow;
    if (global_pooling) {
        oh = 1;
        ow = 1;
    } else if (floor_as_conv) {
        oh = (int)(((float)(in_h + 2 * pad_h - window_h) / stride_h)) + 1;
        ow = (int)(((float)(in_w + 2 * pad_w - window_w) / stride_w)) + 1;
        if (oh <= 0) oh = 1;
        if (ow <= 0) ow = 1;
    } else {
        oh = (int)(ceilf((float)(in_h + 2 * pad_h - window_h) / stride_h)) + 1;
        ow = (int)(ceilf((float)(in_w + 2 * pad_w - window_w) / stride_w)) + 1;
    }
    if (!global_pooling && (pad_h > 0 || pad_w > 0)) { /* PoolingParam::pooling_padded() */
        if ((oh - 1) * stride_h >= in_h + pad_h) --oh;
        if ((ow - 1) * stride_w >= in_w + pad_w) --ow;
    }
    *out_h = oh;
    *out_w = ow;
}

/* ------------------------------------------------------------------ fp32 conv */
/* Reference test/saber/conv_func_helper.h:196-264 (conv_basic_check), NCHW:
 *   dst = beta*dst; dst += sum; dst *= alpha; dst += bias; relu.
 * neg_slope extends relu to the x86 impl's leaky form
 * (saber/funcs/impl/x86/saber_im2col_conv.cpp:161-214); 0 == the oracle's plain relu. */
ORACLE_API void oracle_conv_f32_nchw(const float* src, const float* weights, const float* bias,
                                     float* dst, int n, int c, int h, int w, int k, int group,
                                     int kernel_h, int kernel_w, int stride_h, int stride_w,
                                     int dil_h, int dil_w, int pad_h, int pad_w, int flag_bias,
                                     int flag_relu, float neg_slope, float beta, float alpha) {
    const int out_h = oracle_conv_out_size(h, pad_h, dil_h, kernel_h, stride_h);
    const int out_w = oracle_conv_out_size(w, pad_w, dil_w, kernel_w, stride_w);
    const int out_c_group = k / group, in_c_group = c / group;
#pragma omp parallel for collapse(3) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int g = 0; g < group; ++g) {
            for (int oc = 0; oc < out_c_group; ++oc) {
                for (int oh = 0; oh < out_h; ++oh) {
                    for (int ow = 0; ow < out_w; ++ow) {
                        const size_t out_idx = (((size_t)in_ * group + g) * out_c_group + oc) *
                                                   out_h * out_w + (size_t)oh * out_w + ow;
                        const float bias_d = flag_bias ? bias[g * out_c_group + oc] : 0.f;
                        float acc = dst[out_idx] * beta;
                        for (int ic = 0; ic < in_c_group; ++ic) {
                            for (int kh = 0; kh < kernel_h; ++kh) {
                                for (int kw = 0; kw < kernel_w; ++kw) {
                                    const int iw = ow * stride_w - pad_w + kw * dil_w;
                                    const int ih = oh * stride_h - pad_h + kh * dil_h;
                                    if (iw < 0 || iw >= w) continue;
                                    if (ih < 0 || ih >= h) continue;
                                    const size_t iidx =
                                        (((size_t)in_ * c + g * in_c_group + ic) * h + ih) * w + iw;
                                    const size_t widx =
                                        ((((size_t)g * out_c_group + oc) * in_c_group + ic) *
                                             kernel_h + kh) * kernel_w + kw;
                                    acc += src[iidx] * weights[widx];
                                }
                            }
                        }
                        acc *= alpha;
                        acc += bias_d;
                        if (flag_relu) acc = acc > 0.f ? acc : acc * neg_slope;
                        dst[out_idx] = acc;
                    }
                }
            }
        }
    }
}

/* Same computation on NHWC tensors (src [n,h,w,c], dst [n,ho,wo,k], weights KCRS) for the
 * model-level walker and the CPU baseline: x86 epilogue order
 * (saber/funcs/impl/x86/saber_im2col_conv.cpp:161-214): gemm (+ beta*prev) -> +bias -> relu.
 * The dot product runs in (kh, kw, ic) order with ic contiguous and vectorised -- the
 * im2col+GEMM order of the x86 path rather than the naive oracle's (ic, kh, kw); the two
 * differ by float re-association only (tests/test_oracle.py bounds it at 1e-5 relative). */
ORACLE_API void oracle_conv_f32_nhwc(const float* src, const float* weights, const float* bias,
                                     const float* residual, float* dst, int n, int c, int h, int w,
                                     int k, int group, int kernel_h, int kernel_w, int stride_h,
                                     int stride_w, int dil_h, int dil_w, int pad_h, int pad_w,
                                     int flag_bias, int flag_relu, float neg_slope, float beta) {
    const int out_h = oracle_conv_out_size(h, pad_h, dil_h, kernel_h, stride_h);
    const int out_w = oracle_conv_out_size(w, pad_w, dil_w, kernel_w, stride_w);
    const int out_c_group = k / group, in_c_group = c / group;
    /* weights re-laid as [k][kh][kw][ic] so the inner loop is contiguous in both operands */
    const size_t wsz = (size_t)k * in_c_group * kernel_h * kernel_w;
    float* wt = (float*)malloc(wsz * sizeof(float));
    for (int oc = 0; oc < k; ++oc)
        for (int ic = 0; ic < in_c_group; ++ic)
            for (int kh = 0; kh < kernel_h; ++kh)
                for (int kw = 0; kw < kernel_w; ++kw)
                    wt[(((size_t)oc * kernel_h + kh) * kernel_w + kw) * in_c_group + ic] =
                        weights[(((size_t)oc * in_c_group + ic) * kernel_h + kh) * kernel_w + kw];
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            for (int ow = 0; ow < out_w; ++ow) {
                for (int och = 0; och < k; ++och) {
                    const int g = och / out_c_group;
                    const size_t out_idx = (((size_t)in_ * out_h + oh) * out_w + ow) * k + och;
                    float acc = residual ? residual[out_idx] * beta : 0.f;
                    for (int kh = 0; kh < kernel_h; ++kh) {
                        const int ih = oh * stride_h - pad_h + kh * dil_h;
                        if (ih < 0 || ih >= h) continue;
                        for (int kw = 0; kw < kernel_w; ++kw) {
                            const int iw = ow * stride_w - pad_w + kw * dil_w;
                            if (iw < 0 || iw >= w) continue;
                            const float* ip = src + (((size_t)in_ * h + ih) * w + iw) * c + g * in_c_group;
                            const float* wp = wt + (((size_t)och * kernel_h + kh) * kernel_w + kw) * in_c_group;
                            float part = 0.f;
#pragma omp simd reduction(+ : part)
                            for (int ic = 0; ic < in_c_group; ++ic) part += ip[ic] * wp[ic];
                            acc += part;
                        }
                    }
                    acc += flag_bias ? bias[och] : 0.f;
                    if (flag_relu) acc = acc > 0.f ? acc : acc * neg_slope;
                    dst[out_idx] = acc;
                }
            }
        }
    }
    free(wt);
}

/* ------------------------------------------------------------------ int8 conv */
static inline int8_t saturate_s8(int32_t v) { return (int8_t)(v > 127 ? 127 : (v < -128 ? -128 : v)); }

/* Reference test/saber/conv_func_helper.h:102-192 (conv_basic_check_int8), NHWC,
 * float accumulation, int32 bias, per-oc scale, optional eltwise-sum, nearbyintf +
 * saturate<int8_t> (rm nearest) or floorf (rm down). dst is read for beta / sum. */
ORACLE_API void oracle_conv_s8_nhwc_basic(const void* src, int src_is_unsigned,
                                          const int8_t* weights, const int32_t* bias, int8_t* dst,
                                          int n, int c, int h, int w, int k, int group,
                                          int kernel_h, int kernel_w, int stride_h, int stride_w,
                                          int dil_h, int dil_w, int pad_h, int pad_w, int flag_bias,
                                          int flag_relu, const float* scale, int has_elt_sum,
                                          float sum_scale, float beta, int round_down) {
    const uint8_t* src_u8 = (const uint8_t*)src;
    const int8_t* src_s8 = (const int8_t*)src;
    const int out_h = oracle_conv_out_size(h, pad_h, dil_h, kernel_h, stride_h);
    const int out_w = oracle_conv_out_size(w, pad_w, dil_w, kernel_w, stride_w);
    const int out_c_group = k / group, in_c_group = c / group;
#pragma omp parallel for collapse(3) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            for (int ow = 0; ow < out_w; ++ow) {
                for (int g = 0; g < group; ++g) {
                    for (int oc = 0; oc < out_c_group; ++oc) {
                        const size_t out_idx =
                            (((size_t)in_ * out_h + oh) * out_w + ow) * k + g * out_c_group + oc;
                        const float bias_d = flag_bias ? (float)bias[g * out_c_group + oc] : 0.f;
                        float v = bias_d + dst[out_idx] * beta;
                        for (int ic = 0; ic < in_c_group; ++ic) {
                            for (int kh = 0; kh < kernel_h; ++kh) {
                                for (int kw = 0; kw < kernel_w; ++kw) {
                                    const int iw = ow * stride_w - pad_w + kw * dil_w;
                                    const int ih = oh * stride_h - pad_h + kh * dil_h;
                                    if (iw < 0 || iw >= w) continue;
                                    if (ih < 0 || ih >= h) continue;
                                    const size_t iidx =
                                        (((size_t)in_ * h + ih) * w + iw) * c + g * in_c_group + ic;
                                    const size_t widx =
                                        ((((size_t)g * out_c_group + oc) * in_c_group + ic) *
                                             kernel_h + kh) * kernel_w + kw;
                                    const float a = src_is_unsigned ? (float)src_u8[iidx]
                                                                    : (float)src_s8[iidx];
                                    v += a * weights[widx];
                                }
                            }
                        }
                        v = v * scale[g * out_c_group + oc];
                        if (has_elt_sum) v += dst[out_idx] * sum_scale;
                        if (flag_relu) v = v > 0.f ? v : 0.f;
                        dst[out_idx] = round_down ? saturate_s8((int32_t)floorf(v))
                                                  : saturate_s8((int32_t)nearbyintf(v));
                    }
                }
            }
        }
    }
}

/* x86 Saber int8 conv (what a VNNI CPU executes): exact s32 accumulation
 * (cblas_gemm_s8u8s32 / vpdpbusd, saber/funcs/impl/x86/gemm_x8s8s32x_conv.cpp:244-281)
 * then the JIT epilogue of
 * saber/funcs/impl/x86/kernel/jit_avx512_core_x8s8s32x_conv_kernel.cpp:137-215:
 *   f = (float)acc + bias_f[oc];  f *= scale[oc];
 *   if (relu && !sum) f = max(f,0);
 *   if (sum)  f = (sum_scale==1) ? f + (float)prev : fma((float)prev, sum_scale, f);
 *   if (relu && sum) f = max(f,0);
 *   out = f (fp32) | vcvtps2dq(RN-even) + saturating pack to s8 / u8.
 * bias_f / scale / sum_scale are the host-prepared values of
 * kernel/jit_avx512_core_x8s8s32x_conv.cpp:55-62,174-192,226-255 (see oracle_int8_conv_scales).
 * NHWC; weights KCRS s8; residual has the geometry of dst. group == 1. */
ORACLE_API void oracle_conv_s8_nhwc_x86(const void* src, int src_dtype, const int8_t* weights,
                                        const float* bias_f, const float* scale,
                                        const void* residual, int res_dtype, float sum_scale,
                                        void* dst, int dst_dtype, int n, int c, int h, int w, int k,
                                        int kernel_h, int kernel_w, int stride_h, int stride_w,
                                        int dil_h, int dil_w, int pad_h, int pad_w, int flag_relu) {
    const uint8_t* src_u8 = (const uint8_t*)src;
    const int8_t* src_s8 = (const int8_t*)src;
    const int out_h = oracle_conv_out_size(h, pad_h, dil_h, kernel_h, stride_h);
    const int out_w = oracle_conv_out_size(w, pad_w, dil_w, kernel_w, stride_w);
    const int has_sum = residual != NULL;
    /* re-lay the weights as [k][r][s][c] once so the inner loop is contiguous */
    const size_t wsz = (size_t)k * c * kernel_h * kernel_w;
    int8_t* wt = (int8_t*)malloc(wsz);
    for (int oc = 0; oc < k; ++oc)
        for (int ic = 0; ic < c; ++ic)
            for (int kh = 0; kh < kernel_h; ++kh)
                for (int kw = 0; kw < kernel_w; ++kw)
                    wt[(((size_t)oc * kernel_h + kh) * kernel_w + kw) * c + ic] =
                        weights[(((size_t)oc * c + ic) * kernel_h + kh) * kernel_w + kw];
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            for (int ow = 0; ow < out_w; ++ow) {
                for (int oc = 0; oc < k; ++oc) {
                    int32_t acc = 0;
                    for (int kh = 0; kh < kernel_h; ++kh) {
                        const int ih = oh * stride_h - pad_h + kh * dil_h;
                        if (ih < 0 || ih >= h) continue;
                        for (int kw = 0; kw < kernel_w; ++kw) {
                            const int iw = ow * stride_w - pad_w + kw * dil_w;
                            if (iw < 0 || iw >= w) continue;
                            const size_t ibase = (((size_t)in_ * h + ih) * w + iw) * c;
                            const int8_t* wp = wt + (((size_t)oc * kernel_h + kh) * kernel_w + kw) * c;
                            int32_t part = 0;
                            if (src_dtype == DT_UINT8) {
                                _Pragma("omp simd reduction(+ : part)")
                                for (int ic = 0; ic < c; ++ic) part += (int32_t)src_u8[ibase + ic] * wp[ic];
                            } else {
                                _Pragma("omp simd reduction(+ : part)")
                                for (int ic = 0; ic < c; ++ic) part += (int32_t)src_s8[ibase + ic] * wp[ic];
                            }
                            acc += part;
                        }
                    }
                    const size_t out_idx = (((size_t)in_ * out_h + oh) * out_w + ow) * k + oc;
                    float f = (float)acc + (bias_f ? bias_f[oc] : 0.f);
                    f = f * (scale ? scale[oc] : 1.f);
                    if (flag_relu && !has_sum) f = f > 0.f ? f : 0.f;
                    if (has_sum) {
                        float r;
                        if (res_dtype == DT_FLOAT) r = ((const float*)residual)[out_idx];
                        else if (res_dtype == DT_UINT8) r = (float)((const uint8_t*)residual)[out_idx];
                        else r = (float)((const int8_t*)residual)[out_idx];
                        f = (sum_scale == 1.f) ? f + r : fmaf(r, sum_scale, f);
                        if (flag_relu) f = f > 0.f ? f : 0.f;
                    }
                    if (dst_dtype == DT_FLOAT) {
                        ((float*)dst)[out_idx] = f;
                    } else {
                        /* vcvtps2dq RN-even, then vpmovsdb / vpmovusdb saturation */
                        float rr = nearbyintf(f);
                        int32_t q;
                        if (dst_dtype == DT_INT8) {
                            q = rr > 127.f ? 127 : (rr < -128.f ? -128 : (int32_t)rr);
                            ((int8_t*)dst)[out_idx] = (int8_t)q;
                        } else {
                            q = rr > 255.f ? 255 : (rr < 0.f ? 0 : (int32_t)rr);
                            ((uint8_t*)dst)[out_idx] = (uint8_t)q;
                        }
                    }
                }
            }
        }
    }
    free(wt);
}

/* The same with channel groups (depthwise: group == c == k; weights [k][c/group][r][s]). The arithmetic per output is
 * unchanged -- an exact s32 sum over the group's input channels, then the epilogue above; the reference's depthwise
 * INT8 kernels (jit_avx512_core_x8s8s32x_1x1 / dw variants) share that epilogue. */
ORACLE_API void oracle_conv_s8_nhwc_x86_group(const void* src, int src_dtype, const int8_t* weights,
                                        const float* bias_f, const float* scale,
                                        const void* residual, int res_dtype, float sum_scale,
                                        void* dst, int dst_dtype, int n, int c, int h, int w, int k, int group,
                                        int kernel_h, int kernel_w, int stride_h, int stride_w,
                                        int dil_h, int dil_w, int pad_h, int pad_w, int flag_relu) {
    const uint8_t* src_u8 = (const uint8_t*)src;
    const int8_t* src_s8 = (const int8_t*)src;
    const int out_h = oracle_conv_out_size(h, pad_h, dil_h, kernel_h, stride_h);
    const int out_w = oracle_conv_out_size(w, pad_w, dil_w, kernel_w, stride_w);
    const int has_sum = residual != NULL;
    /* re-lay the weights as [k][r][s][c] once so the inner loop is contiguous */
    const int cg = c / group, kg = k / group;   /* channels per group (weights are [k][c/group][r][s]) */
    const size_t wsz = (size_t)k * cg * kernel_h * kernel_w;
    int8_t* wt = (int8_t*)malloc(wsz);
    for (int oc = 0; oc < k; ++oc)
        for (int ic = 0; ic < cg; ++ic)
            for (int kh = 0; kh < kernel_h; ++kh)
                for (int kw = 0; kw < kernel_w; ++kw)
                    wt[(((size_t)oc * kernel_h + kh) * kernel_w + kw) * cg + ic] =
                        weights[(((size_t)oc * cg + ic) * kernel_h + kh) * kernel_w + kw];
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            for (int ow = 0; ow < out_w; ++ow) {
                for (int oc = 0; oc < k; ++oc) {
                    int32_t acc = 0;
                    for (int kh = 0; kh < kernel_h; ++kh) {
                        const int ih = oh * stride_h - pad_h + kh * dil_h;
                        if (ih < 0 || ih >= h) continue;
                        for (int kw = 0; kw < kernel_w; ++kw) {
                            const int iw = ow * stride_w - pad_w + kw * dil_w;
                            if (iw < 0 || iw >= w) continue;
                            const size_t ibase = (((size_t)in_ * h + ih) * w + iw) * c + (size_t)(oc / kg) * cg;
                            const int8_t* wp = wt + (((size_t)oc * kernel_h + kh) * kernel_w + kw) * cg;
                            int32_t part = 0;
                            if (src_dtype == DT_UINT8) {
                                _Pragma("omp simd reduction(+ : part)")
                                for (int ic = 0; ic < cg; ++ic) part += (int32_t)src_u8[ibase + ic] * wp[ic];
                            } else {
                                _Pragma("omp simd reduction(+ : part)")
                                for (int ic = 0; ic < cg; ++ic) part += (int32_t)src_s8[ibase + ic] * wp[ic];
                            }
                            acc += part;
                        }
                    }
                    const size_t out_idx = (((size_t)in_ * out_h + oh) * out_w + ow) * k + oc;
                    float f = (float)acc + (bias_f ? bias_f[oc] : 0.f);
                    f = f * (scale ? scale[oc] : 1.f);
                    if (flag_relu && !has_sum) f = f > 0.f ? f : 0.f;
                    if (has_sum) {
                        float r;
                        if (res_dtype == DT_FLOAT) r = ((const float*)residual)[out_idx];
                        else if (res_dtype == DT_UINT8) r = (float)((const uint8_t*)residual)[out_idx];
                        else r = (float)((const int8_t*)residual)[out_idx];
                        f = (sum_scale == 1.f) ? f + r : fmaf(r, sum_scale, f);
                        if (flag_relu) f = f > 0.f ? f : 0.f;
                    }
                    if (dst_dtype == DT_FLOAT) {
                        ((float*)dst)[out_idx] = f;
                    } else {
                        /* vcvtps2dq RN-even, then vpmovsdb / vpmovusdb saturation */
                        float rr = nearbyintf(f);
                        int32_t q;
                        if (dst_dtype == DT_INT8) {
                            q = rr > 127.f ? 127 : (rr < -128.f ? -128 : (int32_t)rr);
                            ((int8_t*)dst)[out_idx] = (int8_t)q;
                        } else {
                            q = rr > 255.f ? 255 : (rr < 0.f ? 0 : (int32_t)rr);
                            ((uint8_t*)dst)[out_idx] = (uint8_t)q;
                        }
                    }
                }
            }
        }
    }
    free(wt);
}

/* Host-side scale preparation of the x86 int8 conv:
 * kernel/jit_avx512_core_x8s8s32x_conv.cpp:55-62 (bias), :226-255 (scale), :174-192 (sum_scale).
 *   in_scale / out_scale are the calibrated max|x|/127 edge scales; u8 tensors carry
 *   an extra 127/255 factor. bias_f[i] = bias[i] * (1/(w_scale[i]*in_scale*(u8?127/255:1))). */
ORACLE_API void oracle_int8_conv_scales(const float* w_scale, const float* bias, int k,
                                        float in_scale, int in_dtype, float out_scale,
                                        int out_dtype, float res_scale, int res_dtype,
                                        float* scale_out, float* bias_f_out, float* sum_scale_out) {
    const float u = 127.f / 255.f;
    for (int i = 0; i < k; ++i) {
        float s;
        if (in_dtype == DT_INT8 && out_dtype == DT_INT8) s = (w_scale[i] * in_scale) / out_scale;
        else if (in_dtype == DT_UINT8 && out_dtype == DT_UINT8) s = (w_scale[i] * in_scale * u) / (out_scale * u);
        else if (in_dtype == DT_UINT8 && out_dtype == DT_INT8) s = (w_scale[i] * in_scale * u) / out_scale;
        else if (in_dtype == DT_UINT8 && out_dtype == DT_FLOAT) s = w_scale[i] * in_scale * u;
        else if (in_dtype == DT_INT8 && out_dtype == DT_UINT8) s = (w_scale[i] * in_scale) / (out_scale * u);
        else s = w_scale[i] * in_scale; /* s8 -> f32 */
        scale_out[i] = s;
        if (bias_f_out) {
            float inv = (in_dtype == DT_UINT8) ? (1.f / (w_scale[i] * in_scale * u))
                                               : (1.f / (w_scale[i] * in_scale));
            bias_f_out[i] = bias ? bias[i] * inv : 0.f;
        }
    }
    if (sum_scale_out) {
        /* beta = scale of the residual edge (framework/operators/fusion_ops/conv_eltwise.cpp:182-188) */
        float ss;
        if (res_dtype == DT_INT8 && out_dtype == DT_UINT8) ss = res_scale * (255.f / 127.f) / out_scale;
        else if (res_dtype == DT_UINT8 && out_dtype == DT_INT8) ss = res_scale * (127.f / 255.f) / out_scale;
        else ss = res_scale / out_scale;
        *sum_scale_out = ss;
    }
}

/* ------------------------------------------------------------------ quantisation */
/* Weights: saber/funcs/impl/x86/x86_utils.h:293-323 -- per-output-channel
 * s_w = max|w|/127, truncating static_cast<char>(w / s_w). */
ORACLE_API void oracle_quant_weights_per_oc(const float* w, int k, int per_k, int8_t* out,
                                            float* scale_out) {
    for (int oc = 0; oc < k; ++oc) {
        float mx = 0.f;
        for (int i = 0; i < per_k; ++i) {
            float a = fabsf(w[(size_t)oc * per_k + i]);
            mx = a > mx ? a : mx;
        }
        float s = mx / 127.f;
        if (s == 0.f) s = 1.f;
        scale_out[oc] = s;
        for (int i = 0; i < per_k; ++i) out[(size_t)oc * per_k + i] = (int8_t)(w[(size_t)oc * per_k + i] / s);
    }
}
/* Activations fp32 -> s8: x86_utils.h:318-347 secur_cast2char(x * (1/scale)): roundf + clamp. */
ORACLE_API void oracle_quant_fp32_s8(const float* in, int8_t* out, size_t count, float scale) {
    const float inv = 1.f / scale;
    for (size_t i = 0; i < count; ++i) {
        float t = roundf(in[i] * inv);
        int ti = (int)t;
        ti = ti > 127 ? 127 : ti;
        ti = ti < -128 ? -128 : ti;
        out[i] = (int8_t)ti;
    }
}
/* fp32 -> u8: x86_utils.h:360-372: static_cast<unsigned char>(x * 1/(scale*127/255)) (truncation). */
ORACLE_API void oracle_quant_fp32_u8(const float* in, uint8_t* out, size_t count, float scale) {
    const float inv = 1.f / (scale * (127.f / 255.f));
    for (size_t i = 0; i < count; ++i) {
        float t = in[i] * inv;
        t = t < 0.f ? 0.f : (t > 255.f ? 255.f : t); /* C UB guard; the reference relies on in-range data */
        out[i] = (uint8_t)t;
    }
}

/* ------------------------------------------------------------------ BN/Scale fold */
/* framework/utils/parameter_fusion.cpp:86-131 (fp32): per output channel
 *   f = bn_scale_factor==0 ? 1 : 1/bn_scale_factor
 *   alpha = 1/sqrt(var*f + eps); beta = -(mean*f)*alpha
 *   alpha *= gamma; beta = beta*gamma (+ beta_s)
 *   w *= alpha; b = b*alpha + beta. */
ORACLE_API void oracle_fold_bn_scale(float* weights, float* bias, int k, int per_k,
                                     float bn_scale_factor, float eps, const float* mean,
                                     const float* var, const float* gamma, const float* beta_s) {
    const float f = (bn_scale_factor == 0.f) ? 1.f : 1.f / bn_scale_factor;
    for (int i = 0; i < k; ++i) {
        float alpha = var[i] * f + eps;
        alpha = 1.f / sqrtf(alpha);
        float beta = -1.f * (mean[i] * f);
        beta = beta * alpha;
        alpha = gamma[i] * alpha;
        if (beta_s) beta = beta * gamma[i] + beta_s[i];
        else beta = beta * gamma[i];
        for (int j = 0; j < per_k; ++j) weights[(size_t)i * per_k + j] *= alpha;
        bias[i] *= alpha;
        bias[i] += beta;
    }
}

/* ------------------------------------------------------------------ pooling */
/* fp32: reference test/saber/test_saber_pooling.cpp:14-104. layout_nhwc selects the
 * memory order only; the window logic is the reference's. type: 1 max, 2 avg incl pad,
 * 3 avg excl pad. */
ORACLE_API void oracle_pool_f32(const float* src, float* dst, int n, int c, int in_h, int in_w,
                                int out_h, int out_w, int window_h, int window_w, int pad_h,
                                int pad_w, int stride_h, int stride_w, int type, int layout_nhwc) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int ic = 0; ic < c; ++ic) {
            for (int oh = 0; oh < out_h; ++oh) {
                int sh = oh * stride_h, eh = sh + window_h;
                sh = (sh - pad_h) < 0 ? 0 : sh - pad_h;
                eh = (eh - pad_h) > in_h ? in_h : eh - pad_h;
                for (int ow = 0; ow < out_w; ++ow) {
                    int sw = ow * stride_w, ew = sw + window_w;
                    sw = (sw - pad_w) < 0 ? 0 : sw - pad_w;
                    ew = (ew - pad_w) > in_w ? in_w : ew - pad_w;
                    float result = 0.f;
                    for (int kh = sh; kh < eh; ++kh) {
                        for (int kw = sw; kw < ew; ++kw) {
                            const size_t si = layout_nhwc
                                ? (((size_t)in_ * in_h + kh) * in_w + kw) * c + ic
                                : (((size_t)in_ * c + ic) * in_h + kh) * in_w + kw;
                            if (kh == sh && kw == sw) result = src[si];
                            else if (type == 1) result = result >= src[si] ? result : src[si];
                            else result += src[si];
                        }
                    }
                    if (type == 2) {
                        int bh = window_h, bw = window_w;
                        if (ew == in_w) {
                            bw = sw + window_w >= in_w + pad_w ? in_w + pad_w : sw + window_w;
                            bw -= sw;
                        }
                        if (eh == in_h) {
                            bh = sh + window_h >= in_h + pad_h ? in_h + pad_h : sh + window_h;
                            bh -= sh;
                        }
                        result /= bh * bw;
                    }
                    if (type == 3) result /= (ew - sw) * (eh - sh);
                    const size_t di = layout_nhwc
                        ? (((size_t)in_ * out_h + oh) * out_w + ow) * c + ic
                        : (((size_t)in_ * c + ic) * out_h + oh) * out_w + ow;
                    dst[di] = result;
                }
            }
        }
    }
}

/* int8 NHWC: reference test/saber/conv_func_helper.h:29-100 (pool_basic_check_int8):
 * float accumulation of the raw codes, avg-incl divides by window_h*window_w, result
 * nearbyintf() then cast.  is_unsigned selects u8 vs s8 codes (the reference reads
 * `char`; AK_UINT8 tensors hold relu outputs). */
ORACLE_API void oracle_pool_s8_nhwc(const void* src, void* dst, int is_unsigned, int n, int c,
                                    int in_h, int in_w, int out_h, int out_w, int window_h,
                                    int window_w, int pad_h, int pad_w, int stride_h, int stride_w,
                                    int type) {
    const int8_t* s8 = (const int8_t*)src;
    const uint8_t* u8 = (const uint8_t*)src;
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            int sh = oh * stride_h, eh = sh + window_h;
            if (pad_h > 0) {
                sh = (sh - pad_h) < 0 ? 0 : sh - pad_h;
                eh = (eh - pad_h) > in_h ? in_h : eh - pad_h;
            }
            if (eh > in_h) eh = in_h; /* ceil-mode windows past the edge (reference relies on pad>0) */
            for (int ow = 0; ow < out_w; ++ow) {
                int sw = ow * stride_w, ew = sw + window_w;
                if (pad_w > 0) {
                    sw = (sw - pad_w) < 0 ? 0 : sw - pad_w;
                    ew = (ew - pad_w) > in_w ? in_w : ew - pad_w;
                }
                if (ew > in_w) ew = in_w;
                for (int ic = 0; ic < c; ++ic) {
                    float result = 0.f;
                    for (int kh = sh; kh < eh; ++kh) {
                        for (int kw = sw; kw < ew; ++kw) {
                            const size_t si = (((size_t)in_ * in_h + kh) * in_w + kw) * c + ic;
                            const float v = is_unsigned ? (float)u8[si] : (float)s8[si];
                            if (kh == sh && kw == sw) result = v;
                            else if (type == 1) result = result >= v ? result : v;
                            else result += v;
                        }
                    }
                    if (type == 2) result /= window_h * window_w;
                    if (type == 3) result /= (ew - sw) * (eh - sh);
                    const size_t di = (((size_t)in_ * out_h + oh) * out_w + ow) * c + ic;
                    const float r = nearbyintf(result);
                    if (is_unsigned) ((uint8_t*)dst)[di] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
                    else ((int8_t*)dst)[di] = (int8_t)(r < -128.f ? -128.f : (r > 127.f ? 127.f : r));
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ fc */
/* reference test/saber/test_saber_fc.cpp:14-47: out[i][j] = bias[j] + sum_k in[i][k]*W[j][k]
 * accumulated in k order starting from the bias. */
ORACLE_API void oracle_fc_f32(const float* in, const float* weights, const float* bias, float* out,
                              int m, int k, int n_out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n_out; ++j) {
            float acc = bias ? bias[j] : 0.f;
            for (int q = 0; q < k; ++q) acc += in[(size_t)i * k + q] * weights[(size_t)j * k + q];
            out[(size_t)i * n_out + j] = acc;
        }
    }
}

/* x86 int8 fc (saber/funcs/impl/x86/vender_fc.cpp:363-382 + the scale step): exact s32
 * accumulation, out = (float)acc * scale[j] + bias[j] (fp32 output). */
ORACLE_API void oracle_fc_s8(const void* in, int in_dtype, const int8_t* weights, const float* bias,
                             const float* scale, float* out, int m, int k, int n_out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n_out; ++j) {
            int32_t acc = 0;
            if (in_dtype == DT_UINT8) {
                const uint8_t* a = (const uint8_t*)in + (size_t)i * k;
                for (int q = 0; q < k; ++q) acc += (int32_t)a[q] * weights[(size_t)j * k + q];
            } else {
                const int8_t* a = (const int8_t*)in + (size_t)i * k;
                for (int q = 0; q < k; ++q) acc += (int32_t)a[q] * weights[(size_t)j * k + q];
            }
            out[(size_t)i * n_out + j] = (float)acc * scale[j] + (bias ? bias[j] : 0.f);
        }
    }
}

/* ------------------------------------------------------------------ softmax */
/* reference test/saber/test_saber_softmax.cpp:22-79 / x86 saber_softmax.cpp:147-188:
 * subtract max, exp, divide by sum along the axis of an [outer][axis][inner] view. */
ORACLE_API void oracle_softmax_f32(const float* in, float* out, int outer, int axis_size, int inner) {
    for (int o = 0; o < outer; ++o) {
        for (int i = 0; i < inner; ++i) {
            const float* p = in + (size_t)o * axis_size * inner + i;
            float* q = out + (size_t)o * axis_size * inner + i;
            float mx = -3.402823466e+38f;
            for (int a = 0; a < axis_size; ++a) mx = p[(size_t)a * inner] > mx ? p[(size_t)a * inner] : mx;
            float sum = 0.f;
            for (int a = axis_size - 1; a >= 0; --a) {
                const float e = (float)exp(p[(size_t)a * inner] - mx);
                q[(size_t)a * inner] = e;
                sum += e;
            }
            for (int a = 0; a < axis_size; ++a) q[(size_t)a * inner] = q[(size_t)a * inner] / sum;
        }
    }
}

/* ------------------------------------------------------------------ eltwise / activation / scale */
/* reference test/saber/test_saber_eltwise.cpp:16-98. op: 1 prod, 2 sum (coeff), 3 max. */
ORACLE_API void oracle_eltwise_f32(const float* a, const float* b, float* out, size_t count, int op,
                                   float c0, float c1, int relu) {
    for (size_t e = 0; e < count; ++e) {
        float v;
        if (op == 2) { v = c0 * a[e]; v += c1 * b[e]; }
        else if (op == 1) v = a[e] * b[e];
        else v = a[e] > b[e] ? a[e] : b[e];
        if (relu) v = v > 0.f ? v : 0.f;
        out[e] = v;
    }
}

/* x86 int8 eltwise sum (saber/funcs/impl/x86/saber_eltwise.cpp:72-111, simple_sum):
 *   tmp = coeff0*(float)a*scale_a + coeff1*(float)b*scale_b; relu; saturate(roundf(tmp)).
 * sa / sb are the combined coeff*scale factors (identical to the reference's product
 * order whenever coeff == 1, the ResNet case; the caller also folds 1/out_scale in). */
ORACLE_API void oracle_eltwise_sum_q8(const void* a, int a_dtype, const void* b, int b_dtype,
                                      void* out, int out_dtype, size_t count, float sa, float sb,
                                      int relu) {
    for (size_t e = 0; e < count; ++e) {
        const float fa = (a_dtype == DT_UINT8) ? (float)((const uint8_t*)a)[e] : (float)((const int8_t*)a)[e];
        const float fb = (b_dtype == DT_UINT8) ? (float)((const uint8_t*)b)[e] : (float)((const int8_t*)b)[e];
        float f = fa * sa;
        f += fb * sb;
        if (relu) f = f > 0.f ? f : 0.f;
        const float r = roundf(f);
        if (out_dtype == DT_UINT8) ((uint8_t*)out)[e] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
        else ((int8_t*)out)[e] = (int8_t)(r < -128.f ? -128.f : (r > 127.f ? 127.f : r));
    }
}

/* reference test/saber/test_saber_activation.cpp:16-138. act = ActiveType. */
ORACLE_API void oracle_activation_f32(const float* in, float* out, size_t count, int act,
                                      float neg_slope, float coef) {
    for (size_t i = 0; i < count; ++i) {
        const float x = in[i];
        float y = x;
        switch (act) {
            case 2: y = x > 0.f ? x : x * neg_slope; break;                 /* relu */
            /* the reference's templates call the C double-precision exp / tanh on the promoted float and only
             * round when storing (test_saber_activation.cpp:36-66): restated expression for expression */
            case 1: y = (float)(1.0f / (exp(-x) + 1.0f)); break;                   /* sigmoid */
            case 3: y = (float)tanh(x); break;                                     /* tanh */
            case 4: y = x > 0.f ? x : 0.f; y = y < coef ? y : coef; break;         /* clipped relu */
            case 5: y = x > 0.f ? x : (float)(coef * (exp(x) - 1)); break;         /* elu */
            default: break;
        }
        out[i] = y;
    }
}

/* y = x*w[c] + b[c] over an [outer][c][inner] view (saber/funcs/impl/x86/saber_scale.cpp). */
ORACLE_API void oracle_scale_f32(const float* in, float* out, int outer, int c, int inner,
                                 const float* w, const float* b) {
    for (int o = 0; o < outer; ++o)
        for (int ic = 0; ic < c; ++ic)
            for (int i = 0; i < inner; ++i) {
                const size_t idx = ((size_t)o * c + ic) * inner + i;
                out[idx] = in[idx] * w[ic] + (b ? b[ic] : 0.f);
            }
}

/* ------------------------------------------------------------------ compare */
/* saber/core/tensor_op.cpp:580-599 (tensor_cmp_host): the max abs diff and the relative
 * error 2*|a-b|/(a+b+1e-6) *of that element*. */
ORACLE_API void oracle_tensor_cmp(const float* a, const float* b, size_t count, double* max_ratio,
                                  double* max_diff) {
    const double eps = 1e-6f;
    double md = 0, mr = 0;
    for (size_t i = 0; i < count; ++i) {
        const float df = a[i] - b[i]; /* float subtraction, as tensor_cmp_host<float> does */
        const double d = fabs(df);
        if (md < d) {
            md = d;
            const float sf = a[i] + b[i];
            mr = fabs(2.0 * md / (sf + eps));
        }
    }
    *max_diff = md;
    *max_ratio = mr;
}
