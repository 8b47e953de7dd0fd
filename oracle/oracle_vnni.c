/*
 * oracle_vnni.c -- the x86 INT8 convolution of oracle.c (oracle_conv_s8_nhwc_x86[_group]) computed with AVX-512 VNNI,
 * for the CPU-baseline arm of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle.c). Same arithmetic, bit for bit: the s32 accumulation is exact in any
 * order (vpdpbusd: u8 x s8 quads into s32; an s8 input is shifted to u8 and the shift is taken back out of the sum as
 * 128 * sum(w) per output channel), and the epilogue is the scalar float sequence of oracle.c (kernel/
 * jit_avx512_core_x8s8s32x_conv_kernel.cpp:137-215). tests/test_cpu_oracle.py checks it against the scalar restatement
 * on every dtype pair, so the pins of oracle.c carry over. It exists because the scalar loop nest runs at < 1 GOP/s per
 * core, which made the reported CPU baseline a strawman next to the reference's MKL / xbyak-JIT x86 path (unbuildable
 * here): this is what an honest, still simple, CPU implementation of the same path does on the box's cores.
 *
 * Layout: NHWC input with c % 4 == 0 (the caller pads the 3-channel graph input to 4), weights KCRS s8.
 * Build: gcc -O2 -mavx512f -mavx512bw -mavx512vnni -ffp-contract=off -fopenmp (this file only; guarded at run time by
 * oracle_vnni_available()).
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))
enum { DT_FLOAT = 1, DT_INT8 = 3, DT_UINT8 = 7 };

ORACLE_API int oracle_vnni_available(void) {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vnni");
}

#define PT 6        /* output pixels per register tile */
#define KB 64       /* output channels per register tile (4 zmm of 16 s32 lanes) */

/* group == 1. Returns 0 on success, 1 when the shape is not supported (c % 4 != 0): the caller uses the scalar path. */
ORACLE_API int oracle_conv_s8_nhwc_x86_vnni(const void* src, int src_dtype, const int8_t* weights, const float* bias_f,
                                            const float* scale, const void* residual, int res_dtype, float sum_scale,
                                            void* dst, int dst_dtype, int n, int c, int h, int w, int k, int kernel_h,
                                            int kernel_w, int stride_h, int stride_w, int dil_h, int dil_w, int pad_h,
                                            int pad_w, int flag_relu) {
    if (c % 4 != 0 || !oracle_vnni_available()) return 1;
    const int out_h = (h + 2 * pad_h - (dil_h * (kernel_h - 1) + 1)) / stride_h + 1;
    const int out_w = (w + 2 * pad_w - (dil_w * (kernel_w - 1) + 1)) / stride_w + 1;
    const int taps = kernel_h * kernel_w, cq = c / 4;
    const int kblocks = (k + KB - 1) / KB, kpad = kblocks * KB;
    const int has_sum = residual != NULL;
    const int in_signed = src_dtype != DT_UINT8;
    /* weights: KCRS -> [tap][c/4][kpad][4]: for one (tap, channel quad) the 16 output channels of a zmm are adjacent */
    int8_t* wp = (int8_t*)aligned_alloc(64, (size_t)taps * cq * kpad * 4);
    memset(wp, 0, (size_t)taps * cq * kpad * 4);
    int32_t* wsum = (int32_t*)calloc(kpad, sizeof(int32_t));
    for (int oc = 0; oc < k; ++oc)
        for (int ic = 0; ic < c; ++ic)
            for (int t = 0; t < taps; ++t) {
                const int8_t v = weights[((size_t)oc * c + ic) * taps + t];
                wp[(((size_t)t * cq + ic / 4) * kpad + oc) * 4 + (ic & 3)] = v;
            }
    /* a zero pixel for padding taps; with a signed input the shifted zero is 128 */
    uint8_t* zpix = (uint8_t*)aligned_alloc(64, (size_t)((c + 63) / 64 * 64));
    memset(zpix, in_signed ? 128 : 0, (size_t)((c + 63) / 64 * 64));
    /* u8 view of the input (s8 + 128) */
    const uint8_t* xin = (const uint8_t*)src;
    uint8_t* shifted = NULL;
    const size_t in_elems = (size_t)n * h * w * c;
    if (in_signed) {
        shifted = (uint8_t*)aligned_alloc(64, (in_elems + 63) / 64 * 64);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < in_elems; ++i) shifted[i] = (uint8_t)(((const int8_t*)src)[i] + 128);
        xin = shifted;
        /* sum over ALL taps and channels of w per output channel: every tap contributes (x + 128) * w, padding taps
         * x = 0 included (zpix = 128), so the correction is the same for every output pixel */
        for (int oc = 0; oc < k; ++oc) {
            int32_t s = 0;
            for (size_t i = 0; i < (size_t)c * taps; ++i) s += weights[(size_t)oc * c * taps + i];
            wsum[oc] = 128 * s;
        }
    }
    const int wtiles = (out_w + PT - 1) / PT;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int row = 0; row < n * out_h; ++row) {
        for (int wt = 0; wt < wtiles; ++wt) {
            const int in_ = row / out_h, oh = row - in_ * out_h;
            const int ow0 = wt * PT;
            const int np = out_w - ow0 < PT ? out_w - ow0 : PT;
            for (int kb = 0; kb < kblocks; ++kb) {
                __m512i acc[PT][4];
                for (int p = 0; p < PT; ++p)
                    for (int q = 0; q < 4; ++q) acc[p][q] = _mm512_setzero_si512();
                for (int kh = 0; kh < kernel_h; ++kh) {
                    const int ih = oh * stride_h - pad_h + kh * dil_h;
                    for (int kw = 0; kw < kernel_w; ++kw) {
                        const uint8_t* px[PT];
                        for (int p = 0; p < PT; ++p) {
                            const int iw = (ow0 + p) * stride_w - pad_w + kw * dil_w;
                            const int ok = p < np && ih >= 0 && ih < h && iw >= 0 && iw < w;
                            px[p] = ok ? xin + (((size_t)in_ * h + ih) * w + iw) * c : zpix;
                        }
                        const int8_t* wt_p = wp + ((size_t)(kh * kernel_w + kw) * cq * kpad + (size_t)kb * KB) * 4;
                        for (int q4 = 0; q4 < cq; ++q4) {
                            const __m512i w0 = _mm512_loadu_si512((const void*)(wt_p + (size_t)q4 * kpad * 4));
                            const __m512i w1 = _mm512_loadu_si512((const void*)(wt_p + (size_t)q4 * kpad * 4 + 64));
                            const __m512i w2 = _mm512_loadu_si512((const void*)(wt_p + (size_t)q4 * kpad * 4 + 128));
                            const __m512i w3 = _mm512_loadu_si512((const void*)(wt_p + (size_t)q4 * kpad * 4 + 192));
                            for (int p = 0; p < PT; ++p) {
                                int32_t quad;
                                memcpy(&quad, px[p] + q4 * 4, 4);
                                const __m512i a = _mm512_set1_epi32(quad);
                                acc[p][0] = _mm512_dpbusd_epi32(acc[p][0], a, w0);
                                acc[p][1] = _mm512_dpbusd_epi32(acc[p][1], a, w1);
                                acc[p][2] = _mm512_dpbusd_epi32(acc[p][2], a, w2);
                                acc[p][3] = _mm512_dpbusd_epi32(acc[p][3], a, w3);
                            }
                        }
                    }
                }
                /* epilogue: the scalar float sequence of oracle_conv_s8_nhwc_x86 */
                for (int p = 0; p < np; ++p) {
                    int32_t a32[KB] __attribute__((aligned(64)));
                    for (int q = 0; q < 4; ++q) _mm512_store_si512((void*)(a32 + 16 * q), acc[p][q]);
                    const size_t out_base = (((size_t)in_ * out_h + oh) * out_w + ow0 + p) * k;
                    for (int j = 0; j < KB; ++j) {
                        const int oc = kb * KB + j;
                        if (oc >= k) break;
                        const int32_t av = a32[j] - wsum[oc];
                        const size_t out_idx = out_base + oc;
                        float f = (float)av + (bias_f ? bias_f[oc] : 0.f);
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
                            const float rr = nearbyintf(f);
                            if (dst_dtype == DT_INT8) ((int8_t*)dst)[out_idx] = (int8_t)(rr > 127.f ? 127 : (rr < -128.f ? -128 : (int32_t)rr));
                            else ((uint8_t*)dst)[out_idx] = (uint8_t)(rr > 255.f ? 255 : (rr < 0.f ? 0 : (int32_t)rr));
                        }
                    }
                }
            }
        }
    }
    free(wp);
    free(wsum);
    free(zpix);
    free(shifted);
    return 0;
}
