/*
 * oracle_vnni.c -- the x86 INT8 convolution of oracle.c (oracle_conv_s8_nhwc_x86) computed with AVX-512 VNNI, for the
 * CPU-baseline arm of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle.c). Same arithmetic, bit for bit: the s32 accumulation is exact in any
 * order (vpdpbusd: u8 x s8 quads into s32; an s8 input is shifted to u8 and the shift is taken back out of the sum as
 * 128 * sum(w) per output channel), and the epilogue is the float sequence of oracle.c (kernel/
 * jit_avx512_core_x8s8s32x_conv_kernel.cpp:137-215) carried out 16 lanes at a time: cvt, add, mul, max, fma and the
 * round-to-nearest-even conversion are the same IEEE operations per lane as the scalar statements (clamping before the
 * rounding instead of after it gives the same integer for every finite value). tests/test_cpu_oracle.py checks it against
 * the scalar restatement on every dtype pair, so the pins of oracle.c carry over. It exists because the scalar loop nest
 * runs at < 1 GOP/s per core, which made the reported CPU baseline a strawman next to the reference's MKL / xbyak-JIT x86
 * path (unbuildable here): this is what an honest, still simple, CPU implementation of the same path does on the box's
 * cores.
 *
 * Structure (what the reference's JIT kernels do as well): weights packed ONCE per layer (oracle_vnni_pack -- init-time
 * work like the reference's trans_weights) into [k / 32][tap][c / 4][32][4], a register tile of 14 output pixels x 32
 * output channels (28 zmm accumulators), pixels taken linearly over (n, oh, ow) so that 7 x 7 maps fill tiles too, the
 * block of one 32-channel group (taps * c * 32 bytes) staying in L2 while the pixel tiles stream past it.
 *
 * Layout: NHWC input with c % 4 == 0 (the caller pads the 3-channel graph input to 4), weights KCRS s8.
 * Build: gcc -O2 -mavx512f -mavx512bw -mavx512vl -mavx512vnni -ffp-contract=off -fopenmp (this file only; guarded at run time by
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
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("avx512vnni");
}

#define PT 14       /* output pixels per register tile */
#define KB 32       /* output channels per register tile (2 zmm of 16 s32 lanes) */

typedef struct {
    int k, c, kh, kw, kblocks, cq, taps;
    int8_t* wp;         /* [kblocks][taps][cq][KB][4] */
    int32_t* wsum128;   /* [kblocks * KB]: 128 * sum over taps and channels of w (the shift correction of a signed input) */
} vnni_pack_t;

ORACLE_API void* oracle_vnni_pack(const int8_t* weights, int k, int c, int kernel_h, int kernel_w) {
    if (c % 4 != 0 || !oracle_vnni_available()) return NULL;
    vnni_pack_t* pk = (vnni_pack_t*)calloc(1, sizeof(vnni_pack_t));
    pk->k = k; pk->c = c; pk->kh = kernel_h; pk->kw = kernel_w;
    pk->taps = kernel_h * kernel_w; pk->cq = c / 4; pk->kblocks = (k + KB - 1) / KB;
    const size_t bytes = (size_t)pk->kblocks * pk->taps * pk->cq * KB * 4;
    pk->wp = (int8_t*)aligned_alloc(64, (bytes + 63) / 64 * 64);
    memset(pk->wp, 0, bytes);
    pk->wsum128 = (int32_t*)calloc((size_t)pk->kblocks * KB, sizeof(int32_t));
    const int taps = pk->taps, cq = pk->cq;
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < k; ++oc) {
        const int kb = oc / KB, j = oc % KB;
        int32_t s = 0;
        for (int ic = 0; ic < c; ++ic)
            for (int t = 0; t < taps; ++t) {
                const int8_t v = weights[((size_t)oc * c + ic) * taps + t];
                pk->wp[((((size_t)kb * taps + t) * cq + ic / 4) * KB + j) * 4 + (ic & 3)] = v;
                s += v;
            }
        pk->wsum128[oc] = 128 * s;
    }
    return pk;
}

ORACLE_API void oracle_vnni_free(void* p) {
    vnni_pack_t* pk = (vnni_pack_t*)p;
    if (!pk) return;
    free(pk->wp);
    free(pk->wsum128);
    free(pk);
}

/* 16 finished lanes of one pixel: the epilogue of oracle_conv_s8_nhwc_x86, vectorised. acc: s32 sums (shift already
 * removed); off: index of the first of the 16 output elements; lanes: how many of them exist (k tail). */
static inline void finish16(__m512i acc, __m512 bias, __m512 scale, const void* residual, int res_dtype, float sum_scale,
                            int flag_relu, void* dst, int dst_dtype, size_t off, __mmask16 lanes) {
    const __m512 zero = _mm512_setzero_ps();
    __m512 f = _mm512_cvtepi32_ps(acc);                 /* (float)acc, round to nearest even */
    f = _mm512_add_ps(f, bias);
    f = _mm512_mul_ps(f, scale);
    if (residual) {
        __m512 r;
        if (res_dtype == DT_FLOAT) r = _mm512_maskz_loadu_ps(lanes, (const float*)residual + off);
        else if (res_dtype == DT_UINT8) r = _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(lanes, (const uint8_t*)residual + off)));
        else r = _mm512_cvtepi32_ps(_mm512_cvtepi8_epi32(_mm_maskz_loadu_epi8(lanes, (const int8_t*)residual + off)));
        f = (sum_scale == 1.f) ? _mm512_add_ps(f, r) : _mm512_fmadd_ps(r, _mm512_set1_ps(sum_scale), f);
    }
    if (flag_relu) f = _mm512_max_ps(f, zero);          /* f > 0 ? f : +0 */
    if (dst_dtype == DT_FLOAT) {
        _mm512_mask_storeu_ps((float*)dst + off, lanes, f);
    } else if (dst_dtype == DT_INT8) {
        f = _mm512_min_ps(_mm512_max_ps(f, _mm512_set1_ps(-128.f)), _mm512_set1_ps(127.f));
        const __m512i q = _mm512_cvt_roundps_epi32(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
        _mm_mask_storeu_epi8((int8_t*)dst + off, lanes, _mm512_cvtepi32_epi8(q));
    } else {
        f = _mm512_min_ps(_mm512_max_ps(f, zero), _mm512_set1_ps(255.f));
        const __m512i q = _mm512_cvt_roundps_epi32(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
        _mm_mask_storeu_epi8((uint8_t*)dst + off, lanes, _mm512_cvtepi32_epi8(q));
    }
}

/* group == 1, weights packed by oracle_vnni_pack. Returns 0 on success, 1 when the shape does not match the pack. */
ORACLE_API int oracle_conv_s8_nhwc_x86_vnni_packed(const void* packed, const void* src, int src_dtype, const float* bias_f,
                                                   const float* scale, const void* residual, int res_dtype, float sum_scale,
                                                   void* dst, int dst_dtype, int n, int c, int h, int w, int k, int kernel_h,
                                                   int kernel_w, int stride_h, int stride_w, int dil_h, int dil_w, int pad_h,
                                                   int pad_w, int flag_relu) {
    const vnni_pack_t* pk = (const vnni_pack_t*)packed;
    if (!pk || pk->k != k || pk->c != c || pk->kh != kernel_h || pk->kw != kernel_w) return 1;
    const int out_h = (h + 2 * pad_h - (dil_h * (kernel_h - 1) + 1)) / stride_h + 1;
    const int out_w = (w + 2 * pad_w - (dil_w * (kernel_w - 1) + 1)) / stride_w + 1;
    const int taps = pk->taps, cq = pk->cq, kblocks = pk->kblocks, kpad = kblocks * KB;
    const int in_signed = src_dtype != DT_UINT8;
    /* per-channel tables padded to whole register tiles */
    float* tb = (float*)aligned_alloc(64, (size_t)kpad * 2 * sizeof(float));
    float* ts = tb + kpad;
    for (int i = 0; i < kpad; ++i) {
        tb[i] = (bias_f && i < k) ? bias_f[i] : 0.f;
        ts[i] = (scale && i < k) ? scale[i] : 1.f;
    }
    /* a zero pixel for padding taps; with a signed input the shifted zero is 128 */
    uint8_t* zpix = (uint8_t*)aligned_alloc(64, (size_t)((c + 63) / 64 * 64));
    memset(zpix, in_signed ? 128 : 0, (size_t)((c + 63) / 64 * 64));
    /* u8 view of the input (s8 + 128): every tap then contributes (x + 128) * w, padding taps (x = 0 -> 128) included,
     * so the correction 128 * sum(w) is the same for every output pixel */
    const uint8_t* xin = (const uint8_t*)src;
    uint8_t* shifted = NULL;
    const size_t in_elems = (size_t)n * h * w * c;
    if (in_signed) {
        shifted = (uint8_t*)aligned_alloc(64, (in_elems + 63) / 64 * 64);
        const size_t nv = in_elems / 64;
        const __m512i flip = _mm512_set1_epi8((char)0x80);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nv; ++i)
            _mm512_store_si512((void*)(shifted + 64 * i), _mm512_xor_si512(_mm512_loadu_si512((const void*)((const uint8_t*)src + 64 * i)), flip));
        for (size_t i = nv * 64; i < in_elems; ++i) shifted[i] = (uint8_t)(((const uint8_t*)src)[i] ^ 0x80u);
        xin = shifted;
    }
    const long long M = (long long)n * out_h * out_w;
    const long long mtiles = (M + PT - 1) / PT;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int kb = 0; kb < kblocks; ++kb) {
        for (long long mt = 0; mt < mtiles; ++mt) {
            const long long m0 = mt * PT;
            const int np = (int)(M - m0 < PT ? M - m0 : PT);
            int pn[PT], poh[PT], pow_[PT];
            for (int p = 0; p < PT; ++p) {
                const long long m = m0 + (p < np ? p : 0);
                pn[p] = (int)(m / ((long long)out_h * out_w));
                const int rem = (int)(m - (long long)pn[p] * out_h * out_w);
                poh[p] = rem / out_w;
                pow_[p] = rem - poh[p] * out_w;
            }
            __m512i acc[PT][2];
            for (int p = 0; p < PT; ++p) acc[p][0] = acc[p][1] = _mm512_setzero_si512();
            const int8_t* wkb = pk->wp + (size_t)kb * taps * cq * KB * 4;
            for (int kh = 0; kh < kernel_h; ++kh) {
                for (int kw = 0; kw < kernel_w; ++kw) {
                    const uint8_t* px[PT];
                    for (int p = 0; p < PT; ++p) {
                        const int ih = poh[p] * stride_h - pad_h + kh * dil_h;
                        const int iw = pow_[p] * stride_w - pad_w + kw * dil_w;
                        const int ok = p < np && ih >= 0 && ih < h && iw >= 0 && iw < w;
                        px[p] = ok ? xin + (((size_t)pn[p] * h + ih) * w + iw) * c : zpix;
                    }
                    const int8_t* wt_p = wkb + (size_t)(kh * kernel_w + kw) * cq * KB * 4;
                    for (int q4 = 0; q4 < cq; ++q4) {
                        const __m512i w0 = _mm512_load_si512((const void*)(wt_p + (size_t)q4 * KB * 4));
                        const __m512i w1 = _mm512_load_si512((const void*)(wt_p + (size_t)q4 * KB * 4 + 64));
#pragma GCC unroll 14
                        for (int p = 0; p < PT; ++p) {
                            int32_t quad;
                            memcpy(&quad, px[p] + q4 * 4, 4);
                            const __m512i a = _mm512_set1_epi32(quad);
                            acc[p][0] = _mm512_dpbusd_epi32(acc[p][0], a, w0);
                            acc[p][1] = _mm512_dpbusd_epi32(acc[p][1], a, w1);
                        }
                    }
                }
            }
            const int oc0 = kb * KB;
            const __m512i ws0 = in_signed ? _mm512_loadu_si512((const void*)(pk->wsum128 + oc0)) : _mm512_setzero_si512();
            const __m512i ws1 = in_signed ? _mm512_loadu_si512((const void*)(pk->wsum128 + oc0 + 16)) : _mm512_setzero_si512();
            const __m512 b0 = _mm512_load_ps(tb + oc0), b1 = _mm512_load_ps(tb + oc0 + 16);
            const __m512 s0 = _mm512_load_ps(ts + oc0), s1 = _mm512_load_ps(ts + oc0 + 16);
            const int left = k - oc0;
            const __mmask16 l0 = left >= 16 ? (__mmask16)0xFFFF : (__mmask16)((1u << left) - 1u);
            const __mmask16 l1 = left >= 32 ? (__mmask16)0xFFFF : (left > 16 ? (__mmask16)((1u << (left - 16)) - 1u) : (__mmask16)0);
            for (int p = 0; p < np; ++p) {
                const size_t off = (size_t)(m0 + p) * k + oc0;
                finish16(_mm512_sub_epi32(acc[p][0], ws0), b0, s0, residual, res_dtype, sum_scale, flag_relu, dst, dst_dtype, off, l0);
                if (l1) finish16(_mm512_sub_epi32(acc[p][1], ws1), b1, s1, residual, res_dtype, sum_scale, flag_relu, dst, dst_dtype, off + 16, l1);
            }
        }
    }
    free(tb);
    free(zpix);
    free(shifted);
    return 0;
}

/* Pack + run + free in one call (tests; the baseline arm keeps the packs). */
ORACLE_API int oracle_conv_s8_nhwc_x86_vnni(const void* src, int src_dtype, const int8_t* weights, const float* bias_f,
                                            const float* scale, const void* residual, int res_dtype, float sum_scale,
                                            void* dst, int dst_dtype, int n, int c, int h, int w, int k, int kernel_h,
                                            int kernel_w, int stride_h, int stride_w, int dil_h, int dil_w, int pad_h,
                                            int pad_w, int flag_relu) {
    void* pk = oracle_vnni_pack(weights, k, c, kernel_h, kernel_w);
    if (!pk) return 1;
    const int rc = oracle_conv_s8_nhwc_x86_vnni_packed(pk, src, src_dtype, bias_f, scale, residual, res_dtype, sum_scale, dst,
                                                       dst_dtype, n, c, h, w, k, kernel_h, kernel_w, stride_h, stride_w, dil_h,
                                                       dil_w, pad_h, pad_w, flag_relu);
    oracle_vnni_free(pk);
    return rc;
}

/* ------------------------------------------------------------------ pooling for the baseline arm
 * oracle_pool_s8_nhwc (oracle.c) 64 channels at a time. MAX: the maximum of the raw codes (the float compare of
 * integer-valued floats picks the same element; nearbyintf and the clamp are identities on it). AVG: the scalar code sums
 * the codes as floats -- every partial sum is an integer below 2^24, so an s32 sum converts to the same float -- then
 * divides (one IEEE division) and rounds to nearest even; here 16 lanes at a time. Window clipping as in oracle.c. */
ORACLE_API void oracle_pool_s8_nhwc_fast(const void* src, void* dst, int is_unsigned, int n, int c, int in_h, int in_w,
                                         int out_h, int out_w, int window_h, int window_w, int pad_h, int pad_w,
                                         int stride_h, int stride_w, int type) {
    const uint8_t* s = (const uint8_t*)src;
    uint8_t* d = (uint8_t*)dst;
#pragma omp parallel for collapse(2) schedule(static)
    for (int in_ = 0; in_ < n; ++in_) {
        for (int oh = 0; oh < out_h; ++oh) {
            int sh = oh * stride_h, eh = sh + window_h;
            if (pad_h > 0) {
                sh = (sh - pad_h) < 0 ? 0 : sh - pad_h;
                eh = (eh - pad_h) > in_h ? in_h : eh - pad_h;
            }
            if (eh > in_h) eh = in_h;
            for (int ow = 0; ow < out_w; ++ow) {
                int sw = ow * stride_w, ew = sw + window_w;
                if (pad_w > 0) {
                    sw = (sw - pad_w) < 0 ? 0 : sw - pad_w;
                    ew = (ew - pad_w) > in_w ? in_w : ew - pad_w;
                }
                if (ew > in_w) ew = in_w;
                uint8_t* out_px = d + (((size_t)in_ * out_h + oh) * out_w + ow) * c;
                if (type == 1) {
                    for (int c0 = 0; c0 < c; c0 += 64) {
                        const int left = c - c0;
                        const __mmask64 m = left >= 64 ? ~(__mmask64)0 : (((__mmask64)1 << left) - 1);
                        __m512i best = _mm512_maskz_loadu_epi8(m, s + (((size_t)in_ * in_h + sh) * in_w + sw) * c + c0);
                        for (int kh = sh; kh < eh; ++kh)
                            for (int kw = sw; kw < ew; ++kw) {
                                const __m512i v = _mm512_maskz_loadu_epi8(m, s + (((size_t)in_ * in_h + kh) * in_w + kw) * c + c0);
                                best = is_unsigned ? _mm512_max_epu8(best, v) : _mm512_max_epi8(best, v);
                            }
                        _mm512_mask_storeu_epi8(out_px + c0, m, best);
                    }
                } else {
                    const float div = type == 2 ? (float)(window_h * window_w) : (float)((ew - sw) * (eh - sh));
                    const __m512 vdiv = _mm512_set1_ps(div);
                    for (int c0 = 0; c0 < c; c0 += 16) {
                        const int left = c - c0;
                        const __mmask16 m = left >= 16 ? (__mmask16)0xFFFF : (__mmask16)((1u << left) - 1u);
                        __m512i sum = _mm512_setzero_si512();
                        for (int kh = sh; kh < eh; ++kh)
                            for (int kw = sw; kw < ew; ++kw) {
                                const __m128i b = _mm_maskz_loadu_epi8(m, s + (((size_t)in_ * in_h + kh) * in_w + kw) * c + c0);
                                sum = _mm512_add_epi32(sum, is_unsigned ? _mm512_cvtepu8_epi32(b) : _mm512_cvtepi8_epi32(b));
                            }
                        __m512 f = _mm512_div_ps(_mm512_cvtepi32_ps(sum), vdiv);
                        f = is_unsigned ? _mm512_min_ps(_mm512_max_ps(f, _mm512_setzero_ps()), _mm512_set1_ps(255.f))
                                        : _mm512_min_ps(_mm512_max_ps(f, _mm512_set1_ps(-128.f)), _mm512_set1_ps(127.f));
                        const __m512i q = _mm512_cvt_roundps_epi32(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
                        _mm_mask_storeu_epi8(out_px + c0, m, _mm512_cvtepi32_epi8(q));
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ fp32 convolution for the baseline arm
 * oracle_conv_f32_nhwc (oracle.c; group == 1) with AVX-512: the same sums -- residual * beta, the taps in (kh, kw, ic)
 * order, bias, relu(neg_slope) -- with one fused multiply-add per product instead of the scalar code's separately rounded
 * multiply and its width-dependent `omp simd` reduction tree. The two differ by float re-association / contraction only
 * (tests/test_cpu_oracle.py bounds the difference with the reference's tensor_cmp criterion at 1e-5); goldens and parity
 * tests keep using oracle.c, this one is timed. Weights packed once per layer into [k / 32][tap][c][32]. */
typedef struct {
    int k, c, kh, kw, kblocks, taps;
    float* wp;          /* [kblocks][taps][c][KB] */
} f32_pack_t;

ORACLE_API void* oracle_f32_pack(const float* weights, int k, int c, int kernel_h, int kernel_w) {
    if (!oracle_vnni_available()) return NULL;
    f32_pack_t* pk = (f32_pack_t*)calloc(1, sizeof(f32_pack_t));
    pk->k = k; pk->c = c; pk->kh = kernel_h; pk->kw = kernel_w;
    pk->taps = kernel_h * kernel_w; pk->kblocks = (k + KB - 1) / KB;
    const size_t elems = (size_t)pk->kblocks * pk->taps * c * KB;
    pk->wp = (float*)aligned_alloc(64, elems * sizeof(float));
    memset(pk->wp, 0, elems * sizeof(float));
    const int taps = pk->taps;
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < k; ++oc) {
        const int kb = oc / KB, j = oc % KB;
        for (int ic = 0; ic < c; ++ic)
            for (int t = 0; t < taps; ++t)
                pk->wp[(((size_t)kb * taps + t) * c + ic) * KB + j] = weights[((size_t)oc * c + ic) * taps + t];
    }
    return pk;
}

ORACLE_API void oracle_f32_free(void* p) {
    f32_pack_t* pk = (f32_pack_t*)p;
    if (!pk) return;
    free(pk->wp);
    free(pk);
}

ORACLE_API int oracle_conv_f32_nhwc_packed(const void* packed, const float* src, const float* bias, const float* residual,
                                           float* dst, int n, int c, int h, int w, int k, int kernel_h, int kernel_w,
                                           int stride_h, int stride_w, int dil_h, int dil_w, int pad_h, int pad_w,
                                           int flag_bias, int flag_relu, float neg_slope, float beta) {
    const f32_pack_t* pk = (const f32_pack_t*)packed;
    if (!pk || pk->k != k || pk->c != c || pk->kh != kernel_h || pk->kw != kernel_w) return 1;
    const int out_h = (h + 2 * pad_h - (dil_h * (kernel_h - 1) + 1)) / stride_h + 1;
    const int out_w = (w + 2 * pad_w - (dil_w * (kernel_w - 1) + 1)) / stride_w + 1;
    const int taps = pk->taps, kblocks = pk->kblocks, kpad = kblocks * KB;
    float* tb = (float*)aligned_alloc(64, (size_t)kpad * sizeof(float));
    for (int i = 0; i < kpad; ++i) tb[i] = (flag_bias && bias && i < k) ? bias[i] : 0.f;
    float* zpix = (float*)aligned_alloc(64, ((size_t)c * sizeof(float) + 63) / 64 * 64);
    memset(zpix, 0, ((size_t)c * sizeof(float) + 63) / 64 * 64);
    const long long M = (long long)n * out_h * out_w;
    const long long mtiles = (M + PT - 1) / PT;
    const __m512 vbeta = _mm512_set1_ps(beta), vslope = _mm512_set1_ps(neg_slope), zero = _mm512_setzero_ps();
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int kb = 0; kb < kblocks; ++kb) {
        for (long long mt = 0; mt < mtiles; ++mt) {
            const long long m0 = mt * PT;
            const int np = (int)(M - m0 < PT ? M - m0 : PT);
            const int oc0 = kb * KB;
            const int left = k - oc0;
            const __mmask16 l0 = left >= 16 ? (__mmask16)0xFFFF : (__mmask16)((1u << left) - 1u);
            const __mmask16 l1 = left >= 32 ? (__mmask16)0xFFFF : (left > 16 ? (__mmask16)((1u << (left - 16)) - 1u) : (__mmask16)0);
            int pn[PT], poh[PT], pow_[PT];
            __m512 acc[PT][2];
            for (int p = 0; p < PT; ++p) {
                const long long m = m0 + (p < np ? p : 0);
                pn[p] = (int)(m / ((long long)out_h * out_w));
                const int rem = (int)(m - (long long)pn[p] * out_h * out_w);
                poh[p] = rem / out_w;
                pow_[p] = rem - poh[p] * out_w;
                if (residual && p < np) {
                    const float* rp = residual + (size_t)m * k + oc0;
                    acc[p][0] = _mm512_mul_ps(_mm512_maskz_loadu_ps(l0, rp), vbeta);
                    acc[p][1] = _mm512_mul_ps(_mm512_maskz_loadu_ps(l1, rp + 16), vbeta);
                } else {
                    acc[p][0] = acc[p][1] = zero;
                }
            }
            const float* wkb = pk->wp + (size_t)kb * taps * c * KB;
            for (int kh = 0; kh < kernel_h; ++kh) {
                for (int kw = 0; kw < kernel_w; ++kw) {
                    const float* px[PT];
                    for (int p = 0; p < PT; ++p) {
                        const int ih = poh[p] * stride_h - pad_h + kh * dil_h;
                        const int iw = pow_[p] * stride_w - pad_w + kw * dil_w;
                        const int ok = p < np && ih >= 0 && ih < h && iw >= 0 && iw < w;
                        px[p] = ok ? src + (((size_t)pn[p] * h + ih) * w + iw) * c : zpix;
                    }
                    const float* wt_p = wkb + (size_t)(kh * kernel_w + kw) * c * KB;
                    for (int ic = 0; ic < c; ++ic) {
                        const __m512 w0 = _mm512_load_ps(wt_p + (size_t)ic * KB);
                        const __m512 w1 = _mm512_load_ps(wt_p + (size_t)ic * KB + 16);
#pragma GCC unroll 14
                        for (int p = 0; p < PT; ++p) {
                            const __m512 a = _mm512_set1_ps(px[p][ic]);
                            acc[p][0] = _mm512_fmadd_ps(a, w0, acc[p][0]);
                            acc[p][1] = _mm512_fmadd_ps(a, w1, acc[p][1]);
                        }
                    }
                }
            }
            const __m512 b0 = _mm512_load_ps(tb + oc0), b1 = _mm512_load_ps(tb + oc0 + 16);
            for (int p = 0; p < np; ++p) {
                float* op = dst + (size_t)(m0 + p) * k + oc0;
                __m512 f0 = _mm512_add_ps(acc[p][0], b0), f1 = _mm512_add_ps(acc[p][1], b1);
                if (flag_relu) {
                    f0 = _mm512_mask_mul_ps(f0, _mm512_cmp_ps_mask(f0, zero, _CMP_LE_OQ), f0, vslope);   /* acc > 0 ? acc : acc * slope */
                    f1 = _mm512_mask_mul_ps(f1, _mm512_cmp_ps_mask(f1, zero, _CMP_LE_OQ), f1, vslope);
                }
                _mm512_mask_storeu_ps(op, l0, f0);
                if (l1) _mm512_mask_storeu_ps(op + 16, l1, f1);
            }
        }
    }
    free(tb);
    free(zpix);
    return 0;
}
