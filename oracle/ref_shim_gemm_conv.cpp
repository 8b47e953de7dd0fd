// ref_shim_gemm_conv.cpp -- extern "C" wrapper around the REFERENCE's own x86 INT8 convolution
// GemmX8S8S32XConv (saber/funcs/impl/x86/gemm_x8s8s32x_conv.{h,cpp}), compiled from the sources where they lie:
// its init/create (weight quantisation through utils::ScaleUtils, goihw->hwigo reorder, bias pre-scaling, the
// input/output dtype scale table :145-182) and dispatch (u8 im2col, integer GEMM, epilogue :256-281) run verbatim.
// The one external routine it needs, MKL's cblas_gemm_s8u8s32, is pure integer arithmetic and is supplied below
// from its documented definition (see ref_config/mkl_cblas.h). TEST INFRASTRUCTURE: pins oracle_conv_s8_nhwc_x86.
#include <cstdint>
#include <cstring>
#include <vector>

#include "mkl_cblas.h"
#include "saber/funcs/impl/x86/gemm_x8s8s32x_conv.h"

extern "C" void cblas_gemm_s8u8s32(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa, const CBLAS_TRANSPOSE transb,
                                   const CBLAS_OFFSET offsetc, const MKL_INT m, const MKL_INT n, const MKL_INT k,
                                   const float alpha, const void* a, const MKL_INT lda, const char ao, const void* b,
                                   const MKL_INT ldb, const char bo, const float beta, int* c, const MKL_INT ldc,
                                   const int* co) {
    // column-major, no transposes (the only form the reference uses); alpha = 1, beta = 0 there, kept general
    const int8_t* A = static_cast<const int8_t*>(a);
    const uint8_t* B = static_cast<const uint8_t*>(b);
    if (layout != CblasColMajor || transa != CblasNoTrans || transb != CblasNoTrans) abort();
#pragma omp parallel for
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) {
            long long acc = 0;
            for (int p = 0; p < k; ++p)
                acc += (static_cast<int>(A[i + static_cast<size_t>(p) * lda]) + ao) *
                       (static_cast<int>(B[p + static_cast<size_t>(j) * ldb]) + bo);
            const int off = offsetc == CblasFixOffset ? co[0] : (offsetc == CblasColOffset ? co[i] : co[j]);
            const float prev = beta == 0.f ? 0.f : beta * c[i + static_cast<size_t>(j) * ldc];
            c[i + static_cast<size_t>(j) * ldc] = static_cast<int>(alpha * static_cast<float>(acc) + prev) + off;
        }
    }
}

using namespace anakin::saber;

// x: NHWC s8 | u8 [n,h,w,c]; w: fp32 KCRS; bias fp32 [k] or null; out: NHWC s8 | u8 | f32 [n,oh,ow,k]
extern "C" int ref_gemm_conv_int8(const void* x, int x_unsigned, float in_scale, const float* w, const float* bias,
                                  void* out, int out_dtype, float out_scale, int n, int h, int wd, int c, int k, int r,
                                  int s, int oh, int ow, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h,
                                  int dil_w, int relu) {
    Shape xs({n, h, wd, c}, Layout_NHWC), os({n, oh, ow, k}, Layout_NHWC), ws({k, c, r, s}, Layout_NCHW),
        bs({1, k, 1, 1}, Layout_NCHW);
    Tensor<X86> tin(const_cast<void*>(x), X86(), 0, xs, x_unsigned ? AK_UINT8 : AK_INT8);
    tin.set_scale({in_scale});
    Tensor<X86> tout(out, X86(), 0, os, static_cast<DataType>(out_dtype));
    if (out_dtype != AK_FLOAT) tout.set_scale({out_scale});
    Tensor<X86> tw(const_cast<float*>(w), X86(), 0, ws, AK_FLOAT);
    Tensor<X86> tb(const_cast<float*>(bias ? bias : w), X86(), 0, bs, AK_FLOAT);   // unused without a bias
    ActivationParam<X86> act = relu ? ActivationParam<X86>(Active_relu) : ActivationParam<X86>();
    ConvParam<X86> cp(1, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, &tw, bias ? &tb : nullptr, act);
    EltwiseParam<X86> ep(Eltwise_sum);
    ConvEltwiseParam<X86> param(cp, ep);
    Env<X86>::env_init();
    Context<X86> ctx(0, 0, 0);
    std::vector<Tensor<X86>*> ins{&tin}, outs{&tout};
    GemmX8S8S32XConv op;
    SaberStatus st = op.init(ins, outs, param, ctx);
    if (st != SaberSuccess) return static_cast<int>(st);
    return static_cast<int>(op.dispatch(ins, outs, param));
}
