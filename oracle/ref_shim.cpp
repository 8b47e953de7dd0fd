// ref_shim.cpp -- extern "C" wrappers around the REFERENCE's own naive test oracles,
// compiled from the sources where they lie under /root/reference (never copied):
//   test/saber/conv_func_helper.h : conv_basic_check, conv_basic_check_int8,
//                                    pool_basic_check_int8
//   saber/core/tensor_op.cpp      : tensor_cmp_host
// Output goes to oracle/_ref/libanakin_ref_oracle.so (git-ignored). It is used by
// tests/ to pin oracle/oracle.c and, optionally, as a CPU baseline. TEST INFRASTRUCTURE.
//
// -DANAKIN_SABER_FUNCS_CONV_H pre-defines the include guard of saber/funcs/conv.h, which
// would otherwise pull in mkl-dnn / xbyak headers that are not vendored in the reference.
#include "test/saber/conv_func_helper.h"
#include "saber/core/tensor_op.h"

#include <cmath>
#include <limits>

using namespace anakin::saber;

// The naive CPU oracles the reference's op tests compare against live inside test .cpp files that also
// instantiate the (unbuildable) optimised ops. The Makefile cuts exactly those function templates out of the
// sources where they lie into _ref/ref_test_oracles.inc at build time (never committed, removed after the
// compile):  pooling_cpu_func  test_saber_pooling.cpp:14-104     fc_cpu_base  test_saber_fc.cpp:14-47
//            Count, softmax_cpu  test_saber_softmax.cpp:11-79     eltwise_cpu  test_saber_eltwise.cpp:16-98
//            activation_basic    test_saber_activation.cpp:16-137
#include "_ref/ref_test_oracles.inc"

static Tensor<X86>* wrap(void* data, DataType dt, LayoutType lt, int n, int c, int h, int w) {
    // Layout_NHWC shapes are given as (n,h,w,c) in Shape order.
    Shape sh = (lt == Layout_NHWC) ? Shape({n, h, w, c}, Layout_NHWC) : Shape({n, c, h, w}, Layout_NCHW);
    Tensor<X86>* t = new Tensor<X86>(data, X86(), 0, sh, dt);
    return t;
}

extern "C" {

void ref_conv_basic_check_f32(const float* src, const float* weights, const float* bias, float* dst,
                              int n, int c, int h, int w, int k, int oh, int ow, int group,
                              int kernel_w, int kernel_h, int stride_w, int stride_h, int dil_w,
                              int dil_h, int pad_w, int pad_h, int flag_bias, int flag_relu,
                              float beta, float alpha) {
    Tensor<X86>* tin = wrap(const_cast<float*>(src), AK_FLOAT, Layout_NCHW, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, n, k, oh, ow);
    conv_basic_check<X86, float, float>(*tin, *tout, weights, bias, group, kernel_w, kernel_h,
                                        stride_w, stride_h, dil_w, dil_h, pad_w, pad_h,
                                        flag_bias != 0, flag_relu != 0, beta, alpha);
    delete tin;
    delete tout;
}

void ref_conv_basic_check_int8(const void* src, int src_unsigned, const char* weights,
                               const int* bias, char* dst, int n, int c, int h, int w, int k, int oh,
                               int ow, int group, int kernel_w, int kernel_h, int stride_w,
                               int stride_h, int dil_w, int dil_h, int pad_w, int pad_h,
                               int flag_bias, int flag_relu, const float* scale, int has_elt_sum,
                               float sum_scale, float beta, int round_down) {
    Tensor<X86>* tin = wrap(const_cast<void*>(src), src_unsigned ? AK_UINT8 : AK_INT8, Layout_NHWC, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, AK_INT8, Layout_NHWC, n, k, oh, ow);
    std::vector<float> sc(scale, scale + k);
    EltwiseParam<X86> elt(Eltwise_sum, std::vector<float>({1.f, sum_scale}));
    conv_basic_check_int8<X86>(*tin, *tout, weights, bias, group, kernel_w, kernel_h, stride_w,
                               stride_h, dil_w, dil_h, pad_w, pad_h, flag_bias != 0, flag_relu != 0,
                               sc, has_elt_sum ? &elt : nullptr, beta, round_down ? down : nearest);
    delete tin;
    delete tout;
}

void ref_pool_basic_check_int8(const void* src, void* dst, int is_unsigned, int n, int c, int h,
                               int w, int oh, int ow, int kernel_w, int kernel_h, int stride_w,
                               int stride_h, int pad_w, int pad_h, int pooling_type) {
    DataType dt = is_unsigned ? AK_UINT8 : AK_INT8;
    Tensor<X86>* tin = wrap(const_cast<void*>(src), dt, Layout_NHWC, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, dt, Layout_NHWC, n, c, oh, ow);
    pool_basic_check_int8<X86>(*tin, *tout, kernel_w, kernel_h, stride_w, stride_h, pad_w, pad_h,
                               static_cast<PoolingType>(pooling_type));
    delete tin;
    delete tout;
}

void ref_pooling_cpu_f32(const float* src, float* dst, int n, int c, int h, int w, int oh, int ow, int window_h,
                         int window_w, int pad_h, int pad_w, int stride_h, int stride_w, int pooling_type) {
    Tensor<X86>* tin = wrap(const_cast<float*>(src), AK_FLOAT, Layout_NCHW, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, n, c, oh, ow);
    PoolingParam<X86> param(window_h, window_w, pad_h, pad_w, stride_h, stride_w,
                            static_cast<PoolingType>(pooling_type));
    std::vector<Tensor<X86>*> in{tin}, out{tout};
    pooling_cpu_func<float, X86, X86>(in, out, param);
    delete tin;
    delete tout;
}

void ref_fc_cpu_f32(const float* src, const float* weights, const float* bias, float* dst, int m, int k, int n_out) {
    Tensor<X86>* tin = wrap(const_cast<float*>(src), AK_FLOAT, Layout_NCHW, m, k, 1, 1);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, m, n_out, 1, 1);
    Tensor<X86>* tw = wrap(const_cast<float*>(weights), AK_FLOAT, Layout_NCHW, 1, 1, n_out, k);
    Tensor<X86>* tb = bias ? wrap(const_cast<float*>(bias), AK_FLOAT, Layout_NCHW, 1, 1, 1, n_out) : nullptr;
    FcParam<X86> param(tw, tb, n_out);
    std::vector<Tensor<X86>*> in{tin}, out{tout};
    fc_cpu_base<float, X86, X86>(in, out, param);
    delete tin;
    delete tout;
    delete tw;
    delete tb;
}

void ref_softmax_cpu_f32(const float* src, float* dst, int n, int c, int h, int w, int axis) {
    Tensor<X86>* tin = wrap(const_cast<float*>(src), AK_FLOAT, Layout_NCHW, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, n, c, h, w);
    SoftmaxParam<X86> param(axis);
    std::vector<Tensor<X86>*> in{tin}, out{tout};
    softmax_cpu<float, X86, X86>(in, out, param);
    delete tin;
    delete tout;
}

void ref_eltwise_cpu_f32(const float* a, const float* b, float* dst, int size, int op, float c0, float c1, int relu) {
    Tensor<X86>* ta = wrap(const_cast<float*>(a), AK_FLOAT, Layout_NCHW, 1, 1, 1, size);
    Tensor<X86>* tb = wrap(const_cast<float*>(b), AK_FLOAT, Layout_NCHW, 1, 1, 1, size);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, 1, 1, 1, size);
    ActivationParam<X86> act = relu ? ActivationParam<X86>(Active_relu) : ActivationParam<X86>();
    EltwiseParam<X86> param(static_cast<EltwiseType>(op), std::vector<float>({c0, c1}), act);
    std::vector<Tensor<X86>*> in{ta, tb}, out{tout};
    eltwise_cpu<float, X86, X86>(in, out, param);
    delete ta;
    delete tb;
    delete tout;
}

void ref_activation_f32(const float* src, float* dst, int n, int c, int h, int w, int act, float neg_slope, float coef) {
    Tensor<X86>* tin = wrap(const_cast<float*>(src), AK_FLOAT, Layout_NCHW, n, c, h, w);
    Tensor<X86>* tout = wrap(dst, AK_FLOAT, Layout_NCHW, n, c, h, w);
    ActivationParam<X86> param(static_cast<ActiveType>(act), neg_slope, coef);
    std::vector<Tensor<X86>*> in{tin}, out{tout};
    activation_basic<float, X86, X86>(in, out, param);
    delete tin;
    delete tout;
}

void ref_tensor_cmp_host(const float* a, const float* b, int size, double* max_ratio, double* max_diff) {
    tensor_cmp_host<float>(a, b, size, *max_ratio, *max_diff);
}

}  // extern "C"
