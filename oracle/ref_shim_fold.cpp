// ref_shim_fold.cpp -- extern "C" wrapper around the REFERENCE's own BatchNorm / Scale folding,
// WeightsFusion<float, X86>::update_weights (framework/utils/parameter_fusion.cpp:86-131), compiled from the
// source where it lies. TEST INFRASTRUCTURE: pins oracle_fold_bn_scale and the product's fold (operators.cpp).
#include <cstring>
#include <vector>

#include "framework/utils/parameter_fusion.h"

using namespace anakin;
using namespace anakin::saber;

extern "C" void ref_fold_bn_scale(float* weights, float* bias, int n, int c, int h, int w, int conv_bias_term,
                                  float bn_scale_factor, float eps, const float* mean, const float* var,
                                  const float* gamma, const float* beta_s, int scale_bias_term) {
    Shape4d ws({n, c, h, w}), bs({1, n, 1, 1});
    PBlock<X86> wb(ws, AK_FLOAT), bb(bs, AK_FLOAT);
    std::memcpy(wb.h_tensor().mutable_data(), weights, sizeof(float) * n * c * h * w);
    if (conv_bias_term) std::memcpy(bb.h_tensor().mutable_data(), bias, sizeof(float) * n);
    std::vector<float> vm(mean, mean + n), vv(var, var + n), vg(gamma, gamma + n), vb(n, 0.f);
    if (beta_s) vb.assign(beta_s, beta_s + n);
    WeightsFusion<float, X86>::update_weights(wb, bb, n, c, h, w, conv_bias_term != 0, bn_scale_factor, eps, vm, vv,
                                              vg, vb, scale_bias_term != 0);
    std::memcpy(weights, wb.h_tensor().data(), sizeof(float) * n * c * h * w);
    std::memcpy(bias, bb.h_tensor().data(), sizeof(float) * n);
}
