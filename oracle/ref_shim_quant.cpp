// ref_shim_quant.cpp -- extern "C" wrappers around the REFERENCE's own x86 INT8 quantisation helpers,
// utils::ScaleUtils in saber/funcs/impl/x86/x86_utils.h (header-only; compiled from where it lies):
//   scale_conv_weights_to_nchw_host :293-323  per-output-channel max|w|/127, truncating cast
//   scale_fp32_int8                 :325-347  roundf + clamp
//   scale_fp32_uint8                :360-372  scale * 127/255, truncating cast
// TEST INFRASTRUCTURE: pins oracle_quant_weights_per_oc / oracle_quant_fp32_s8 / oracle_quant_fp32_u8.
#include <cstring>
#include <vector>

#include "saber/funcs/impl/x86/x86_utils.h"

using namespace anakin::saber;

extern "C" void ref_quant_weights_per_oc(const float* w, int k, int c, int r, int s, signed char* out, float* scale) {
    Shape sh({k, c, r, s}, Layout_NCHW);
    Tensor<X86> tin(const_cast<float*>(w), X86(), 0, sh, AK_FLOAT);
    Tensor<X86> tout(sh, AK_INT8);
    utils::ScaleUtils::scale_conv_weights_to_nchw_host(tout, tin);
    std::memcpy(out, tout.data(), static_cast<size_t>(k) * c * r * s);
    std::vector<float> sc = tout.get_scale();
    for (int i = 0; i < k; ++i) scale[i] = sc[i];
}

extern "C" void ref_quant_fp32_s8(const float* x, int count, float scale, signed char* out) {
    Shape sh({1, 1, 1, count}, Layout_NCHW);
    Tensor<X86> tin(const_cast<float*>(x), X86(), 0, sh, AK_FLOAT);
    tin.set_scale({scale});
    Tensor<X86> tout(sh, AK_INT8);
    utils::ScaleUtils::scale_fp32_int8(tout, tin);
    std::memcpy(out, tout.data(), count);
}

extern "C" void ref_quant_fp32_u8(const float* x, int count, float scale, unsigned char* out) {
    Shape sh({1, 1, 1, count}, Layout_NCHW);
    Tensor<X86> tin(const_cast<float*>(x), X86(), 0, sh, AK_FLOAT);
    tin.set_scale({scale});
    Tensor<X86> tout(sh, AK_UINT8);
    utils::ScaleUtils::scale_fp32_uint8(tout, tin);
    std::memcpy(out, tout.data(), count);
}
