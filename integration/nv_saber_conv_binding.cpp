// The reference-side binding of INTEGRATION.md section 2, written against the REFERENCE'S OWN headers
// (saber/funcs/impl/impl_base.h, saber/saber_funcs_param.h, saber/core/{tensor,context}.h for target NV) and the C ABI
// of include/b200_saber.h -- the file a maintainer would add as saber/funcs/impl/cuda/saber_conv.cpp in place of the
// dispatchers over the closed SASS kernels (saber_conv.cpp:190-585). It is compiled, by
// tests/test_cpu_host.py::test_reference_side_binding_compiles_against_reference_headers, with
//     g++ -std=c++11 -DUSE_CUDA -DNVIDIA_GPU -I oracle/ref_config -I $REF -I $REF/saber ... -I include -c
// where /root/reference exists: the class template arity (ImplBase<TargetType, OpDtype, Param>), the Param field names,
// the Tensor / Context accessors and the SaberStatus / DataType enum values the binding relies on are the reference's,
// not this repo's mirror. Nothing here is part of the product build.
#include <vector>

#include "saber/funcs/impl/impl_base.h"
#include "saber/saber_funcs_param.h"

#include "b200_saber.h"

namespace anakin {
namespace saber {

// enum values are passed through the C ABI unconverted: they must be the reference's
static_assert(static_cast<int>(SaberSuccess) == B200_SUCCESS, "SaberStatus values");
static_assert(static_cast<int>(SaberUnImplError) == B200_UNIMPL_ERROR, "SaberStatus values");
static_assert(static_cast<int>(SaberWrongDevice) == B200_WRONG_DEVICE, "SaberStatus values");
static_assert(static_cast<int>(AK_INT8) == B200_INT8 && static_cast<int>(AK_UINT8) == B200_UINT8 &&
              static_cast<int>(AK_FLOAT) == B200_FLOAT && static_cast<int>(AK_HALF) == B200_HALF, "DataType values");
static_assert(static_cast<int>(Pooling_max) == B200_POOL_MAX && static_cast<int>(Eltwise_sum) == B200_ELT_SUM &&
              static_cast<int>(Active_relu) == B200_ACT_RELU, "op enum values");

// the class the reference declares in saber/funcs/impl/cuda/saber_conv.h, with the plan in place of its SASS dispatchers
template <typename TargetType, DataType OpDtype>
class B200SaberConv2D : public ImplBase<TargetType, OpDtype, ConvParam<TargetType>> {
public:
    typedef std::vector<Tensor<TargetType>*> TensorVec;
    B200SaberConv2D() : _plan(nullptr), _packed_w(nullptr), _bias_f(nullptr), _scale(nullptr) {}
    ~B200SaberConv2D() { if (_plan) b200_conv_plan_destroy(_plan); }

    // device tables produced once by trans_weights (Conv::trans_weights, saber/funcs/conv.h:103-119):
    // b200_conv_pack_weights on the host, uploaded by the caller
    void set_tables(const void* packed_w, const float* bias_f, const float* scale) {
        _packed_w = packed_w; _bias_f = bias_f; _scale = scale;
    }

    SaberStatus init(const TensorVec& inputs, TensorVec& outputs, ConvParam<TargetType>& param,
                     Context<TargetType>& ctx) override {
        this->_ctx = &ctx;
        return create(inputs, outputs, param, ctx);
    }

    SaberStatus create(const TensorVec& inputs, TensorVec& outputs, ConvParam<TargetType>& param,
                       Context<TargetType>& ctx) override {
        this->_ctx = &ctx;
        b200_conv_desc_t d = b200_conv_desc_t();
        d.math = OpDtype == AK_INT8 ? B200_MATH_I8 : (OpDtype == AK_HALF ? B200_MATH_F16 : B200_MATH_TF32X3);
        d.in_dtype = inputs[0]->get_dtype();
        d.out_dtype = outputs[0]->get_dtype();
        d.res_dtype = -1;                                  // SaberConvEltwise: param.beta_type
        d.n = inputs[0]->num(); d.h = inputs[0]->height(); d.w = inputs[0]->width(); d.c = inputs[0]->channel();
        d.k = param.weight()->num();
        d.ldc = outputs[0]->channel();
        d.r = param.weight()->height(); d.s = param.weight()->width();
        d.pad_h = param.pad_h; d.pad_w = param.pad_w;
        d.stride_h = param.stride_h; d.stride_w = param.stride_w;
        d.dil_h = param.dilation_h; d.dil_w = param.dilation_w;
        d.relu = param.activation_param.has_active && param.activation_param.active == Active_relu;
        d.neg_slope = param.activation_param.negative_slope;
        d.sum_scale = 1.f;                                 // SaberConvEltwise: param.beta / s_out (b200_saber.h)
        if (_plan) { b200_conv_plan_destroy(_plan); _plan = nullptr; }
        return static_cast<SaberStatus>(b200_conv_plan_create(&d, _packed_w, _bias_f, _scale, &_plan));
    }

    SaberStatus dispatch(const TensorVec& inputs, TensorVec& outputs, ConvParam<TargetType>& param) override {
        (void)param;
        return static_cast<SaberStatus>(b200_conv_plan_run(_plan, inputs[0]->data(), nullptr, outputs[0]->mutable_data(),
                                                           this->_ctx->get_compute_stream()));
    }

private:
    b200_conv_plan_t* _plan;
    const void* _packed_w;
    const float* _bias_f;
    const float* _scale;
};

// pooling / softmax one-liners of the same shape (saber_pooling.cu:20-229, saber_softmax.cu:358-430)
template <typename TargetType>
SaberStatus b200_pooling_dispatch(const Tensor<TargetType>& in, Tensor<TargetType>& out, const PoolingParam<TargetType>& p,
                                  Context<TargetType>& ctx) {
    b200_pool_desc_t d = b200_pool_desc_t();
    d.dtype = in.get_dtype(); d.type = p.pooling_type;
    d.n = in.num(); d.h = in.height(); d.w = in.width(); d.c = in.channel();
    d.window_h = p.window_h; d.window_w = p.window_w; d.pad_h = p.pad_h; d.pad_w = p.pad_w;
    d.stride_h = p.stride_h; d.stride_w = p.stride_w;
    d.global_pooling = p.global_pooling ? 1 : 0;
    d.floor_as_conv = p.cmp_out_shape_floor_as_conv ? 1 : 0;
    return static_cast<SaberStatus>(b200_pool_run(&d, in.data(), out.mutable_data(), ctx.get_compute_stream()));
}

// instantiate against the reference's NV target: every member used above must exist with these names and types
template class B200SaberConv2D<NV, AK_INT8>;
template class B200SaberConv2D<NV, AK_FLOAT>;
template SaberStatus b200_pooling_dispatch<NV>(const Tensor<NV>&, Tensor<NV>&, const PoolingParam<NV>&, Context<NV>&);

}  // namespace saber
}  // namespace anakin
