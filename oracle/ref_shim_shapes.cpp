// ref_shim_shapes.cpp -- extern "C" wrapper around the REFERENCE's own Pooling<X86,AK_FLOAT>::compute_output_shape
// (saber/funcs/pooling.h:69-132: the ceil / floor-as-conv / "last window starts in the padding" rule), compiled
// from the header where it lies. -DANAKIN_SABER_FUNCS_IMPL_X86_SABER_POOLING_H pre-defines the include guard of
// the optimised x86 implementation (xbyak JIT, not vendored); the generic DEFINE_OP_CLASS stubs of
// impl/impl_pooling.h stand in for it -- only the shape function is called. TEST INFRASTRUCTURE.
#include "saber/funcs/pooling.h"

using namespace anakin::saber;

extern "C" int ref_pooling_output_shape(int n, int c, int h, int w, int window_h, int window_w, int pad_h, int pad_w,
                                        int stride_h, int stride_w, int global_pooling, int floor_as_conv,
                                        int* out_h, int* out_w) {
    Tensor<X86> in(Shape({n, c, h, w}, Layout_NCHW)), out;
    PoolingParam<X86> param(window_h, window_w, pad_h, pad_w, stride_h, stride_w, Pooling_max, global_pooling != 0,
                            floor_as_conv != 0);
    std::vector<Tensor<X86>*> ins{&in}, outs{&out};
    Pooling<X86, AK_FLOAT> op;
    SaberStatus st = op.compute_output_shape(ins, outs, param);
    *out_h = out.height();
    *out_w = out.width();
    return static_cast<int>(st);
}
