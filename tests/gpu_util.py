"""Helpers for the -m gpu parity tests: drive the C ABI (libb200saber.so) with torch
device buffers. torch is plumbing only (allocation, H2D/D2H, stream handle)."""
import ctypes as C

import numpy as np
import torch

from anakin_b200 import saber_abi as A

_NP2DT = {np.dtype(np.float32): A.FLOAT, np.dtype(np.float16): A.HALF, np.dtype(np.int8): A.INT8,
          np.dtype(np.uint8): A.UINT8}
_DT2TORCH = {A.FLOAT: torch.float32, A.HALF: torch.float16, A.INT8: torch.int8, A.UINT8: torch.uint8}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def pad_channels(x_nhwc, c_pad):
    n, h, w, c = x_nhwc.shape
    if c == c_pad:
        return np.ascontiguousarray(x_nhwc)
    out = np.zeros((n, h, w, c_pad), x_nhwc.dtype)
    out[..., :c] = x_nhwc
    return out


class ConvRunner:
    """One fused conv plan; mirrors SaberConv2D::init/create (plan) + dispatch (run)."""

    def __init__(self, math, x_shape_nhwc, in_dtype, w_kcrs, bias_f, scale, out_dtype, res_dtype=-1,
                 stride=(1, 1), pad=(0, 0), dil=(1, 1), relu=False, neg_slope=0.0, sum_scale=1.0,
                 ldc=None, fuse_pool=0, pool_stride=0, pool_pad=0, pool_floor_as_conv=False):
        lib = A.load()
        n, h, w, c = x_shape_nhwc
        k, c_real, r, s = w_kcrs.shape
        d = A.ConvDesc()
        d.math, d.in_dtype, d.out_dtype, d.res_dtype = math, in_dtype, out_dtype, res_dtype
        d.n, d.h, d.w, d.c, d.k = n, h, w, c, k
        d.ldc = ldc or k
        d.r, d.s = r, s
        d.pad_h, d.pad_w = pad
        d.stride_h, d.stride_w = stride
        d.dil_h, d.dil_w = dil
        d.relu, d.neg_slope, d.sum_scale = int(relu), neg_slope, sum_scale
        d.fuse_pool, d.pool_stride, d.pool_pad, d.pool_floor_as_conv = fuse_pool, pool_stride, pool_pad, int(pool_floor_as_conv)
        self.d = d
        ho, wo = C.c_int32(), C.c_int32()
        # (the size of what the plan stores: the pooled size with a fused pooling)
        A.check(lib.b200_conv_pooled_hw(C.byref(d), C.byref(ho), C.byref(wo)), "conv_pooled_hw")
        self.ho, self.wo = ho.value, wo.value
        nbytes = lib.b200_conv_packed_weight_bytes(C.byref(d))
        packed = np.zeros(nbytes, np.uint8)
        wsrc = np.ascontiguousarray(w_kcrs)
        A.check(lib.b200_conv_pack_weights(C.byref(d), wsrc.ctypes.data_as(C.c_void_p), c_real,
                                           packed.ctypes.data_as(C.c_void_p)), "pack_weights")
        self.w_dev = dev(packed)
        self.bias_dev = dev(np.asarray(bias_f, np.float32)) if bias_f is not None else None
        self.scale_dev = dev(np.asarray(scale, np.float32)) if scale is not None else None
        plan = C.c_void_p()
        A.check(lib.b200_conv_plan_create(C.byref(d), ptr(self.w_dev), ptr(self.bias_dev),
                                          ptr(self.scale_dev), C.byref(plan)), "conv_plan_create")
        self.plan = plan
        self.lib = lib

    def info(self):
        v = [C.c_int32() for _ in range(5)]
        self.lib.b200_conv_plan_info(self.plan, *[C.byref(x) for x in v])
        d = dict(zip(("block_n", "grid_x", "grid_y", "k_steps", "smem"), [x.value for x in v]))
        d["split"] = self.lib.b200_conv_plan_split(self.plan)
        d["slab"] = bool(self.lib.b200_conv_plan_is_slab(self.plan))
        d["persistent"] = bool(self.lib.b200_conv_plan_is_persistent(self.plan))
        return d

    def run(self, x_dev, res_dev=None, out_dev=None):
        d = self.d
        if out_dev is None:
            out_dev = torch.zeros((d.n, self.ho, self.wo, d.ldc), dtype=_DT2TORCH[d.out_dtype], device="cuda")
        A.check(self.lib.b200_conv_plan_run(self.plan, ptr(x_dev), ptr(res_dev), ptr(out_dev), stream_ptr()),
                "conv_plan_run")
        return out_dev

    def __del__(self):
        try:
            self.lib.b200_conv_plan_destroy(self.plan)
        except Exception:
            pass
