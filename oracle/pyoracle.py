"""ctypes front-end of the CPU oracle (oracle/oracle.c) and, when built, of the
reference's own naive test oracle compiled from /root/reference (oracle/_ref/).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by anakin_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libanakin_ref_oracle.so")
_REF_SHAPES = os.path.join(_HERE, "_ref", "libanakin_ref_shapes.so")

DT_FLOAT, DT_INT8, DT_UINT8 = 1, 3, 7


def build(ref=True):
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_vnni.c", "Makefile")]
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(_LIB) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
    if ref and os.path.isdir("/root/reference/test/saber") and not all(os.path.exists(os.path.join(_HERE, "_ref", n)) for n in
                    ("libanakin_ref_oracle.so", "libanakin_ref_shapes.so", "libanakin_ref_fold.so",
                     "libanakin_ref_quant.so", "libanakin_ref_gemmconv.so")):
        subprocess.check_call(["make", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build(ref=False)
        # idle OpenMP threads must sleep, not spin: on a CPU-quota'd container spinning waiters starve the workers
        # (measured here: 64 ms instead of 1 ms for a small parallel loop). Only effective when libgomp is not loaded yet.
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        _lib = C.CDLL(_LIB)
        _lib.oracle_conv_out_size.restype = C.c_int
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


_ref_shapes = None


def ref_pool_out_size(h, w, wh, ww, ph, pw, sh, sw, global_pooling=False, floor_as_conv=False):
    """The reference's own Pooling<X86,AK_FLOAT>::compute_output_shape (oracle/_ref); None when not built."""
    global _ref_shapes
    if _ref_shapes is None:
        if not os.path.exists(_REF_SHAPES):
            return None
        _ref_shapes = C.CDLL(_REF_SHAPES)
    oh, ow = C.c_int(), C.c_int()
    _ref_shapes.ref_pooling_output_shape(1, 1, h, w, wh, ww, ph, pw, sh, sw, int(global_pooling), int(floor_as_conv),
                                         C.byref(oh), C.byref(ow))
    return oh.value, ow.value


def ref_lib():
    """The reference's own oracle (None when it has not been / cannot be built here)."""
    global _ref
    if _ref is None and os.path.exists(_REF):
        _ref = C.CDLL(_REF)
    return _ref


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(x):
    return C.c_float(float(x))


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    lib().oracle_set_threads(int(n))


def conv_out_size(i, pad, dil, k, stride):
    return (i + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


def pool_out_size(h, w, wh, ww, ph, pw, sh, sw, global_pooling=False, floor_as_conv=False):
    oh, ow = C.c_int(), C.c_int()
    lib().oracle_pool_out_size(h, w, wh, ww, ph, pw, sh, sw, int(global_pooling), int(floor_as_conv),
                               C.byref(oh), C.byref(ow))
    return oh.value, ow.value


def conv_f32_nchw(x, w, bias, group=1, stride=(1, 1), dil=(1, 1), pad=(0, 0), relu=False,
                  neg_slope=0.0, beta=0.0, alpha=1.0, dst=None):
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
    out = np.zeros((n, k, oh, ow), np.float32) if dst is None else np.ascontiguousarray(dst, np.float32).copy()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    lib().oracle_conv_f32_nchw(_p(x), _p(w), _p(b), _p(out), n, c, h, wd, k, group, r, s, stride[0],
                               stride[1], dil[0], dil[1], pad[0], pad[1], int(b is not None),
                               int(relu), _f(neg_slope), _f(beta), _f(alpha))
    return out


_F32_PACKS = {}    # id(weight array) -> (the array, pack handle)


def _f32_pack(w):
    ent = _F32_PACKS.get(id(w))
    if ent is not None and ent[0] is w:
        return ent[1]
    if len(_F32_PACKS) >= 512:
        f32_release()
    k, c, r, s = w.shape
    lib().oracle_f32_pack.restype = C.c_void_p
    handle = C.c_void_p(lib().oracle_f32_pack(_p(w), k, c, r, s))
    assert handle.value, "oracle_f32_pack failed"
    _F32_PACKS[id(w)] = (w, handle)
    return handle


def f32_release():
    for _, handle in _F32_PACKS.values():
        lib().oracle_f32_free(handle)
    _F32_PACKS.clear()


def conv_f32_nhwc(x, w, bias, residual=None, group=1, stride=(1, 1), dil=(1, 1), pad=(0, 0),
                  relu=False, neg_slope=0.0, beta=1.0, fast=False):
    """fast=True (bench.py's CPU arm only): the AVX-512 implementation in oracle_vnni.c -- same sums with fused multiply-adds,
    equal to the scalar restatement up to float re-association (not bit-identical: goldens and parity tests use the default).
    The weight pack is cached per weight ARRAY."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    n, h, wd, c = x.shape
    k, _, r, s = w.shape
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
    out = np.zeros((n, oh, ow, k), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    res = None if residual is None else np.ascontiguousarray(residual, np.float32)
    if fast and group == 1 and vnni_available():
        rc = lib().oracle_conv_f32_nhwc_packed(_f32_pack(w), _p(x), _p(b), _p(res), _p(out), n, c, h, wd, k, r, s,
                                               stride[0], stride[1], dil[0], dil[1], pad[0], pad[1],
                                               int(b is not None), int(relu), _f(neg_slope), _f(beta))
        if rc == 0:
            return out
    lib().oracle_conv_f32_nhwc(_p(x), _p(w), _p(b), _p(res), _p(out), n, c, h, wd, k, group, r, s,
                               stride[0], stride[1], dil[0], dil[1], pad[0], pad[1],
                               int(b is not None), int(relu), _f(neg_slope), _f(beta))
    return out


_NP = {DT_FLOAT: np.float32, DT_INT8: np.int8, DT_UINT8: np.uint8}


def _dt(a):
    return {np.dtype(np.float32): DT_FLOAT, np.dtype(np.int8): DT_INT8, np.dtype(np.uint8): DT_UINT8}[a.dtype]


def vnni_available():
    return bool(lib().oracle_vnni_available())


_VNNI_PACKS = {}   # id(weight array) -> (the array itself: keeps the id alive, pack handle, padded channel count)


def _vnni_pack(w_s8):
    """The VNNI weight pack of a KCRS int8 array, made once per array (init-time work like the reference's trans_weights;
    the baseline arm keeps its quantised weights in a cache, so every layer is packed once)."""
    ent = _VNNI_PACKS.get(id(w_s8))
    if ent is not None and ent[0] is w_s8:
        return ent[1], ent[2]
    if len(_VNNI_PACKS) >= 512:      # callers that pass a fresh array every time (tests) must not grow this without bound
        vnni_release()
    k, c, r, s = w_s8.shape
    cp = (c + 3) // 4 * 4
    src = w_s8
    if cp != c:
        src = np.zeros((k, cp, r, s), np.int8); src[:, :c] = w_s8
    lib().oracle_vnni_pack.restype = C.c_void_p
    handle = C.c_void_p(lib().oracle_vnni_pack(_p(src), k, cp, r, s))
    assert handle.value, "oracle_vnni_pack failed"
    _VNNI_PACKS[id(w_s8)] = (w_s8, handle, cp)
    return handle, cp


def vnni_release():
    """Free every cached weight pack."""
    for _, handle, _ in _VNNI_PACKS.values():
        lib().oracle_vnni_free(handle)
    _VNNI_PACKS.clear()


def conv_s8_nhwc_x86(x, w_s8, bias_f, scale, residual=None, sum_scale=1.0, out_dtype=DT_INT8,
                     stride=(1, 1), dil=(1, 1), pad=(0, 0), relu=False, group=1, fast=False):
    """x: NHWC s8/u8, w_s8: KCRS int8 ([k][c/group][r][s]); returns NHWC out_dtype.
    fast=True: the AVX-512 VNNI implementation of the same arithmetic (oracle_vnni.c; bit-identical, used by the
    CPU-baseline arm of bench.py) where the CPU has it and group == 1; channels are zero-padded to a multiple of 4. The
    weight pack is cached per weight ARRAY (pass the same object again to reuse it)."""
    x = np.ascontiguousarray(x)
    w_s8 = np.ascontiguousarray(w_s8, np.int8)
    n, h, wd, c = x.shape
    k, cw, r, s = w_s8.shape
    assert cw * group == c and k % group == 0
    if fast and group == 1 and vnni_available():
        handle, cp = _vnni_pack(w_s8)
        if cp != c:
            xp = np.zeros((n, h, wd, cp), x.dtype); xp[..., :c] = x
            x, c = xp, cp
        oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
        ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
        out = np.empty((n, oh, ow, k), _NP[out_dtype])
        b = None if bias_f is None else np.ascontiguousarray(bias_f, np.float32)
        sc = None if scale is None else np.ascontiguousarray(scale, np.float32)
        res = None if residual is None else np.ascontiguousarray(residual)
        rc = lib().oracle_conv_s8_nhwc_x86_vnni_packed(handle, _p(x), _dt(x), _p(b), _p(sc), _p(res),
                                                       _dt(res) if res is not None else -1, _f(sum_scale), _p(out),
                                                       out_dtype, n, c, h, wd, k, r, s, stride[0], stride[1], dil[0],
                                                       dil[1], pad[0], pad[1], int(relu))
        if rc == 0:
            return out
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
    out = np.zeros((n, oh, ow, k), _NP[out_dtype])
    b = None if bias_f is None else np.ascontiguousarray(bias_f, np.float32)
    sc = None if scale is None else np.ascontiguousarray(scale, np.float32)
    res = None if residual is None else np.ascontiguousarray(residual)
    if group != 1:
        lib().oracle_conv_s8_nhwc_x86_group(_p(x), _dt(x), _p(w_s8), _p(b), _p(sc), _p(res),
                                            _dt(res) if res is not None else -1, _f(sum_scale), _p(out),
                                            out_dtype, n, c, h, wd, k, group, r, s, stride[0], stride[1], dil[0],
                                            dil[1], pad[0], pad[1], int(relu))
        return out
    lib().oracle_conv_s8_nhwc_x86(_p(x), _dt(x), _p(w_s8), _p(b), _p(sc), _p(res),
                                  _dt(res) if res is not None else -1, _f(sum_scale), _p(out),
                                  out_dtype, n, c, h, wd, k, r, s, stride[0], stride[1], dil[0],
                                  dil[1], pad[0], pad[1], int(relu))
    return out


def conv_s8_nhwc_basic(x, w_s8, bias_i32, scale, dst=None, has_elt_sum=False, sum_scale=1.0,
                       beta=0.0, group=1, stride=(1, 1), dil=(1, 1), pad=(0, 0), relu=False,
                       round_down=False, use_ref=False):
    """Restatement (or, use_ref=True, the reference itself) of conv_basic_check_int8."""
    x = np.ascontiguousarray(x)
    w_s8 = np.ascontiguousarray(w_s8, np.int8)
    n, h, wd, c = x.shape
    k, _, r, s = w_s8.shape
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
    out = np.zeros((n, oh, ow, k), np.int8) if dst is None else np.ascontiguousarray(dst, np.int8).copy()
    b = None if bias_i32 is None else np.ascontiguousarray(bias_i32, np.int32)
    sc = np.ascontiguousarray(scale, np.float32)
    uns = int(x.dtype == np.uint8)
    if use_ref:
        ref_lib().ref_conv_basic_check_int8(_p(x), uns, _p(w_s8), _p(b), _p(out), n, c, h, wd, k, oh,
                                            ow, group, s, r, stride[1], stride[0], dil[1], dil[0],
                                            pad[1], pad[0], int(b is not None), int(relu), _p(sc),
                                            int(has_elt_sum), _f(sum_scale), _f(beta), int(round_down))
    else:
        lib().oracle_conv_s8_nhwc_basic(_p(x), uns, _p(w_s8), _p(b), _p(out), n, c, h, wd, k, group,
                                        r, s, stride[0], stride[1], dil[0], dil[1], pad[0], pad[1],
                                        int(b is not None), int(relu), _p(sc), int(has_elt_sum),
                                        _f(sum_scale), _f(beta), int(round_down))
    return out


def ref_conv_f32_nchw(x, w, bias, group=1, stride=(1, 1), dil=(1, 1), pad=(0, 0), relu=False,
                      beta=0.0, alpha=1.0, dst=None):
    """The reference's conv_basic_check itself (oracle/_ref)."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s, stride[1])
    out = np.zeros((n, k, oh, ow), np.float32) if dst is None else np.ascontiguousarray(dst, np.float32).copy()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    ref_lib().ref_conv_basic_check_f32(_p(x), _p(w), _p(b), _p(out), n, c, h, wd, k, oh, ow, group, s,
                                       r, stride[1], stride[0], dil[1], dil[0], pad[1], pad[0],
                                       int(b is not None), int(relu), _f(beta), _f(alpha))
    return out


def int8_conv_scales(w_scale, bias, in_scale, in_dtype, out_scale, out_dtype, res_scale=1.0,
                     res_dtype=DT_INT8):
    w_scale = np.ascontiguousarray(w_scale, np.float32)
    k = w_scale.shape[0]
    sc = np.zeros(k, np.float32)
    bf = np.zeros(k, np.float32)
    ss = C.c_float()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    lib().oracle_int8_conv_scales(_p(w_scale), _p(b), k, _f(in_scale), in_dtype, _f(out_scale),
                                  out_dtype, _f(res_scale), res_dtype, _p(sc), _p(bf), C.byref(ss))
    return sc, bf, ss.value


def quant_weights_per_oc(w):
    w = np.ascontiguousarray(w, np.float32)
    k = w.shape[0]
    per = w.size // k
    out = np.zeros(w.shape, np.int8)
    sc = np.zeros(k, np.float32)
    lib().oracle_quant_weights_per_oc(_p(w), k, per, _p(out), _p(sc))
    return out, sc


def quant_fp32_s8(x, scale):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.int8)
    lib().oracle_quant_fp32_s8(_p(x), _p(out), C.c_size_t(x.size), _f(scale))
    return out


def quant_fp32_u8(x, scale):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.uint8)
    lib().oracle_quant_fp32_u8(_p(x), _p(out), C.c_size_t(x.size), _f(scale))
    return out


def fold_bn_scale(w, bias, bn_scale_factor, eps, mean, var, gamma, beta_s):
    w = np.ascontiguousarray(w, np.float32).copy()
    k = w.shape[0]
    b = np.zeros(k, np.float32) if bias is None else np.ascontiguousarray(bias, np.float32).copy()
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    mean, var, gamma, beta_s = f32(mean), f32(var), f32(gamma), f32(beta_s)
    lib().oracle_fold_bn_scale(_p(w), _p(b), k, w.size // k, _f(bn_scale_factor), _f(eps), _p(mean),
                               _p(var), _p(gamma), _p(beta_s))
    return w, b


_REF_QUANT = os.path.join(_HERE, "_ref", "libanakin_ref_quant.so")
_ref_quant = None


def _ref_quant_lib():
    global _ref_quant
    if _ref_quant is None and os.path.exists(_REF_QUANT):
        _ref_quant = C.CDLL(_REF_QUANT)
    return _ref_quant


def ref_quant_weights_per_oc(w):
    """utils::ScaleUtils::scale_conv_weights_to_nchw_host of the reference (oracle/_ref); None when not built."""
    L = _ref_quant_lib()
    if L is None:
        return None
    w = np.ascontiguousarray(w, np.float32)
    k, c, r, s_ = w.shape
    out = np.zeros(w.shape, np.int8)
    sc = np.zeros(k, np.float32)
    L.ref_quant_weights_per_oc(_p(w), k, c, r, s_, _p(out), _p(sc))
    return out, sc


def ref_quant_fp32(x, scale, unsigned=False):
    """scale_fp32_int8 / scale_fp32_uint8 of the reference (oracle/_ref); None when not built."""
    L = _ref_quant_lib()
    if L is None:
        return None
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.uint8 if unsigned else np.int8)
    (L.ref_quant_fp32_u8 if unsigned else L.ref_quant_fp32_s8)(_p(x), x.size, _f(scale), _p(out))
    return out


_REF_GEMMCONV = os.path.join(_HERE, "_ref", "libanakin_ref_gemmconv.so")
_ref_gemmconv = None


def ref_gemm_conv_int8(x, w_f32, bias, in_scale, out_dtype, out_scale, stride=(1, 1), pad=(0, 0), dil=(1, 1),
                       relu=False):
    """The reference's own x86 INT8 convolution GemmX8S8S32XConv (init + dispatch, oracle/_ref): x NHWC s8|u8,
    fp32 KCRS weights quantised by the reference itself. None when not built."""
    global _ref_gemmconv
    if _ref_gemmconv is None:
        if not os.path.exists(_REF_GEMMCONV):
            return None
        _ref_gemmconv = C.CDLL(_REF_GEMMCONV)
    x = np.ascontiguousarray(x)
    w_f32 = np.ascontiguousarray(w_f32, np.float32)
    n, h, wd, c = x.shape
    k, _, r, s_ = w_f32.shape
    oh = conv_out_size(h, pad[0], dil[0], r, stride[0])
    ow = conv_out_size(wd, pad[1], dil[1], s_, stride[1])
    out = np.zeros((n, oh, ow, k), _NP[out_dtype])
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    st = _ref_gemmconv.ref_gemm_conv_int8(_p(x), int(x.dtype == np.uint8), _f(in_scale), _p(w_f32), _p(b), _p(out),
                                          out_dtype, _f(out_scale), n, h, wd, c, k, r, s_, oh, ow, stride[0],
                                          stride[1], pad[0], pad[1], dil[0], dil[1], int(relu))
    if st != -1:
        raise RuntimeError("reference GemmX8S8S32XConv returned SaberStatus %d" % st)
    return out


_REF_FOLD = os.path.join(_HERE, "_ref", "libanakin_ref_fold.so")
_ref_fold = None


def ref_fold_bn_scale(w, bias, bn_scale_factor, eps, mean, var, gamma, beta_s):
    """The reference's own WeightsFusion<float,X86>::update_weights (oracle/_ref); None when not built."""
    global _ref_fold
    if _ref_fold is None:
        if not os.path.exists(_REF_FOLD):
            return None
        _ref_fold = C.CDLL(_REF_FOLD)
    w = np.ascontiguousarray(w, np.float32).copy()
    k, c, r, s_ = w.shape
    b = np.zeros(k, np.float32) if bias is None else np.ascontiguousarray(bias, np.float32).copy()
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    mean, var, gamma, beta_s = f32(mean), f32(var), f32(gamma), f32(beta_s)
    _ref_fold.ref_fold_bn_scale(_p(w), _p(b), k, c, r, s_, int(bias is not None), _f(bn_scale_factor), _f(eps),
                                _p(mean), _p(var), _p(gamma), _p(beta_s), int(beta_s is not None))
    return w, b


def pool_f32(x, window, pad, stride, ptype, nhwc=False, global_pooling=False, floor_as_conv=False):
    x = np.ascontiguousarray(x, np.float32)
    if nhwc:
        n, h, w, c = x.shape
    else:
        n, c, h, w = x.shape
    if global_pooling:
        window, stride, pad = (h, w), (h, w), (0, 0)
    oh, ow = pool_out_size(h, w, window[0], window[1], pad[0], pad[1], stride[0], stride[1],
                           global_pooling, floor_as_conv)
    out = np.zeros((n, oh, ow, c) if nhwc else (n, c, oh, ow), np.float32)
    lib().oracle_pool_f32(_p(x), _p(out), n, c, h, w, oh, ow, window[0], window[1], pad[0], pad[1],
                          stride[0], stride[1], int(ptype), int(nhwc))
    return out


def pool_s8_nhwc(x, window, pad, stride, ptype, global_pooling=False, floor_as_conv=False,
                 use_ref=False, fast=False):
    """fast=True: the AVX-512 implementation of the same arithmetic (oracle_vnni.c, bit-identical; bench.py's CPU arm)."""
    x = np.ascontiguousarray(x)
    n, h, w, c = x.shape
    if global_pooling:
        window, stride, pad = (h, w), (h, w), (0, 0)
    oh, ow = pool_out_size(h, w, window[0], window[1], pad[0], pad[1], stride[0], stride[1],
                           global_pooling, floor_as_conv)
    out = np.zeros((n, oh, ow, c), x.dtype)
    uns = int(x.dtype == np.uint8)
    if use_ref:
        ref_lib().ref_pool_basic_check_int8(_p(x), _p(out), uns, n, c, h, w, oh, ow, window[1],
                                            window[0], stride[1], stride[0], pad[1], pad[0], int(ptype))
    elif fast and vnni_available():
        lib().oracle_pool_s8_nhwc_fast(_p(x), _p(out), uns, n, c, h, w, oh, ow, window[0], window[1],
                                       pad[0], pad[1], stride[0], stride[1], int(ptype))
    else:
        lib().oracle_pool_s8_nhwc(_p(x), _p(out), uns, n, c, h, w, oh, ow, window[0], window[1],
                                  pad[0], pad[1], stride[0], stride[1], int(ptype))
    return out


def fc_f32(x, w, bias):
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    m = x.shape[0]
    k = x.size // m
    n_out = w.size // k
    out = np.zeros((m, n_out), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    lib().oracle_fc_f32(_p(x), _p(w), _p(b), _p(out), m, k, n_out)
    return out


def fc_s8(x, w_s8, bias, scale):
    x = np.ascontiguousarray(x)
    w_s8 = np.ascontiguousarray(w_s8, np.int8)
    m = x.shape[0]
    k = x.size // m
    n_out = w_s8.size // k
    out = np.zeros((m, n_out), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    sc = np.ascontiguousarray(scale, np.float32)
    lib().oracle_fc_s8(_p(x), _dt(x), _p(w_s8), _p(b), _p(sc), _p(out), m, k, n_out)
    return out


def softmax_f32(x, outer, axis_size, inner=1):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.float32)
    lib().oracle_softmax_f32(_p(x), _p(out), outer, axis_size, inner)
    return out


def eltwise_f32(a, b, op=2, c0=1.0, c1=1.0, relu=False):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.zeros(a.shape, np.float32)
    lib().oracle_eltwise_f32(_p(a), _p(b), _p(out), C.c_size_t(a.size), op, _f(c0), _f(c1), int(relu))
    return out


def eltwise_sum_q8(a, b, sa, sb, out_dtype, relu=False):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = np.zeros(a.shape, _NP[out_dtype])
    lib().oracle_eltwise_sum_q8(_p(a), _dt(a), _p(b), _dt(b), _p(out), out_dtype, C.c_size_t(a.size),
                                _f(sa), _f(sb), int(relu))
    return out


def activation_f32(x, act, neg_slope=0.0, coef=1.0):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.float32)
    lib().oracle_activation_f32(_p(x), _p(out), C.c_size_t(x.size), act, _f(neg_slope), _f(coef))
    return out


def scale_f32(x, outer, c, inner, w, b):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    lib().oracle_scale_f32(_p(x), _p(out), outer, c, inner, _p(w), _p(b))
    return out


# ---- the reference's own op-test oracles (oracle/_ref, compiled from /root/reference): pin the restatements above
def ref_pool_f32(x, window, pad, stride, ptype):
    x = np.ascontiguousarray(x, np.float32)
    n, c, h, w = x.shape
    oh, ow = pool_out_size(h, w, window[0], window[1], pad[0], pad[1], stride[0], stride[1])
    out = np.zeros((n, c, oh, ow), np.float32)
    ref_lib().ref_pooling_cpu_f32(_p(x), _p(out), n, c, h, w, oh, ow, window[0], window[1], pad[0], pad[1],
                                  stride[0], stride[1], int(ptype))
    return out


def ref_fc_f32(x, w, bias):
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    m = x.shape[0]
    k = x.size // m
    n_out = w.size // k
    out = np.zeros((m, n_out), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    ref_lib().ref_fc_cpu_f32(_p(x), _p(w), _p(b), _p(out), m, k, n_out)
    return out


def ref_softmax_f32(x, axis):
    x = np.ascontiguousarray(x, np.float32)
    n, c, h, w = x.shape
    out = np.zeros(x.shape, np.float32)
    ref_lib().ref_softmax_cpu_f32(_p(x), _p(out), n, c, h, w, axis)
    return out


def ref_eltwise_f32(a, b, op=2, c0=1.0, c1=1.0, relu=False):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.zeros(a.shape, np.float32)
    ref_lib().ref_eltwise_cpu_f32(_p(a), _p(b), _p(out), a.size, op, _f(c0), _f(c1), int(relu))
    return out


def ref_activation_f32(x, act, neg_slope=0.0, coef=1.0):
    x = np.ascontiguousarray(x, np.float32)
    n, c, h, w = x.shape
    out = np.zeros(x.shape, np.float32)
    ref_lib().ref_activation_f32(_p(x), _p(out), n, c, h, w, act, _f(neg_slope), _f(coef))
    return out


def tensor_cmp(a, b, use_ref=False):
    a = np.ascontiguousarray(a, np.float32).ravel()
    b = np.ascontiguousarray(b, np.float32).ravel()
    mr, md = C.c_double(), C.c_double()
    if use_ref:
        ref_lib().ref_tensor_cmp_host(_p(a), _p(b), a.size, C.byref(mr), C.byref(md))
    else:
        lib().oracle_tensor_cmp(_p(a), _p(b), C.c_size_t(a.size), C.byref(mr), C.byref(md))
    return mr.value, md.value
