"""Model-level CPU oracle: walks an Anakin graph (dict form of anakin_b200.anakin_bin) with the
op-level oracle (oracle/oracle.c) and x86 Saber semantics.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  It is deliberately independent of
the C++ framework: it does its own pattern grouping, so that a disagreement between the
two shows up as a parity failure instead of cancelling out.

Semantics restated (reference file:line):
  * grouping: Convolution [+BatchNorm] [+Scale] [+ReLU], Eltwise [+ReLU]  --
    framework/graph/llvm/fusion/fusion_op_register.cpp:45-175; conv + eltwise(+relu) ->
    ConvEltwise with the *other* eltwise input computed first --
    framework/graph/llvm/optimizer/conv_elewise_fusion_scheduler.cpp:31-136.
  * BN/Scale fold -- framework/utils/parameter_fusion.cpp:86-131.
  * fp32 op order -- saber/funcs/impl/x86/saber_im2col_conv.cpp:161-214.
  * INT8: weights per-output-channel max/127 with truncating cast (x86_utils.h:293-323);
    first-layer input roundf+clamp (x86_utils.h:318-347); conv epilogue and scale tables
    of the VNNI JIT path (kernel/jit_avx512_core_x8s8s32x_conv{,_kernel}.cpp); relu outputs
    are u8 with scale*127/255, bare conv outputs s8 (docs/Manual/int8_design_ch.md);
    int8 pooling passes its input scale through (saber/funcs/impl/x86/saber_pooling.cpp:583-584);
    edge dtype is int8 iff both endpoints are int8 nodes
    (framework/core/net/calibrator_parse.cpp:88-127).
"""
import numpy as np

from . import pyoracle as O

DT_FLOAT, DT_INT8, DT_UINT8 = O.DT_FLOAT, O.DT_INT8, O.DT_UINT8


def _attr_tensor(v):
    if isinstance(v, dict):
        return np.asarray(v["tensor"], np.float32)
    return np.asarray(v, np.float32)


class Group:
    """One fused execution unit."""

    def __init__(self, kind, head):
        self.kind = kind          # conv | pool | dense | softmax | eltwise | relu | input | output | split | bn | scale
        self.head = head          # node dict of the first op
        self.nodes = [head]
        self.relu = False
        self.relu_alpha = 0.0
        self.bn = None
        self.scale = None
        self.elt = None           # eltwise node fused into a conv (ConvEltwise)
        self.residual = None      # producer name of the residual input
        self.out_name = head["name"]  # name of the LAST original node (its output edge)
        self.inputs = list(head["ins"])


def plan(graph, fuse_conv_eltwise=True):
    """Group nodes the way Graph::Optimize does; returns groups in execution order."""
    nodes = {n["name"]: n for n in graph["nodes"]}
    consumed = set()
    groups = []
    by_out = {}

    def sole_consumer(n):
        return nodes[n["outs"][0]] if len(n["outs"]) == 1 else None

    for n in graph["nodes"]:
        if n["name"] in consumed:
            continue
        op = n["op"]
        if op == "Convolution":
            g = Group("conv", n)
            cur = n
            nxt = sole_consumer(cur)
            if nxt is not None and nxt["op"] == "BatchNorm":
                g.bn = nxt; g.nodes.append(nxt); consumed.add(nxt["name"]); cur = nxt; nxt = sole_consumer(cur)
            if nxt is not None and nxt["op"] == "Scale":
                g.scale = nxt; g.nodes.append(nxt); consumed.add(nxt["name"]); cur = nxt; nxt = sole_consumer(cur)
            if nxt is not None and nxt["op"] == "ReLU":
                g.relu = True; g.relu_alpha = float(nxt["attrs"].get("alpha", 0.0))
                g.nodes.append(nxt); consumed.add(nxt["name"]); cur = nxt
            g.out_name = cur["name"]
        elif op == "Eltwise":
            g = Group("eltwise", n)
            nxt = sole_consumer(n)
            if nxt is not None and nxt["op"] == "ReLU":
                g.relu = True; g.relu_alpha = float(nxt["attrs"].get("alpha", 0.0))
                g.nodes.append(nxt); consumed.add(nxt["name"]); g.out_name = nxt["name"]
        elif op == "Dense":
            g = Group("dense", n)
            nxt = sole_consumer(n)
            if nxt is not None and nxt["op"] == "ReLU":
                g.relu = True; g.relu_alpha = float(nxt["attrs"].get("alpha", 0.0))
                g.nodes.append(nxt); consumed.add(nxt["name"]); g.out_name = nxt["name"]
        else:
            kind = {"Pooling": "pool", "Softmax": "softmax", "ReLU": "relu", "Input": "input",
                    "Output": "output", "Split": "split", "BatchNorm": "bn", "Scale": "scale",
                    "Flatten": "flatten"}.get(op)
            if kind is None:
                raise NotImplementedError("oracle walker: op %s" % op)
            g = Group(kind, n)
        groups.append(g)
        for m in g.nodes:
            by_out[m["name"]] = g

    if fuse_conv_eltwise:
        # ConvEltwise: fold an Eltwise(Add, coeff 1,1)[+ReLU] group into the conv group that
        # produces its later input, provided that conv has no activation and feeds only the eltwise.
        order = {id(g): i for i, g in enumerate(groups)}
        fused = []
        for g in groups:
            if g.kind != "eltwise" or g.head["attrs"].get("type") != "Add":
                continue
            prods = [by_out[b] for b in g.inputs]
            cands = [p for p in prods if p.kind == "conv" and not p.relu and p.elt is None and
                     len(nodes[p.out_name]["outs"]) == 1]
            if not cands:
                continue
            conv_g = max(cands, key=lambda p: order[id(p)])
            others = [b for b, p in zip(g.inputs, prods) if p is not conv_g]
            if len(others) != 1:
                continue
            conv_g.elt = g.head
            conv_g.residual = others[0]
            conv_g.relu = g.relu
            conv_g.relu_alpha = g.relu_alpha
            conv_g.out_name = g.out_name
            conv_g.nodes += g.nodes
            for m in g.nodes:
                by_out[m["name"]] = conv_g
            fused.append(g)
        groups = [g for g in groups if g not in fused]
    return groups


def _folded_conv_weights(g):
    n = g.head
    a = n["attrs"]
    w = _attr_tensor(a["weight_1"]).copy()
    k = w.shape[0]
    bias = _attr_tensor(a["weight_2"]).reshape(-1).copy() if a.get("bias_term") else np.zeros(k, np.float32)
    if g.bn is not None or g.scale is not None:
        if g.bn is not None:
            mean = _attr_tensor(g.bn["attrs"]["weight_1"]).reshape(-1)
            var = _attr_tensor(g.bn["attrs"]["weight_2"]).reshape(-1)
            factor = float(_attr_tensor(g.bn["attrs"]["weight_3"]).reshape(-1)[0])
            eps = float(g.bn["attrs"]["epsilon"])
        else:
            mean, var, factor, eps = np.zeros(k, np.float32), np.ones(k, np.float32), 1.0, 0.0
        if g.scale is not None:
            gamma = _attr_tensor(g.scale["attrs"]["weight_1"]).reshape(-1)
            beta = _attr_tensor(g.scale["attrs"]["weight_2"]).reshape(-1) if g.scale["attrs"].get("bias_term") else None
        else:
            gamma, beta = np.ones(k, np.float32), None
        w, bias = O.fold_bn_scale(w, bias, factor, eps, mean, var, gamma, beta)
    return w, bias


def _conv_kw(a):
    return dict(stride=tuple(a["strides"]), pad=tuple(a["padding"]), dil=tuple(a["dilation_rate"]))


def run_fp32(graph, x_nchw, collect_absmax=False, return_values=False, fast=False, weight_cache=None):
    """FP32 forward (NHWC internally). Returns {output_name: ndarray}, and with collect_absmax
    also {node_name: max|x|} of every node's output (for max-abs calibration,
    CalibrationAlgoType::MAXABS, saber/saber_types.h:357-360).
    fast / weight_cache (bench.py's CPU arm): AVX-512 convolutions and inner products (equal up to float re-association), the
    BN/Scale fold and the weight packs kept between calls (init-time work in the reference: Net::init)."""
    groups = plan(graph)
    vals = {}
    absmax = {}

    def put(g, v):
        for m in g.nodes:
            vals[m["name"]] = v
        if collect_absmax:
            am = float(np.abs(v).max())
            for m in g.nodes:
                absmax[m["name"]] = am

    outputs = {}
    for g in groups:
        a = g.head["attrs"]
        if g.kind == "input":
            x = np.ascontiguousarray(np.transpose(np.asarray(x_nchw, np.float32), (0, 2, 3, 1)))
            put(g, x)
        elif g.kind == "split":
            put(g, vals[g.inputs[0]])
        elif g.kind == "conv":
            cached = weight_cache.get(g.head["name"]) if weight_cache is not None else None
            if cached is None:
                cached = _folded_conv_weights(g)
                if weight_cache is not None:
                    weight_cache[g.head["name"]] = cached
            w, bias = cached
            res = vals[g.residual] if g.elt is not None else None
            src = vals[[b for b in g.inputs][0]]
            y = O.conv_f32_nhwc(src, w, bias, residual=res, group=int(a["group"]), relu=g.relu,
                                neg_slope=g.relu_alpha, beta=1.0, fast=fast, **_conv_kw(a))
            put(g, y)
        elif g.kind == "eltwise":
            op = {"Add": 2, "Mul": 1, "Prod": 1, "Max": 3}[a["type"]]
            c = a.get("coeff") or [1.0, 1.0]
            y = O.eltwise_f32(vals[g.inputs[0]], vals[g.inputs[1]], op, c[0], c[1] if len(c) > 1 else 1.0, g.relu)
            put(g, y)
        elif g.kind == "relu":
            put(g, O.activation_f32(vals[g.inputs[0]], 2, float(a.get("alpha", 0.0))))
        elif g.kind == "pool":
            ptype = {"MAX": 1, "AVG": 2, "AVGEXC": 3}[a["method"]]
            y = O.pool_f32(vals[g.inputs[0]], tuple(a["pool_size"]), tuple(a["padding"]), tuple(a["strides"]),
                           ptype, nhwc=True, global_pooling=bool(a["global_pooling"]),
                           floor_as_conv=bool(a.get("cmp_out_shape_floor_as_conv", False)))
            put(g, y)
        elif g.kind == "dense":
            x = vals[g.inputs[0]]
            if x.ndim == 4:  # Dense flattens in NCHW order (saber_fc.cu:27-35)
                x = np.transpose(x, (0, 3, 1, 2))
            x = np.ascontiguousarray(x).reshape(x.shape[0], -1)
            if fast and weight_cache is not None:
                # the inner product as a 1x1 convolution over [m, 1, 1, K] through the packed AVX-512 path
                cached = weight_cache.get(g.head["name"])
                if cached is None:
                    w = np.ascontiguousarray(_attr_tensor(a["weight_1"]).reshape(int(a["out_dim"]), -1, 1, 1), np.float32)
                    b = _attr_tensor(a["weight_2"]).reshape(-1) if a.get("bias_term") else None
                    cached = weight_cache[g.head["name"]] = (w, b)
                w, b = cached
                y = O.conv_f32_nhwc(x.reshape(x.shape[0], 1, 1, -1), w, b, relu=bool(g.relu), neg_slope=g.relu_alpha,
                                    fast=True).reshape(x.shape[0], -1)
            else:
                w = _attr_tensor(a["weight_1"]).reshape(int(a["out_dim"]), -1)
                b = _attr_tensor(a["weight_2"]).reshape(-1) if a.get("bias_term") else None
                y = O.fc_f32(x, w, b)
                if g.relu:
                    y = O.activation_f32(y, 2, g.relu_alpha)
            put(g, y.reshape(y.shape[0], 1, 1, -1))
        elif g.kind == "softmax":
            x = vals[g.inputs[0]]
            n = x.shape[0]
            y = O.softmax_f32(x.reshape(n, -1), n, x.size // n, 1)
            put(g, y.reshape(x.shape))
        elif g.kind == "output":
            v = vals[g.inputs[0]]
            outputs[g.head["name"]] = v.reshape(v.shape[0], -1) if v.shape[1] == 1 and v.shape[2] == 1 else \
                np.transpose(v, (0, 3, 1, 2))
        else:
            raise NotImplementedError(g.kind)
    if return_values:          # every node's fp32 output (NHWC), for goldens of intermediate edges (logits)
        return outputs, vals
    if collect_absmax:
        return outputs, absmax
    return outputs


def calibrate(graph, images_nchw):
    """Max-abs calibration over a batch of images: scale = max|x| / 127 per node output."""
    _, absmax = run_fp32(graph, images_nchw, collect_absmax=True)
    return {k: (v / 127.0 if v > 0 else 1.0) for k, v in absmax.items()}


def run_int8(graph, x_nchw, edge_scales, return_intermediate=False, fast=False, weight_cache=None):
    """INT8 forward with x86 Saber semantics. edge_scales: {node_name: output scale}."""
    groups = plan(graph)
    int8_ops = {"Convolution", "BatchNorm", "Scale", "ReLU", "Pooling", "Eltwise", "Dense", "Split", "Input"}
    nodes = {n["name"]: n for n in graph["nodes"]}

    def node_is_int8(name):
        n = nodes[name]
        bt = n.get("bit_type")
        return (bt == "INT8") if bt is not None else (n["op"] in int8_ops)

    vals = {}  # node name -> (array NHWC, dtype code, scale)
    outputs = {}
    trace = {}

    def put(g, arr, dt, scale):
        for m in g.nodes:
            vals[m["name"]] = (arr, dt, scale)
        trace[g.out_name] = (arr, dt, scale)
        trace[g.head["name"]] = (arr, dt, scale)

    def consumers_int8(g):
        outs = nodes[g.out_name]["outs"]
        return all(node_is_int8(o) for o in outs) and len(outs) > 0

    def as_float(v):
        arr, dt, sc = v
        if dt == DT_FLOAT:
            return arr
        f = arr.astype(np.float32) * np.float32(sc)
        if dt == DT_UINT8:
            f = f * np.float32(127.0 / 255.0)
        return f

    for g in groups:
        a = g.head["attrs"]
        if g.kind == "input":
            x = np.ascontiguousarray(np.transpose(np.asarray(x_nchw, np.float32), (0, 2, 3, 1)))
            put(g, x, DT_FLOAT, edge_scales[g.out_name])
        elif g.kind == "split":
            arr, dt, sc = vals[g.inputs[0]]
            put(g, arr, dt, sc)
        elif g.kind in ("conv", "dense"):
            src, sdt, s_in = vals[g.inputs[0]]
            # weight_cache (a dict the caller keeps between calls): BN/Scale fold and weight quantisation are init-time
            # work in the reference (Net::init, trans_weights) -- the CPU-baseline arm times inference, not them
            ckey = (g.head["name"], tuple(src.shape[1:]))
            cached = weight_cache.get(ckey) if weight_cache is not None else None
            if g.kind == "conv":
                w, bias = (None, cached[2]) if cached else _folded_conv_weights(g)
                kw = _conv_kw(a)
                kw["group"] = int(a["group"])
            else:
                if src.ndim == 4 and src.shape[1] * src.shape[2] > 1:
                    n_, h_, w_, c_ = src.shape
                if not cached:
                    w = _attr_tensor(a["weight_1"]).reshape(int(a["out_dim"]), -1)
                    if src.ndim == 4 and src.shape[1] * src.shape[2] > 1:
                        # NCHW-flatten order -> permute weight columns to NHWC order
                        w = w.reshape(-1, c_, h_, w_).transpose(0, 2, 3, 1).reshape(w.shape[0], -1)
                    w = w.reshape(w.shape[0], -1, 1, 1)
                    bias = _attr_tensor(a["weight_2"]).reshape(-1) if a.get("bias_term") else np.zeros(w.shape[0], np.float32)
                else:
                    bias = cached[2]
                if src.ndim == 4 and src.shape[1] * src.shape[2] > 1:
                    src = src.reshape(n_, 1, 1, -1)
                kw = dict(stride=(1, 1), pad=(0, 0), dil=(1, 1))
            if sdt == DT_FLOAT:  # quantise the fp32 input with the input edge's scale
                src = O.quant_fp32_s8(src, s_in)
                sdt = DT_INT8
            if cached:
                wq, w_scale = cached[0], cached[1]
            else:
                wq, w_scale = O.quant_weights_per_oc(w)
                if weight_cache is not None:
                    weight_cache[ckey] = (wq, w_scale, bias)
            s_out = edge_scales[g.out_name]
            if consumers_int8(g):
                out_dt = DT_UINT8 if g.relu else DT_INT8
            else:
                out_dt = DT_FLOAT
            res = None
            res_dt, res_scale = DT_INT8, 1.0
            if g.kind == "conv" and g.elt is not None:
                res, res_dt, res_scale = vals[g.residual]
                if res_dt == DT_FLOAT:
                    raise NotImplementedError("fp32 residual into int8 ConvEltwise")
            scale, bias_f, sum_scale = O.int8_conv_scales(w_scale, bias, s_in, sdt, s_out, out_dt,
                                                          res_scale, res_dt)
            # fast: the AVX-512 VNNI implementation of the same arithmetic (bit-identical; bench.py's CPU arm)
            y = O.conv_s8_nhwc_x86(src, wq, bias_f, scale, residual=res, sum_scale=sum_scale,
                                   out_dtype=out_dt, relu=g.relu, fast=fast, **kw)
            put(g, y, out_dt, s_out)
        elif g.kind == "pool":
            src, sdt, s_in = vals[g.inputs[0]]
            ptype = {"MAX": 1, "AVG": 2, "AVGEXC": 3}[a["method"]]
            args = (tuple(a["pool_size"]), tuple(a["padding"]), tuple(a["strides"]), ptype)
            kwp = dict(global_pooling=bool(a["global_pooling"]),
                       floor_as_conv=bool(a.get("cmp_out_shape_floor_as_conv", False)))
            if sdt == DT_FLOAT:
                put(g, O.pool_f32(src, *args, nhwc=True, **kwp), DT_FLOAT, s_in)
            else:
                put(g, O.pool_s8_nhwc(src, *args, fast=fast, **kwp), sdt, s_in)  # scale passes through
        elif g.kind == "eltwise":
            (x0, d0, s0), (x1, d1, s1) = vals[g.inputs[0]], vals[g.inputs[1]]
            s_out = edge_scales[g.out_name]
            out_dt = DT_UINT8 if g.relu else DT_INT8
            u = np.float32(127.0 / 255.0)
            f0 = np.float32(s0) * (u if d0 == DT_UINT8 else np.float32(1))
            f1 = np.float32(s1) * (u if d1 == DT_UINT8 else np.float32(1))
            fo = np.float32(s_out) * (u if out_dt == DT_UINT8 else np.float32(1))
            y = O.eltwise_sum_q8(x0, x1, float(f0 / fo), float(f1 / fo), out_dt, relu=g.relu)
            put(g, y, out_dt, s_out)
        elif g.kind == "softmax":
            x = as_float(vals[g.inputs[0]])
            n = x.shape[0]
            y = O.softmax_f32(x.reshape(n, -1), n, x.size // n, 1)
            put(g, y.reshape(x.shape), DT_FLOAT, 1.0)
        elif g.kind == "relu":
            arr, dt, sc = vals[g.inputs[0]]
            put(g, O.activation_f32(as_float((arr, dt, sc)), 2, float(a.get("alpha", 0.0))), DT_FLOAT, sc)
        elif g.kind == "output":
            v = as_float(vals[g.inputs[0]])
            outputs[g.head["name"]] = v.reshape(v.shape[0], -1) if v.shape[1] == 1 and v.shape[2] == 1 else \
                np.transpose(v, (0, 3, 1, 2))
        else:
            raise NotImplementedError(g.kind)
    if return_intermediate:
        return outputs, trace
    return outputs
