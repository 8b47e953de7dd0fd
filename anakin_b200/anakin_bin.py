"""Reader / writer for Anakin's model file (*.anakin.bin) without protobuf.

The file is ONE serialized proto3 `GraphProto` (reference
framework/model_parser/proto/{graph,node,tensor,operator}.proto; read by
framework/model_parser/parser/parser.cpp:67-237 and model_io.cpp:9-260).  This module
hand-codes the proto3 wire format for exactly those four schemas.  It is used to
synthesise the benchmark models (modelzoo.py) and by the tests to round-trip what the C++
parser (csrc/framework/model_parser.cpp) reads and writes.

In-memory form (plain dicts, no classes):
  graph = {"name": str, "nodes": [node...], "ins": [str], "outs": [str],
           "edges_in":  {node_name: [(bottom_name, scale_or_None), ...]},
           "edges_out": {node_name: [(top_name,    scale_or_None), ...]},
           "edges_info": {edge_name: {"shared": bool, "share_from": str}},
           "version": (major, minor, patch, version), "is_optimized": bool}
  node  = {"name": str, "op": str, "ins": [str], "outs": [str], "attrs": {key: value},
           "bit_type": None | "FLOAT" | "INT8", "lane": int, "need_wait": bool}
  attr values: str | bool | int | float | list[int|float|bool|str] | np.ndarray(float32, 4-D)
               | {"tensor": ndarray, "scale": [float]} (tensor with int8 scale; an int8 ndarray is stored as the
                 INT8 payload CacheDate.c, model_io.cpp:204-216)
               | {"share_from": node_name} (TensorProto.shared: the owner node's tensor of the same key,
                 model_io.cpp:147-151)
"""
import struct

import numpy as np

# DateTypeProto (tensor.proto)
STR, INT8, INT32, FLOAT16, FLOAT, DOUBLE, BOOLEN, CACHE_LIST, TENSOR = 0, 2, 4, 8, 13, 14, 20, 30, 31
_BIT = {None: 0, "INT8": INT8, "FLOAT": FLOAT}
_BIT_INV = {0: None, INT8: "INT8", FLOAT: "FLOAT"}
LP_NCHW, LP_NHWC = 8, 9


# ----------------------------------------------------------------------------- wire encode
def _varint(n):
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wt):
    return _varint((field << 3) | wt)


def _f_varint(field, v):
    return _tag(field, 0) + _varint(int(v))


def _f_bytes(field, b):
    if isinstance(b, str):
        b = b.encode("utf-8")
    return _tag(field, 2) + _varint(len(b)) + b


def _f_float(field, v):
    return _tag(field, 5) + struct.pack("<f", v)


def _enc_cache(values, dtype_code):
    """CacheDate{s=1,i=2,f=3,b=4,l=5,type=6,size=7,c=8}"""
    out = bytearray()
    n = len(values)
    if dtype_code == FLOAT:
        arr = np.ascontiguousarray(values, dtype="<f4")
        if n:
            out += _tag(3, 2) + _varint(arr.nbytes)
            out += arr.tobytes()
    elif dtype_code == INT32:
        if n:
            body = b"".join(_varint(int(v)) for v in values)
            out += _tag(2, 2) + _varint(len(body)) + body
    elif dtype_code == BOOLEN:
        if n:
            body = bytes(1 if v else 0 for v in values)
            out += _tag(4, 2) + _varint(len(body)) + body
    elif dtype_code == STR:
        for v in values:
            out += _f_bytes(1, v)
    else:
        raise ValueError("unsupported list type %r" % dtype_code)
    if dtype_code:
        out += _f_varint(6, dtype_code)
    if n:
        out += _f_varint(7, n)
    return bytes(out)


def _enc_shape(dims):
    body = b"".join(_varint(int(d)) for d in dims)
    dim = _tag(1, 2) + _varint(len(body)) + body + _f_varint(2, len(dims))
    return _f_bytes(3, dim)


def _enc_tensor(arr, scale=None, name=None):
    """TensorProto{name=1,shared=2,share_from=3,shape=8,valid_shape=9,data=10,scale=11}"""
    q8 = isinstance(arr, np.ndarray) and arr.dtype == np.int8
    arr = np.ascontiguousarray(arr, dtype=np.int8 if q8 else np.float32)
    shape = list(arr.shape)
    while len(shape) < 4:
        shape.insert(0, 1)
    out = bytearray()
    if name:
        out += _f_bytes(1, name)
    out += _f_bytes(8, _enc_shape(shape))
    out += _f_bytes(9, _enc_shape(shape))
    if q8:
        cache = (_f_bytes(8, arr.tobytes()) if arr.size else b"") + _f_varint(6, INT8) + \
                (_f_varint(7, arr.size) if arr.size else b"")
        out += _f_bytes(10, cache)
    else:
        out += _f_bytes(10, _enc_cache(arr.ravel(), FLOAT))
    if scale is not None and len(scale):
        out += _f_bytes(11, _enc_cache(list(scale), FLOAT))
    return bytes(out)


def _enc_value(v):
    """valueType{s=1,i=2,f=3,b=4,cache_list=8,tensor=10,type=14}"""
    if isinstance(v, dict) and "share_from" in v:
        return _f_bytes(10, _f_varint(2, 1) + _f_bytes(3, v["share_from"])) + _f_varint(14, TENSOR)
    if isinstance(v, dict) and "tensor" in v:
        return _f_bytes(10, _enc_tensor(v["tensor"], v.get("scale"))) + _f_varint(14, TENSOR)
    if isinstance(v, np.ndarray):
        return _f_bytes(10, _enc_tensor(v)) + _f_varint(14, TENSOR)
    if isinstance(v, bool):
        return (_f_varint(4, 1) if v else b"") + _f_varint(14, BOOLEN)
    if isinstance(v, (int, np.integer)):
        return (_f_varint(2, int(v)) if v else b"") + _f_varint(14, INT32)
    if isinstance(v, (float, np.floating)):
        return (_f_float(3, float(v)) if float(v) != 0.0 else b"") + _f_varint(14, FLOAT)
    if isinstance(v, (str, bytes)):
        return _f_bytes(1, v)  # type STR == 0 is the proto3 default and is not emitted
    if isinstance(v, (list, tuple)):
        if len(v) and isinstance(v[0], bool):
            code = BOOLEN
        elif len(v) and isinstance(v[0], (int, np.integer)):
            code = INT32
        elif len(v) and isinstance(v[0], (float, np.floating)):
            code = FLOAT
        elif len(v) and isinstance(v[0], (str, bytes)):
            code = STR
        else:
            code = INT32
        return _f_bytes(8, _enc_cache(list(v), code)) + _f_varint(14, CACHE_LIST)
    raise TypeError("cannot encode attr value of type %s" % type(v))


def _enc_node(node):
    out = bytearray()
    out += _f_bytes(1, node["name"])
    for s in node.get("ins", []):
        out += _f_bytes(2, s)
    for s in node.get("outs", []):
        out += _f_bytes(3, s)
    # map entries in key order: what protobuf's deterministic serialisation (and the C++ writer's std::map) emit
    for k, v in sorted(node.get("attrs", {}).items(), key=lambda kv: kv[0].encode("utf-8")):
        entry = _f_bytes(1, k) + _f_bytes(2, _enc_value(v))
        out += _f_bytes(10, entry)
    if node.get("lane"):
        out += _f_varint(11, node["lane"])
    if node.get("need_wait"):
        out += _f_varint(12, 1)
    n_in, n_out = len(node.get("ins", [])), len(node.get("outs", []))   # proto3: zero scalars are not emitted
    op = _f_bytes(1, node["op"]) + (_f_varint(3, n_in) if n_in else b"") + (_f_varint(4, n_out) if n_out else b"")
    out += _f_bytes(15, op)
    bt = _BIT[node.get("bit_type")]
    if bt:
        out += _f_varint(16, bt)
    return bytes(out)


def _enc_list(targets):
    """List{val=1, target=2}; TargetProto{node=1, scale=2(packed float), layout=3}"""
    out = bytearray()
    for name, scale in targets:
        if scale is None:
            out += _f_bytes(1, name)
        else:
            sc = np.atleast_1d(np.asarray(scale, dtype="<f4"))
            t = _f_bytes(1, name) + _tag(2, 2) + _varint(sc.nbytes) + sc.tobytes() + _f_varint(3, LP_NCHW)
            out += _f_bytes(2, t)
    return bytes(out)


def dumps(graph):
    out = bytearray()
    out += _f_bytes(1, graph.get("name", "graph"))
    for node in graph["nodes"]:
        out += _f_bytes(2, _enc_node(node))
    for field, key in ((3, "edges_in"), (4, "edges_out")):
        for name, targets in sorted(graph.get(key, {}).items(), key=lambda kv: kv[0].encode("utf-8")):
            out += _f_bytes(field, _f_bytes(1, name) + _f_bytes(2, _enc_list(targets)))
    for ename, info in sorted(graph.get("edges_info", {}).items(), key=lambda kv: kv[0].encode("utf-8")):
        t = _f_bytes(1, ename)
        if info.get("shared"):
            t += _f_varint(2, 1) + _f_bytes(3, info.get("share_from", ""))
        out += _f_bytes(5, _f_bytes(1, ename) + _f_bytes(2, t))
    for s in graph.get("ins", []):
        out += _f_bytes(6, s)
    for s in graph.get("outs", []):
        out += _f_bytes(7, s)
    ver = graph.get("version", (2, 0, 0, 200))
    out += _f_bytes(10, _f_varint(1, ver[0]) + (_f_varint(2, ver[1]) if ver[1] else b"") +
                    (_f_varint(3, ver[2]) if ver[2] else b"") + _f_varint(4, ver[3]))
    out += _f_bytes(11, _f_varint(10, 1) if graph.get("is_optimized") else b"")
    return bytes(out)


def save(graph, path):
    with open(path, "wb") as f:
        f.write(dumps(graph))


# ----------------------------------------------------------------------------- wire decode
def _read_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field, wiretype, value) over a message; value is int (wt 0/1/5 raw) or memoryview (wt 2)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield field, wt, v


def _s(v):
    return bytes(v).decode("utf-8")


def _packed_varints(v):
    out, pos = [], 0
    while pos < len(v):
        x, pos = _read_varint(v, pos)
        if x >= 1 << 63:
            x -= 1 << 64
        elif x >= 1 << 31 and x < 1 << 32:
            x -= 1 << 32
        out.append(x)
    return out


def _dec_cache(buf):
    d = {"s": [], "i": [], "f": None, "b": [], "c": b"", "type": 0, "size": 0}
    fchunks = []
    for field, wt, v in _fields(buf):
        if field == 1:
            d["s"].append(_s(v))
        elif field == 2:
            d["i"].extend(_packed_varints(v) if wt == 2 else [v])
        elif field == 3:
            fchunks.append(np.frombuffer(v, dtype="<f4") if wt == 2 else np.frombuffer(v, dtype="<f4"))
        elif field == 4:
            d["b"].extend([bool(x) for x in bytes(v)] if wt == 2 else [bool(v)])
        elif field == 8:
            d["c"] = bytes(v)
        elif field == 6:
            d["type"] = v
        elif field == 7:
            d["size"] = v
    d["f"] = np.concatenate(fchunks) if fchunks else np.zeros(0, np.float32)
    return d


def _dec_shape(buf):
    dims = []
    for field, wt, v in _fields(buf):
        if field == 3:
            for f2, wt2, v2 in _fields(v):
                if f2 == 1:
                    dims.extend(_packed_varints(v2) if wt2 == 2 else [v2])
    return dims


def _dec_tensor(buf):
    t = {"name": "", "shared": False, "share_from": "", "shape": [], "valid_shape": [], "data": None, "scale": []}
    for field, wt, v in _fields(buf):
        if field == 1:
            t["name"] = _s(v)
        elif field == 2:
            t["shared"] = bool(v)
        elif field == 3:
            t["share_from"] = _s(v)
        elif field == 8:
            t["shape"] = _dec_shape(v)
        elif field == 9:
            t["valid_shape"] = _dec_shape(v)
        elif field == 10:
            c = _dec_cache(v)
            t["data"] = np.frombuffer(c["c"], dtype=np.int8) if c["type"] == INT8 else c["f"]
        elif field == 11:
            t["scale"] = [float(x) for x in _dec_cache(v)["f"]]
    return t


def _dec_value(buf):
    raw = {}
    typ = STR
    for field, wt, v in _fields(buf):
        if field == 14:
            typ = v
        else:
            raw[field] = v
    if typ == STR:
        return _s(raw.get(1, b""))
    if typ == INT32:
        x = raw.get(2, 0)
        return x - (1 << 64) if x >= 1 << 63 else x
    if typ in (FLOAT, DOUBLE):
        return struct.unpack("<f", raw[3])[0] if 3 in raw else 0.0
    if typ == BOOLEN:
        return bool(raw.get(4, 0))
    if typ == CACHE_LIST:
        c = _dec_cache(raw.get(8, b""))
        if c["type"] == FLOAT:
            return [float(x) for x in c["f"]]
        if c["type"] == INT32:
            return list(c["i"])
        if c["type"] == BOOLEN:
            return list(c["b"])
        return list(c["s"])
    if typ == TENSOR:
        t = _dec_tensor(raw.get(10, b""))
        if t["shared"]:
            return {"share_from": t["share_from"]}
        arr = None
        if t["data"] is not None:
            arr = np.array(t["data"], dtype=np.int8 if t["data"].dtype == np.int8 else np.float32).reshape(t["shape"])
        if t["scale"]:
            return {"tensor": arr, "scale": t["scale"]}
        return arr
    raise ValueError("unsupported attr type %d" % typ)


def _dec_node(buf):
    node = {"name": "", "op": "", "ins": [], "outs": [], "attrs": {}, "bit_type": None, "lane": 0,
            "need_wait": False}
    for field, wt, v in _fields(buf):
        if field == 1:
            node["name"] = _s(v)
        elif field == 2:
            node["ins"].append(_s(v))
        elif field == 3:
            node["outs"].append(_s(v))
        elif field == 10:
            k, val = None, None
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    k = _s(v2)
                elif f2 == 2:
                    val = _dec_value(v2)
            node["attrs"][k] = val
        elif field == 11:
            node["lane"] = v
        elif field == 12:
            node["need_wait"] = bool(v)
        elif field == 15:
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    node["op"] = _s(v2)
        elif field == 16:
            node["bit_type"] = _BIT_INV.get(v)
    return node


def _dec_list(buf):
    out = []
    for field, wt, v in _fields(buf):
        if field == 1:
            out.append((_s(v), None))
        elif field == 2:
            name, scale = "", []
            for f2, wt2, v2 in _fields(v):
                if f2 == 1:
                    name = _s(v2)
                elif f2 == 2:
                    scale.extend(np.frombuffer(v2, dtype="<f4").tolist() if wt2 == 2
                                 else [struct.unpack("<f", v2)[0]])
            out.append((name, scale))
    return out


def loads(data):
    buf = memoryview(data)
    g = {"name": "", "nodes": [], "ins": [], "outs": [], "edges_in": {}, "edges_out": {}, "edges_info": {},
         "version": (0, 0, 0, 0), "is_optimized": False}
    for field, wt, v in _fields(buf):
        if field == 1:
            g["name"] = _s(v)
        elif field == 2:
            g["nodes"].append(_dec_node(v))
        elif field in (3, 4):
            k, lst = None, []
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    k = _s(v2)
                elif f2 == 2:
                    lst = _dec_list(v2)
            g["edges_in" if field == 3 else "edges_out"][k] = lst
        elif field == 5:
            k, t = None, None
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    k = _s(v2)
                elif f2 == 2:
                    t = _dec_tensor(v2)
            g["edges_info"][k] = {"shared": t["shared"], "share_from": t["share_from"]} if t else {}
        elif field == 6:
            g["ins"].append(_s(v))
        elif field == 7:
            g["outs"].append(_s(v))
        elif field == 10:
            ver = [0, 0, 0, 0]
            for f2, _, v2 in _fields(v):
                if 1 <= f2 <= 4:
                    ver[f2 - 1] = v2
            g["version"] = tuple(ver)
        elif field == 11:
            for f2, _, v2 in _fields(v):
                if f2 == 10:
                    g["is_optimized"] = bool(v2)
    return g


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())
