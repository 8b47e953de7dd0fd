"""Python front end over the framework C ABI (include/anakin_b200.h, libanakin_b200.so).

Mirrors the reference's user API names (Graph.load / ResetBatchSize / Optimize / save,
Net.init / prediction / get_in / get_out -- examples/cuda/example_nv_cnn_net.cpp:21-71) for
tests and bench.py.  All compute happens in the native libraries; there is no Python or
CPU fallback path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ANAKIN_B200_LIBDIR: alternative directory holding both .so files (A/B experiments between builds)
_LIBDIR = os.environ.get("ANAKIN_B200_LIBDIR") or os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(_LIBDIR, "libanakin_b200.so")

FP32, FP16, INT8 = 0, -1, -2
PRECISIONS = {"fp32": FP32, "fp16": FP16, "int8": INT8}
_NP_OF_DTYPE = {0: np.float16, 1: np.float32, 3: np.int8, 7: np.uint8}

_vp, _i, _sz, _cp = C.c_void_p, C.c_int, C.c_size_t, C.c_char_p
SYMBOLS = {
    "anakin_last_error": (_cp, []),
    "anakin_graph_load": (_i, [_cp, C.POINTER(_vp)]),
    "anakin_graph_load_buffer": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "anakin_graph_reset_batch_size": (_i, [_vp, _cp, _i]),
    "anakin_graph_reshape": (_i, [_vp, _cp, C.POINTER(_i)]),
    "anakin_graph_optimize": (_i, [_vp, _i]),
    "anakin_graph_save": (_i, [_vp, _cp]),
    "anakin_graph_describe": (_sz, [_vp, _vp, _sz]),
    "anakin_graph_destroy": (None, [_vp]),
    "anakin_net_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "anakin_net_num_inputs": (_i, [_vp]),
    "anakin_net_num_outputs": (_i, [_vp]),
    "anakin_net_input_name": (_cp, [_vp, _i]),
    "anakin_net_output_name": (_cp, [_vp, _i]),
    "anakin_net_tensor_info": (_i, [_vp, _cp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i),
                                    C.POINTER(C.c_float), C.POINTER(_sz)]),
    "anakin_net_tensor_device_ptr": (_vp, [_vp, _cp]),
    "anakin_net_set_input": (_i, [_vp, _cp, _vp, _sz]),
    "anakin_net_prediction": (_i, [_vp]),
    "anakin_net_sync": (_i, [_vp]),
    "anakin_net_read_tensor": (_i, [_vp, _cp, _vp, _sz]),
    "anakin_net_stream": (_vp, [_vp]),
    "anakin_net_launched_ops": (_i, [_vp]),
    "anakin_net_cuda_graph_active": (_i, [_vp]),
    "anakin_net_set_cuda_graph": (_i, [_vp, _i]),
    "anakin_net_exec_order": (_sz, [_vp, _vp, _sz]),
    "anakin_net_activation_bytes": (_sz, [_vp]),
    "anakin_net_activation_bytes_unshared": (_sz, [_vp]),
    "anakin_net_weight_ptrs": (_i, [_vp, C.POINTER(_vp), _i]),
    "anakin_weight_arena_stats": (_sz, [C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "anakin_weight_arena_set_receive": (None, [C.c_int]),
    "anakin_weight_arena_flat_bytes": (_sz, [C.c_int]),
    "anakin_weight_arena_export": (C.c_int, [C.c_int, _vp, _sz]),
    "anakin_weight_arena_import": (C.c_int, [C.c_int, _vp, _sz]),
    "anakin_net_create_ex": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "anakin_net_profile_ops": (_i, [_vp, _i, _i, C.POINTER(C.c_float), _i]),
    "anakin_net_destroy": (None, [_vp]),
    "anakin_worker_create": (_i, [_cp, _i, _i, C.POINTER(_i), _i, _i, C.POINTER(_vp)]),
    "anakin_worker_sync_prediction": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "anakin_worker_wait_ready": (_i, [_vp]),
    "anakin_worker_async_prediction": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "anakin_worker_async_get_result": (_i, [_vp]),
    "anakin_worker_destroy": (None, [_vp]),
}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libanakin_b200.so is not built (%s); run `python -m anakin_b200.build`. "
                           "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AnakinError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise AnakinError("%s: %s" % (what, load().anakin_last_error().decode()))


def _text(fn, handle):
    n = fn(handle, None, 0)
    buf = C.create_string_buffer(n + 1)
    fn(handle, buf, n + 1)
    return buf.value.decode()


class Graph:
    """graph::Graph<NV, P> (load / ResetBatchSize / Reshape / Optimize / save)."""

    def __init__(self):
        self._h = _vp()
        self._lib = load()

    @staticmethod
    def from_file(path):
        g = Graph()
        _check(g._lib.anakin_graph_load(path.encode(), C.byref(g._h)), "Graph.load(%s)" % path)
        return g

    @staticmethod
    def from_bytes(data):
        g = Graph()
        buf = C.create_string_buffer(data, len(data))
        _check(g._lib.anakin_graph_load_buffer(buf, len(data), C.byref(g._h)), "Graph.load(buffer)")
        return g

    def ResetBatchSize(self, in_name, batch):
        _check(self._lib.anakin_graph_reset_batch_size(self._h, in_name.encode(), batch), "ResetBatchSize")

    def Reshape(self, in_name, shape):
        arr = (_i * 4)(*shape)
        _check(self._lib.anakin_graph_reshape(self._h, in_name.encode(), arr), "Reshape")

    def Optimize(self, with_fusion=True):
        _check(self._lib.anakin_graph_optimize(self._h, int(with_fusion)), "Optimize")

    def save(self, path):
        _check(self._lib.anakin_graph_save(self._h, path.encode()), "save")

    def describe(self):
        """[(name, op, [ins], [outs])] in execution order."""
        out = []
        for line in _text(self._lib.anakin_graph_describe, self._h).splitlines():
            name, op, ins, outs = line.split("|")
            out.append((name, op, [s for s in ins.split(",") if s], [s for s in outs.split(",") if s]))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.anakin_graph_destroy(self._h)
            self._h = _vp()


class Net:
    """Net<NV, P>: init(graph) on a device, prediction(), tensors by node name."""

    def __init__(self, graph, precision="fp32", device=-1, keep_edges=False):
        """keep_edges=True gives every edge tensor its own buffer so intermediate tensors can be read back after
        prediction() (parity tests); the default shares buffers between edges whose live ranges do not overlap."""
        self._lib = load()
        self._h = _vp()
        prec = PRECISIONS[precision] if isinstance(precision, str) else precision
        _check(self._lib.anakin_net_create_ex(graph._h, prec, device, 1 if keep_edges else 0, C.byref(self._h)),
               "Net.init")
        self.in_names = [self._lib.anakin_net_input_name(self._h, i).decode()
                         for i in range(self._lib.anakin_net_num_inputs(self._h))]
        self.out_names = [self._lib.anakin_net_output_name(self._h, i).decode()
                          for i in range(self._lib.anakin_net_num_outputs(self._h))]

    def tensor_info(self, node):
        dims = (_i * 4)()
        cs, layout, dtype = _i(), _i(), _i()
        scale, nbytes = C.c_float(), _sz()
        _check(self._lib.anakin_net_tensor_info(self._h, node.encode(), dims, C.byref(cs), C.byref(layout),
                                                C.byref(dtype), C.byref(scale), C.byref(nbytes)), "tensor_info")
        return {"dims": list(dims), "c_stored": cs.value, "layout": layout.value, "dtype": dtype.value,
                "scale": scale.value, "bytes": nbytes.value}

    def device_ptr(self, node):
        return self._lib.anakin_net_tensor_device_ptr(self._h, node.encode())

    def set_input(self, name, host_nchw):
        a = np.ascontiguousarray(host_nchw, np.float32)
        self._keep = a
        _check(self._lib.anakin_net_set_input(self._h, name.encode(), a.ctypes.data_as(_vp), a.size), "set_input")

    def set_input_ptr(self, name, host_ptr, count):
        _check(self._lib.anakin_net_set_input(self._h, name.encode(), _vp(host_ptr), count), "set_input")

    def prediction(self):
        _check(self._lib.anakin_net_prediction(self._h), "prediction")

    def sync(self):
        _check(self._lib.anakin_net_sync(self._h), "sync")

    def read_tensor(self, node):
        """Raw storage of a node's output as numpy: NHWC tensors come back [n,h,w,c_stored]."""
        info = self.tensor_info(node)
        n, c, h, w = info["dims"]
        dt = _NP_OF_DTYPE[info["dtype"]]
        shape = (n, h, w, info["c_stored"]) if info["layout"] == 9 else (n, c, h, w)
        out = np.empty(shape, dt)
        assert out.nbytes == info["bytes"], (out.nbytes, info)
        _check(self._lib.anakin_net_read_tensor(self._h, node.encode(), out.ctypes.data_as(_vp), out.nbytes),
               "read_tensor")
        return out, info

    def read_tensor_into(self, node, host_ptr, nbytes):
        _check(self._lib.anakin_net_read_tensor(self._h, node.encode(), _vp(host_ptr), nbytes), "read_tensor")

    def get_output(self, name=None):
        """fp32 output as [N, C] (or NCHW) numpy."""
        name = name or self.out_names[0]
        arr, info = self.read_tensor(name)
        n, c, h, w = info["dims"]
        if info["layout"] == 9:
            arr = arr[..., :c]
            arr = arr.reshape(n, c) if h == 1 and w == 1 else np.transpose(arr, (0, 3, 1, 2))
        else:
            arr = arr.reshape(n, c) if h == 1 and w == 1 else arr
        return np.ascontiguousarray(arr)

    @property
    def stream(self):
        return self._lib.anakin_net_stream(self._h)

    def launched_ops(self):
        return self._lib.anakin_net_launched_ops(self._h)

    def cuda_graph_active(self):
        return bool(self._lib.anakin_net_cuda_graph_active(self._h))

    def set_cuda_graph(self, enable):
        _check(self._lib.anakin_net_set_cuda_graph(self._h, int(enable)), "set_cuda_graph")

    def exec_order(self):
        return [l.split(":") for l in _text(self._lib.anakin_net_exec_order, self._h).splitlines()]

    def profile_ops(self, iters=5, reps=1):
        """[(node, op, ms)] device time per launched op (eager, CUDA-event pair per op; reps > 1 =
        that many back-to-back launches per pair, i.e. steady-state time)."""
        order = self.exec_order()
        buf = (C.c_float * len(order))()
        _check(self._lib.anakin_net_profile_ops(self._h, iters, reps, buf, len(order)), "profile_ops")
        return [(n, o, float(buf[i])) for i, (n, o) in enumerate(order)]

    def activation_bytes(self):
        return self._lib.anakin_net_activation_bytes(self._h)

    def activation_bytes_unshared(self):
        return self._lib.anakin_net_activation_bytes_unshared(self._h)

    def weight_ptrs(self):
        n = self._lib.anakin_net_weight_ptrs(self._h, None, 0)
        buf = (_vp * max(1, n))()
        self._lib.anakin_net_weight_ptrs(self._h, buf, n)
        return [int(buf[i] or 0) for i in range(n)]

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.anakin_net_destroy(self._h)
            self._h = _vp()


def weight_arena_stats():
    """(device bytes, entries, hits, misses) of the process-wide packed-weight arena."""
    lib = load()
    e, h, m = _sz(), _sz(), _sz()
    b = lib.anakin_weight_arena_stats(C.byref(e), C.byref(h), C.byref(m))
    return int(b), e.value, h.value, m.value


def weight_arena_set_receive(on):
    """Receive mode: Nets initialised while it is on allocate their packed-weight images without building them; the
    images then arrive with weight_arena_import (multi-GPU replicas, see anakin_b200/dist.py)."""
    load().anakin_weight_arena_set_receive(1 if on else 0)


def weight_arena_flat_bytes(device):
    return int(load().anakin_weight_arena_flat_bytes(device))


def weight_arena_export(device, dev_ptr, nbytes):
    _check(load().anakin_weight_arena_export(device, _vp(dev_ptr), nbytes), "weight_arena_export")


def weight_arena_import(device, dev_ptr, nbytes):
    _check(load().anakin_weight_arena_import(device, _vp(dev_ptr), nbytes), "weight_arena_import")


class Worker:
    """Worker<NV, P>: thread pool of per-thread Nets, optionally one GPU per thread."""

    def __init__(self, model_path, precision="fp32", threads=1, devices=(), batch=0):
        self._lib = load()
        self._h = _vp()
        devs = (_i * max(1, len(devices)))(*devices) if devices else None
        _check(self._lib.anakin_worker_create(model_path.encode(), PRECISIONS[precision], threads, devs,
                                              len(devices), batch, C.byref(self._h)), "Worker")

    def sync_prediction(self, x_nchw, out_count):
        a = np.ascontiguousarray(x_nchw, np.float32)
        out = np.empty(out_count, np.float32)
        _check(self._lib.anakin_worker_sync_prediction(self._h, a.ctypes.data_as(_vp), a.size,
                                                       out.ctypes.data_as(_vp), out.size), "sync_prediction")
        return out

    def wait_ready(self):
        _check(self._lib.anakin_worker_wait_ready(self._h), "Worker init")

    def async_prediction_ptr(self, in_ptr, in_count, out_ptr, out_count):
        """Queue one request on caller-owned (pinned) fp32 host buffers; pair with async_get_result()."""
        _check(self._lib.anakin_worker_async_prediction(self._h, _vp(in_ptr), in_count, _vp(out_ptr), out_count),
               "async_prediction")

    def async_get_result(self):
        _check(self._lib.anakin_worker_async_get_result(self._h), "async_get_result")

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.anakin_worker_destroy(self._h)
            self._h = _vp()
