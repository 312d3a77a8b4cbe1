"""Writes .onnx files (protobuf wire format, no `onnx` package in this image).

Counterpart of onnx_loader.py: what `torch.onnx.export(..., opset_version=12)` does at the end of
/root/reference/TRAIN/face_landmark/tools/convert_to_onnx.py:55-60, for graphs assembled in Python
(teacher_graph.py).  Field numbers follow onnx.proto: ModelProto ir_version=1, producer_name=2,
graph=7, opset_import=8; GraphProto node=1, name=2, initializer=5, input=11, output=12; NodeProto
input=1, output=2, name=3, op_type=4, attribute=5; AttributeProto name=1, f=2, i=3, s=4, t=5,
floats=7, ints=8, type=20; TensorProto dims=1, data_type=2, name=8, raw_data=9.
"""
import struct

import numpy as np

_DT = {np.dtype(np.float32): 1, np.dtype(np.int32): 6, np.dtype(np.int64): 7, np.dtype(np.bool_): 9,
       np.dtype(np.float64): 11}


def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno, wt):
    return _varint((fno << 3) | wt)


def _ld(fno, payload):
    return _key(fno, 2) + _varint(len(payload)) + payload


def _vi(fno, v):
    return _key(fno, 0) + _varint(int(v))


def tensor_proto(name, arr):
    arr = np.asarray(arr, order="C")      # (ascontiguousarray would turn 0-d into 1-d)
    out = b"".join(_vi(1, d) for d in arr.shape)
    out += _vi(2, _DT[arr.dtype])
    out += _ld(8, name.encode())
    out += _ld(9, arr.tobytes())
    return out


def _attr(name, val):
    out = _ld(1, name.encode())
    if isinstance(val, np.ndarray):
        return out + _ld(5, tensor_proto("", val)) + _vi(20, 4)
    if isinstance(val, (bool, int, np.integer)):
        return out + _vi(3, int(val)) + _vi(20, 2)
    if isinstance(val, (float, np.floating)):
        return out + _key(2, 5) + struct.pack("<f", float(val)) + _vi(20, 1)
    if isinstance(val, str):
        return out + _ld(4, val.encode()) + _vi(20, 3)
    val = list(val)
    if val and all(isinstance(v, (float, np.floating)) for v in val):
        return out + b"".join(_key(7, 5) + struct.pack("<f", float(v)) for v in val) + _vi(20, 6)
    return out + b"".join(_vi(8, int(v)) for v in val) + _vi(20, 7)


def node_proto(op, name, inputs, outputs, attrs):
    out = b"".join(_ld(1, i.encode()) for i in inputs)
    out += b"".join(_ld(2, o.encode()) for o in outputs)
    out += _ld(3, name.encode()) + _ld(4, op.encode())
    out += b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items())
    return out


def _value_info(name, dims, elem_type=1):
    shape = b"".join(_ld(1, _vi(1, d)) for d in dims)
    ttype = _vi(1, elem_type) + _ld(2, shape)
    return _ld(1, name.encode()) + _ld(2, _ld(1, ttype))


def save_onnx(path, nodes, initializers, inputs, outputs, graph_name="main_graph", opset=12, producer="skps_b200"):
    """nodes: iterable of objects with .op/.name/.inputs/.outputs/.attrs (onnx_loader.OnnxNode works);
    initializers: {name: ndarray}; inputs: [(name, dims)]; outputs: [(name, dims)]."""
    g = b"".join(_ld(1, node_proto(n.op, n.name, n.inputs, n.outputs, n.attrs)) for n in nodes)
    g += _ld(2, graph_name.encode())
    g += b"".join(_ld(5, tensor_proto(k, v)) for k, v in initializers.items())
    g += b"".join(_ld(11, _value_info(n, d)) for n, d in inputs)
    g += b"".join(_ld(12, _value_info(n, d)) for n, d in outputs)
    model = _vi(1, 7) + _ld(2, producer.encode()) + _ld(7, g) + _ld(8, _ld(1, b"") + _vi(2, opset))
    with open(path, "wb") as f:
        f.write(model)
