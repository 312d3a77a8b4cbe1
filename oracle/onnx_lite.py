"""ORACLE (test infrastructure, never on the product path).

Minimal ONNX protobuf reader: decodes the wire format directly because the
`onnx` package is not installed in this image.  Only the message fields that the
two shipped graphs (yolov5n-0.5.onnx, kps_student.onnx) use are decoded.

Field numbers follow onnx.proto (ModelProto.graph=7; GraphProto.node=1,
initializer=5, input=11, output=12; NodeProto input=1, output=2, name=3,
op_type=4, attribute=5; AttributeProto name=1, f=2, i=3, s=4, t=5, floats=7,
ints=8; TensorProto dims=1, data_type=2, float_data=4, int64_data=7, name=8,
raw_data=9).

The reference reaches these graphs through onnxruntime
(/root/reference/Skps/core/api/onnx_model_base.py:14); this reader plus
oracle/onnx_exec.py restate what that session does.
"""
import struct
import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fno, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v):
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64}


def _tensor(buf):
    dims, dtype, name, raw = [], 1, "", None
    fdata, idata = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 2:
            dtype = v
        elif fno == 4:
            if wt == 2:
                fdata += list(struct.unpack("<%df" % (len(v) // 4), v))
            else:
                fdata.append(struct.unpack("<f", v)[0])
        elif fno == 7:
            idata += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dtype).copy()
    elif fdata:
        arr = np.asarray(fdata, dtype=np_dtype)
    elif idata:
        arr = np.asarray(idata, dtype=np_dtype)
    else:
        arr = np.zeros(0, dtype=np_dtype)
    return name, arr.reshape(dims)


def _attribute(buf):
    name, val = "", None
    floats, ints = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v).decode(errors="replace")
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            if wt == 2:
                floats += list(struct.unpack("<%df" % (len(v) // 4), v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            ints += _packed_varints(v) if wt == 2 else [_signed(v)]
    if val is None:
        val = ints if ints else (floats if floats else [])
    return name, val


class Node:
    __slots__ = ("op", "name", "inputs", "outputs", "attrs")

    def __init__(self, op, name, inputs, outputs, attrs):
        self.op, self.name, self.inputs, self.outputs, self.attrs = op, name, inputs, outputs, attrs

    def __repr__(self):
        return "Node(%s %s %s -> %s %s)" % (self.op, self.name, self.inputs, self.outputs,
                                            {k: (v if not isinstance(v, np.ndarray) else v.shape)
                                             for k, v in self.attrs.items()})


def _node(buf):
    ins, outs, name, op, attrs = [], [], "", "", {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            ins.append(bytes(v).decode())
        elif fno == 2:
            outs.append(bytes(v).decode())
        elif fno == 3:
            name = bytes(v).decode()
        elif fno == 4:
            op = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(v)
            attrs[k] = a
    return Node(op, name, ins, outs, attrs)


def _value_info_name(buf):
    for fno, wt, v in _fields(buf):
        if fno == 1:
            return bytes(v).decode()
    return ""


class Graph:
    def __init__(self, nodes, initializers, inputs, outputs):
        self.nodes = nodes
        self.initializers = initializers
        self.inputs = inputs      # graph inputs that are not initializers
        self.outputs = outputs


def load(path):
    with open(path, "rb") as f:
        data = memoryview(f.read())
    graph_buf = None
    for fno, wt, v in _fields(data):
        if fno == 7:
            graph_buf = v
    nodes, inits, inputs, outputs = [], {}, [], []
    for fno, wt, v in _fields(graph_buf):
        if fno == 1:
            nodes.append(_node(v))
        elif fno == 5:
            n, a = _tensor(v)
            inits[n] = a
        elif fno == 11:
            inputs.append(_value_info_name(v))
        elif fno == 12:
            outputs.append(_value_info_name(v))
    inputs = [i for i in inputs if i not in inits]
    return Graph(nodes, inits, inputs, outputs)
