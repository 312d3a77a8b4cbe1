"""Loads the reference's shipped .onnx files unchanged (weights + node list).

The reference hands these files to onnxruntime
(/root/reference/Skps/core/api/onnx_model_base.py:14).  Neither `onnx` nor
`onnxruntime` exists in this image, so the protobuf wire format is walked
directly; only the fields the two shipped graphs use are understood.
"""
import struct
from collections import namedtuple

import numpy as np

OnnxNode = namedtuple("OnnxNode", "op name inputs outputs attrs")
OnnxGraph = namedtuple("OnnxGraph", "nodes weights inputs outputs input_shapes")

_NP = {1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64}


class _Reader:
    """Cursor over one protobuf message body."""

    def __init__(self, mv):
        self.mv = mv
        self.p = 0

    def more(self):
        return self.p < len(self.mv)

    def varint(self):
        x = 0
        s = 0
        while True:
            b = self.mv[self.p]
            self.p += 1
            x |= (b & 127) << s
            if b < 128:
                return x
            s += 7

    def field(self):
        key = self.varint()
        num, wire = key >> 3, key & 7
        if wire == 0:
            return num, wire, self.varint()
        if wire == 2:
            n = self.varint()
            v = self.mv[self.p:self.p + n]
            self.p += n
            return num, wire, v
        n = 8 if wire == 1 else 4
        if wire not in (1, 5):
            raise ValueError("bad wire type")
        v = self.mv[self.p:self.p + n]
        self.p += n
        return num, wire, v


def _i64(x):
    return x - (1 << 64) if x >> 63 else x


def _ints(wire, v):
    if wire == 0:
        return [_i64(v)]
    r = _Reader(v)
    out = []
    while r.more():
        out.append(_i64(r.varint()))
    return out


def _floats(wire, v):
    if wire == 5:
        return list(struct.unpack("<f", v))
    return list(np.frombuffer(bytes(v), "<f4"))


def _parse_tensor(mv):
    r = _Reader(mv)
    dims, dt, name, raw, fl, il = [], 1, "", None, [], []
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            dims += _ints(wire, v)
        elif num == 2:
            dt = v
        elif num == 4:
            fl += _floats(wire, v)
        elif num == 7:
            il += _ints(wire, v)
        elif num == 8:
            name = bytes(v).decode()
        elif num == 9:
            raw = bytes(v)
    if raw is not None:
        arr = np.frombuffer(raw, _NP[dt]).copy()
    else:
        arr = np.asarray(fl if fl else il, dtype=_NP[dt])
    return name, arr.reshape(dims)


def _parse_attr(mv):
    r = _Reader(mv)
    name, scalar, fl, il = "", None, [], []
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            name = bytes(v).decode()
        elif num == 2:
            scalar = struct.unpack("<f", v)[0]
        elif num == 3:
            scalar = _i64(v)
        elif num == 4:
            scalar = bytes(v).decode(errors="replace")
        elif num == 5:
            scalar = _parse_tensor(v)[1]
        elif num == 7:
            fl += _floats(wire, v)
        elif num == 8:
            il += _ints(wire, v)
    if scalar is None:
        scalar = il if il else fl
    return name, scalar


def _parse_node(mv):
    r = _Reader(mv)
    ins, outs, name, op, attrs = [], [], "", "", {}
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            ins.append(bytes(v).decode())
        elif num == 2:
            outs.append(bytes(v).decode())
        elif num == 3:
            name = bytes(v).decode()
        elif num == 4:
            op = bytes(v).decode()
        elif num == 5:
            k, a = _parse_attr(v)
            attrs[k] = a
    return OnnxNode(op, name, ins, outs, attrs)


def _value_info(mv):
    """ValueInfoProto -> (name, [dims]); dims of unknown size come back as -1."""
    r = _Reader(mv)
    name, dims = "", []
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            name = bytes(v).decode()
        elif num == 2:                      # TypeProto
            rt = _Reader(v)
            while rt.more():
                n2, w2, v2 = rt.field()
                if n2 != 1:                 # tensor_type
                    continue
                rtt = _Reader(v2)
                while rtt.more():
                    n3, w3, v3 = rtt.field()
                    if n3 != 2:             # shape
                        continue
                    rs = _Reader(v3)
                    while rs.more():
                        n4, w4, v4 = rs.field()
                        if n4 != 1:         # dim
                            continue
                        rd = _Reader(v4)
                        val = -1
                        while rd.more():
                            n5, w5, v5 = rd.field()
                            if n5 == 1:
                                val = _i64(v5)
                        dims.append(val)
    return name, dims


def _first_string(mv):
    r = _Reader(mv)
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            return bytes(v).decode()
    return ""


def load_onnx(path):
    with open(path, "rb") as f:
        mv = memoryview(f.read())
    r = _Reader(mv)
    graph = None
    while r.more():
        num, wire, v = r.field()
        if num == 7:
            graph = v
    if graph is None:
        raise ValueError("%s: no graph in model" % path)
    nodes, weights, gin, gout, shapes = [], {}, [], [], {}
    r = _Reader(graph)
    while r.more():
        num, wire, v = r.field()
        if num == 1:
            nodes.append(_parse_node(v))
        elif num == 5:
            k, a = _parse_tensor(v)
            weights[k] = a
        elif num == 11:
            nm, dims = _value_info(v)
            gin.append(nm)
            shapes[nm] = dims
        elif num == 12:
            gout.append(_first_string(v))
    gin = [g for g in gin if g not in weights]
    return OnnxGraph(nodes, weights, gin, gout, {g: shapes[g] for g in gin})
