"""Re-targets a Skps landmark export to another square input size.

`tools/convert_to_onnx.py --img_size S` (/root/reference/TRAIN/face_landmark/tools/convert_to_onnx.py:15-60)
traces the same network at S x S: every convolution is size-agnostic, and the only constants the tracer bakes
in are (a) the ASPP pooling branch's `F.interpolate(x, size=size)` target (model.py:58-61) = S/16 and (b) the
heat-map side S/4 used by postp (model.py:511-554: `idx % W`, `idx // W`, `/ W`, `/ H`).  Patching those
reproduces the graph the exporter would have written for the README's @128 variants from the shipped @256
file (the weights are whatever the source file holds).
"""
import numpy as np

from .onnx_loader import OnnxNode, load_onnx
from .onnx_writer import save_onnx


def retarget_input_size(src_onnx, dst_onnx, size):
    g = load_onnx(src_onnx)
    shp = g.input_shapes[g.inputs[0]]
    old = int(shp[2])
    if shp[2] != shp[3] or size % 32 or size <= 0:
        raise ValueError("retarget_input_size: square inputs, multiples of 32 (got %s -> %d)" % (shp, size))
    old_hm, new_hm, old_p, new_p = old // 4, size // 4, old // 16, size // 16
    users = {}
    for n in g.nodes:
        for i in n.inputs:
            users.setdefault(i, []).append(n)
    has_argmax = any(n.op == "ArgMax" for n in g.nodes)
    nodes, patched = [], 0
    for n in g.nodes:
        attrs = dict(n.attrs)
        if n.op == "Constant":
            v = np.asarray(attrs["value"])
            us = users.get(n.outputs[0], [])
            if has_argmax and v.size == 1 and float(v.reshape(-1)[0]) == old_hm and us and \
                    all(u.op in ("Mod", "Div") and u.inputs[1] == n.outputs[0] for u in us):
                attrs["value"] = np.full(v.shape, new_hm, v.dtype)               # postp: % W, // W, / W, / H
                patched += 1
            elif v.dtype == np.int64 and v.shape == (2,) and list(v) == [old_p, old_p] and "fm_pool" in n.name:
                attrs["value"] = np.array([new_p, new_p], np.int64)                # ASPP pooling branch target size
                patched += 1
        nodes.append(OnnxNode(n.op, n.name, list(n.inputs), list(n.outputs), attrs))
    if patched != 5:
        raise ValueError("retarget_input_size: expected 5 size constants in a Skps landmark export, found %d" % patched)
    save_onnx(dst_onnx, nodes, g.weights, [(g.inputs[0], [1, 3, size, size])],
              [(o, [1, 196] if o == "output" else [1, 98]) for o in g.outputs])
    return dst_onnx
