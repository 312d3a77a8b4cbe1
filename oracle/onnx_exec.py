"""ORACLE (test infrastructure, never on the product path).

Node-by-node executor for the two ONNX graphs the reference ships, on PyTorch CPU
ops in fp32 (or fp64 as a tie-breaker).  It stands in for
`onnxruntime.InferenceSession(...).run` at
/root/reference/Skps/core/api/onnx_model_base.py:14,23 — onnxruntime itself is
un-vendored, unpinned (setup.py:21) and absent from this image, so the published
ONNX operator semantics (opset 12) are restated here op by op.

Cross-check available here: cv2.dnn loads the detector graph (not the student
graph); see tests/test_oracle.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_lite


class Session:
    """Executes a graph.  `run(feed)` returns the list of graph outputs (numpy)."""

    def __init__(self, path, dtype=torch.float32):
        self.graph = onnx_lite.load(path)
        self.dtype = dtype
        self.consts = {}
        for k, v in self.graph.initializers.items():
            t = torch.from_numpy(np.ascontiguousarray(v))
            if t.dtype == torch.float32:
                t = t.to(dtype)
            self.consts[k] = t
        self.input_name = self.graph.inputs[0]

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _ints(t):
        return [int(x) for x in t.reshape(-1).tolist()]

    def run(self, x, keep=None):
        """x: numpy NCHW.  keep: optional set of tensor names to also return (dict)."""
        env = dict(self.consts)
        env[self.input_name] = torch.from_numpy(np.ascontiguousarray(x)).to(self.dtype)
        kept = {}
        with torch.no_grad():
            for node in self.graph.nodes:
                outs = self._exec(node, env)
                if not isinstance(outs, (list, tuple)):
                    outs = [outs]
                for name, val in zip(node.outputs, outs):
                    env[name] = val
                    if keep is not None and (keep == "all" or name in keep):
                        kept[name] = val
        res = []
        for o in self.graph.outputs:
            v = env[o]
            if v.dtype in (torch.float64,):
                v = v.to(torch.float32) if self.dtype == torch.float32 else v
            res.append(v.numpy())
        if keep is not None:
            return res, kept
        return res

    # ------------------------------------------------------------------ ops
    def _exec(self, n, env):
        op = n.op
        a = n.attrs
        i = [env[k] if k != "" else None for k in n.inputs]
        if op == "Constant":
            t = torch.from_numpy(np.ascontiguousarray(a["value"]))
            if t.dtype == torch.float32:
                t = t.to(self.dtype)
            return t
        if op == "Conv":
            pads = a.get("pads", [0, 0, 0, 0])
            assert pads[0] == pads[2] and pads[1] == pads[3]
            return F.conv2d(i[0], i[1], i[2] if len(i) > 2 else None,
                            stride=a.get("strides", [1, 1]), padding=(pads[0], pads[1]),
                            dilation=a.get("dilations", [1, 1]), groups=a.get("group", 1))
        if op == "Relu":
            return torch.relu(i[0])
        if op == "Sigmoid":
            return torch.sigmoid(i[0])
        if op == "HardSigmoid":
            alpha = a.get("alpha", 0.2)
            beta = a.get("beta", 0.5)
            # ONNX: max(0, min(1, alpha*x + beta)); alpha is the f32 attribute value
            al = torch.tensor(np.float32(alpha)).to(i[0].dtype)
            be = torch.tensor(np.float32(beta)).to(i[0].dtype)
            return torch.clamp(i[0] * al + be, 0.0, 1.0)
        if op == "Mul":
            return i[0] * i[1]
        if op == "Add":
            return i[0] + i[1]
        if op == "Sub":
            return i[0] - i[1]
        if op == "Div":
            if not i[0].is_floating_point() and not i[1].is_floating_point():
                return torch.div(i[0], i[1], rounding_mode="trunc")
            return i[0] / i[1]
        if op == "Pow":
            return torch.pow(i[0], i[1].to(i[0].dtype))
        if op == "Mod":
            assert a.get("fmod", 0) == 0
            return torch.remainder(i[0], i[1])
        if op == "Concat":
            return torch.cat(i, dim=a["axis"])
        if op == "Reshape":
            shape = self._ints(i[1])
            src = list(i[0].shape)
            shape = [src[k] if s == 0 else s for k, s in enumerate(shape)]
            return i[0].reshape(shape)
        if op == "Transpose":
            return i[0].permute(a["perm"]).contiguous()
        if op == "Shape":
            return torch.tensor(list(i[0].shape), dtype=torch.int64)
        if op == "Gather":
            axis = a.get("axis", 0)
            idx = i[1]
            if idx.dim() == 0:
                return i[0].select(axis, int(idx))
            return torch.index_select(i[0], axis, idx.reshape(-1)).reshape(
                list(i[0].shape[:axis]) + list(idx.shape) + list(i[0].shape[axis + 1:]))
        if op == "Slice":
            data = i[0]
            starts, ends = self._ints(i[1]), self._ints(i[2])
            axes = self._ints(i[3]) if len(i) > 3 and i[3] is not None else list(range(len(starts)))
            steps = self._ints(i[4]) if len(i) > 4 and i[4] is not None else [1] * len(starts)
            idx = [slice(None)] * data.dim()
            for s, e, ax, st in zip(starts, ends, axes, steps):
                assert st > 0
                dim = data.shape[ax]
                if s < 0:
                    s += dim
                if e < 0:
                    e += dim
                s = max(0, min(s, dim))
                e = max(0, min(e, dim))
                idx[ax] = slice(s, e, st)
            return data[tuple(idx)]
        if op == "MaxPool":
            assert a["kernel_shape"] == [2, 2] and a["strides"] == [2, 2]
            return F.max_pool2d(i[0], 2, 2, 0, ceil_mode=bool(a.get("ceil_mode", 0)))
        if op == "Resize":
            mode = a["mode"]
            ctm = a.get("coordinate_transformation_mode", "half_pixel")
            x = i[0]
            if len(i) > 3 and i[3] is not None and i[3].numel() > 0:
                size = self._ints(i[3])[2:]
            else:
                sc = i[2].reshape(-1).tolist()
                size = [int(np.floor(x.shape[2] * sc[2])), int(np.floor(x.shape[3] * sc[3]))]
            if mode == "nearest":
                assert ctm == "asymmetric" and a.get("nearest_mode", "round_prefer_floor") == "floor"
                # asymmetric + floor: src = floor(dst * in/out)
                ys = (torch.arange(size[0]) * x.shape[2]) // size[0]
                xs = (torch.arange(size[1]) * x.shape[3]) // size[1]
                return x[:, :, ys][:, :, :, xs]
            assert mode == "linear" and ctm == "half_pixel"
            return F.interpolate(x, size=size, mode="bilinear", align_corners=False)
        if op == "ReduceMean":
            return i[0].mean(dim=a["axes"], keepdim=bool(a.get("keepdims", 1)))
        if op == "GlobalAveragePool":
            return i[0].mean(dim=(2, 3), keepdim=True)
        if op == "BatchNormalization":
            x, s, b, m, v = i
            eps = a.get("epsilon", 1e-5)
            return F.batch_norm(x, m, v, s, b, training=False, eps=eps)
        if op == "ArgMax":
            assert a.get("select_last_index", 0) == 0
            return self._argmax_first(i[0], a["axis"], bool(a.get("keepdims", 1)))
        if op == "GatherElements":
            return torch.gather(i[0], a.get("axis", 0) % i[0].dim(), i[1])
        if op == "Squeeze":
            out = i[0]
            for ax in sorted([ax % i[0].dim() for ax in a["axes"]], reverse=True):
                out = out.squeeze(ax)
            return out
        if op == "Unsqueeze":
            out = i[0]
            for ax in sorted(a["axes"]):
                out = out.unsqueeze(ax if ax >= 0 else ax + out.dim() + 1)
            return out
        if op == "Cast":
            to = {1: self.dtype, 7: torch.int64, 6: torch.int32, 9: torch.bool}[a["to"]]
            return i[0].to(to)
        if op == "Equal":
            return i[0] == i[1]
        if op == "Where":
            return torch.where(i[0], i[1], i[2])
        if op == "Expand":
            shape = self._ints(i[1])
            return i[0] * torch.ones(shape, dtype=i[0].dtype) if i[0].dtype != torch.bool \
                else i[0].expand(torch.broadcast_shapes(tuple(i[0].shape), tuple(shape)))
        if op == "ConstantOfShape":
            val = a.get("value", np.zeros(1, np.float32))
            t = torch.from_numpy(np.ascontiguousarray(val)).reshape(())
            return torch.full(self._ints(i[0]), t.item(), dtype=t.dtype)
        if op == "ScatterND":
            data, idx, upd = i[0].clone(), i[1], i[2]
            k = idx.shape[-1]
            flat_idx = idx.reshape(-1, k)
            flat_upd = upd.reshape((flat_idx.shape[0],) + tuple(data.shape[k:]))
            data[tuple(flat_idx[:, j] for j in range(k))] = flat_upd
            return data
        raise NotImplementedError(op)

    @staticmethod
    def _argmax_first(x, axis, keepdims):
        # first index of the maximum (ONNX select_last_index=0)
        m = x.max(dim=axis, keepdim=True).values
        n = x.shape[axis]
        shape = [1] * x.dim()
        shape[axis] = n
        ar = torch.arange(n).reshape(shape)
        idx = torch.where(x == m, ar, torch.full_like(ar, n)).min(dim=axis, keepdim=keepdims).values
        return idx
