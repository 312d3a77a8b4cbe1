"""Teacher (HRNet-w18 + Decoder + heat-map head) inference graph, written as the .onnx file that
`tools/convert_to_onnx.py --model teacher --img_size {128,256}` would produce.

Reference: /root/reference/TRAIN/face_landmark/lib/core/base_trainer/model.py:302-345 (TeacherNet),
:212-244 (Decoder), :64-98 (ASPP), :133-196 (DecoderBlock), :117-130 (SCSEModule), :511-554 (postp);
the encoder is timm's public `hrnet_w18` with `features_only=True, out_indices=[0,1,2,3]`
(stem 64@s2, then the 'incre' bottlenecks 128@s4, 256@s8, 512@s16).

The reference ships no teacher weights (README model table only), so the weights here are SYNTHETIC:
Kaiming-normal(fan_out) convolutions as in model.py:199-209 / timm's HRNet init, drawn from a seeded
numpy generator, with every BatchNorm's running statistics calibrated on a few seeded synthetic crops
so that activations stay O(1) as in a trained network.  BatchNorms that follow a convolution are folded
into it exactly as the exporter's eval-mode constant folding does (the student file shows the same:
Conv nodes with `onnx::Conv_*` weight+bias pairs).  The arg-max decode tail (postp) is taken node for
node from the shipped kps_student.onnx, where it was traced from the same function.

The file goes through the normal ONNXEngine path (lowering.py -> engine); nothing here runs on the
forward path.  torch is used as a CPU calculator for the calibration pass only.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from .onnx_loader import OnnxNode, load_onnx
from .onnx_writer import save_onnx

_HERE = os.path.dirname(os.path.abspath(__file__))
STUDENT_ONNX = os.path.join(_HERE, "pretrained", "kps_student.onnx")

# timm hrnet_w18 configuration (public model definition)
_STAGES = [
    dict(modules=1, block="bottleneck", blocks=(4,), channels=(64,)),
    dict(modules=1, block="basic", blocks=(4, 4), channels=(18, 36)),
    dict(modules=4, block="basic", blocks=(4, 4, 4), channels=(18, 36, 72)),
    dict(modules=3, block="basic", blocks=(4, 4, 4, 4), channels=(18, 36, 72, 144)),
]
_HEAD_CHANNELS = (32, 64, 128, 256)      # 'incre' bottlenecks, expansion 4
_BN_EPS = 1e-5


def synthetic_crops(n, size, seed):
    """Smooth random fields + pixel noise, uint8 (n, size, size, 3): stand-ins for face crops."""
    rng = np.random.default_rng(seed)
    low = torch.from_numpy(rng.uniform(0, 255, (n, 3, 8, 8)).astype(np.float32))
    img = F.interpolate(low, size=(size, size), mode="bilinear", align_corners=False).numpy()
    img = img + rng.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8).transpose(0, 2, 3, 1).copy()


class _Builder:
    def __init__(self, seed, calib_nchw):
        self.rng = np.random.default_rng(seed)
        self.nodes, self.inits = [], {}
        self.val = {"input": calib_nchw}          # calibration activations, torch float32 NCHW
        self.params = 0                             # trainable parameters of the un-folded torch module
        self.macs = 0                               # conv multiply-accumulates per sample (emitted nodes)

    # ---- raw parameter draws ------------------------------------------------------------------------
    def _kaiming(self, cout, cin_g, k):
        std = np.sqrt(2.0 / (cout * k * k))         # fan_out, relu gain (model.py:205)
        return (self.rng.standard_normal((cout, cin_g, k, k)) * std).astype(np.float32)

    def _default_bias(self, cout, cin_g, k):
        bound = 1.0 / np.sqrt(cin_g * k * k)        # nn.Conv2d default bias init
        return self.rng.uniform(-bound, bound, cout).astype(np.float32)

    def count_only(self, cin, cout, k, groups=1, bias=False, bn=False):
        """Parameters of a module that exists in the torch model but is dead in the exported graph."""
        self.params += cout * (cin // groups) * k * k + (cout if bias else 0) + (2 * cout if bn else 0)

    def _emit(self, op, scope, ins, attrs, value, n_out=1):
        out = scope + "_output_0"
        self.nodes.append(OnnxNode(op, scope, list(ins), [out], attrs))
        self.val[out] = value
        return out

    # ---- layers ---------------------------------------------------------------------------------------
    def conv(self, x, cout, k, scope, stride=1, dil=1, groups=1, bias=False, bn=False, act=None, pad=None):
        xv = self.val[x]
        cin = xv.shape[1]
        pad = dil * (k - 1) // 2 if pad is None else pad
        w = self._kaiming(cout, cin // groups, k)
        b = self._default_bias(cout, cin // groups, k) if bias else None
        self.params += w.size + (cout if bias else 0) + (2 * cout if bn else 0)
        wt = torch.from_numpy(w)
        y = F.conv2d(xv, wt, torch.from_numpy(b) if b is not None else None, stride, pad, dil, groups)
        if bn:
            # BatchNorm (gamma 1, beta 0) with running statistics = statistics of the calibration batch, folded
            mean = y.mean(dim=(0, 2, 3))
            var = y.var(dim=(0, 2, 3), unbiased=False)
            inv = (1.0 / torch.sqrt(var + _BN_EPS)).numpy().astype(np.float32)
            w = (w * inv[:, None, None, None]).astype(np.float32)
            b0 = b if b is not None else np.zeros(cout, np.float32)
            b = ((b0 - mean.numpy().astype(np.float32)) * inv).astype(np.float32)
            y = F.conv2d(xv, torch.from_numpy(w), torch.from_numpy(b), stride, pad, dil, groups)
        wn, bn_ = scope + ".weight", scope + ".bias"
        self.inits[wn] = w
        ins = [x, wn]
        if b is not None:
            self.inits[bn_] = b
            ins.append(bn_)
        out = self._emit("Conv", scope + "/Conv", ins,
                         dict(dilations=[dil, dil], group=groups, kernel_shape=[k, k], pads=[pad] * 4,
                              strides=[stride, stride]), y)
        self.macs += int(y.shape[1] * y.shape[2] * y.shape[3] * (cin // groups) * k * k)
        if act == "relu":
            out = self.relu(out, scope)
        elif act == "sigmoid":
            out = self._emit("Sigmoid", scope + "/Sigmoid", [out], {}, torch.sigmoid(self.val[out]))
        return out

    def relu(self, x, scope):
        return self._emit("Relu", scope + "/Relu", [x], {}, torch.relu(self.val[x]))

    def add(self, a, b, scope):
        return self._emit("Add", scope + "/Add", [a, b], {}, self.val[a] + self.val[b])

    def mul(self, a, b, scope):
        return self._emit("Mul", scope + "/Mul", [a, b], {}, self.val[a] * self.val[b])

    def concat(self, xs, scope):
        return self._emit("Concat", scope + "/Concat", xs, dict(axis=1), torch.cat([self.val[x] for x in xs], 1))

    def const(self, scope, arr):
        out = scope + "_output_0"
        self.nodes.append(OnnxNode("Constant", scope, [], [out], dict(value=np.asarray(arr))))
        return out

    def resize_scale(self, x, factor, mode, scope):
        """nn.Upsample(scale_factor) / F.interpolate(scale_factor=2, mode='bilinear') as opset-12 Resize."""
        sc = self.const(scope + "/Constant", np.array([1, 1, factor, factor], np.float32))
        roi = self.const(scope + "/Constant_1", np.zeros(0, np.float32))
        xv = self.val[x]
        if mode == "nearest":
            y = xv.repeat_interleave(factor, 2).repeat_interleave(factor, 3)
            attrs = dict(coordinate_transformation_mode="asymmetric", cubic_coeff_a=-0.75, mode="nearest",
                         nearest_mode="floor")
        else:
            y = F.interpolate(xv, scale_factor=factor, mode="bilinear", align_corners=False)
            attrs = dict(coordinate_transformation_mode="half_pixel", cubic_coeff_a=-0.75, mode="linear",
                         nearest_mode="floor")
        return self._emit("Resize", scope + "/Resize", [x, roi, sc], attrs, y)

    def resize_to(self, x, hw, scope):
        """F.interpolate(x, size=size) (nearest) of the ASPP pooling branch (model.py:58-61)."""
        xv = self.val[x]
        sizes = self.const(scope + "/Constant", np.array([1, xv.shape[1], hw[0], hw[1]], np.int64))
        roi = self.const(scope + "/Constant_1", np.zeros(0, np.float32))
        scales = self.const(scope + "/Constant_2", np.zeros(0, np.float32))
        y = xv.expand(xv.shape[0], xv.shape[1], hw[0], hw[1]).contiguous()
        return self._emit("Resize", scope + "/Resize", [x, roi, scales, sizes],
                          dict(coordinate_transformation_mode="asymmetric", cubic_coeff_a=-0.75, mode="nearest",
                               nearest_mode="floor"), y)

    def gap(self, x, scope):
        return self._emit("GlobalAveragePool", scope + "/GlobalAveragePool", [x], {},
                          self.val[x].mean(dim=(2, 3), keepdim=True))

    def batchnorm(self, x, scope):
        xv = self.val[x]
        c = xv.shape[1]
        mean = xv.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
        var = xv.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32)
        self.params += 2 * c
        names = [scope + s for s in (".weight", ".bias", ".running_mean", ".running_var")]
        for nm, a in zip(names, (np.ones(c, np.float32), np.zeros(c, np.float32), mean, var)):
            self.inits[nm] = a
        y = F.batch_norm(xv, torch.from_numpy(mean), torch.from_numpy(var), None, None, False, 0.0, _BN_EPS)
        return self._emit("BatchNormalization", scope + "/BatchNormalization", [x] + names,
                          dict(epsilon=float(np.float32(_BN_EPS)), momentum=float(np.float32(0.9))), y)

    # ---- HRNet blocks (timm hrnet.py / resnet.py) -------------------------------------------------------
    def bottleneck(self, x, planes, scope, downsample):
        y = self.conv(x, planes, 1, scope + "/conv1", bn=True, act="relu")
        y = self.conv(y, planes, 3, scope + "/conv2", bn=True, act="relu")
        y = self.conv(y, planes * 4, 1, scope + "/conv3", bn=True)
        sc = self.conv(x, planes * 4, 1, scope + "/downsample/downsample.0", bn=True) if downsample else x
        return self.relu(self.add(y, sc, scope), scope + "/act3")

    def basic(self, x, planes, scope):
        y = self.conv(x, planes, 3, scope + "/conv1", bn=True, act="relu")
        y = self.conv(y, planes, 3, scope + "/conv2", bn=True)
        return self.relu(self.add(y, x, scope), scope + "/act2")

    def hr_module(self, xs, channels, blocks, scope, live_outputs):
        """HighResolutionModule (fuse_method SUM, multi_scale_output True)."""
        nb = len(xs)
        for i in range(nb):
            for j in range(blocks[i]):
                xs[i] = self.basic(xs[i], channels[i], "%s/branches.%d/branches.%d.%d" % (scope, i, i, j))
        outs = []
        for i in range(nb):
            if i >= live_outputs:
                # exists in the torch module, dead in the exported graph (nothing reads this branch)
                for j in range(nb):
                    if j > i:
                        self.count_only(channels[j], channels[i], 1, bn=True)
                    for k in range(i - j):
                        self.count_only(channels[j], channels[i] if k == i - j - 1 else channels[j], 3, bn=True)
                outs.append(None)
                continue
            y = None
            for j in range(nb):
                fs = "%s/fuse_layers.%d/fuse_layers.%d.%d" % (scope, i, i, j)
                if j == i:
                    t = xs[j]
                elif j > i:
                    t = self.conv(xs[j], channels[i], 1, fs + "/0", bn=True)
                    t = self.resize_scale(t, 2 ** (j - i), "nearest", fs + "/2")
                else:
                    t = xs[j]
                    for k in range(i - j):
                        last = k == i - j - 1
                        t = self.conv(t, channels[i] if last else channels[j], 3, "%s/%d/0" % (fs, k), stride=2,
                                      bn=True, act=None if last else "relu")
                y = t if y is None else self.add(y, t, "%s/fuse_add.%d.%d" % (scope, i, j))
            outs.append(self.relu(y, "%s/fuse_act.%d" % (scope, i)))
        return outs


def build_teacher_onnx(path, size=256, seed=0, n_calib=4, student_onnx=STUDENT_ONNX):
    """Write the synthetic-weight Teacher graph to `path`; returns {'params', 'macs', 'nodes'}."""
    if size % 32 or size < 64:
        raise ValueError("teacher input size must be a multiple of 32 (README variants: 128, 256)")
    calib = torch.from_numpy(synthetic_crops(n_calib, size, seed + 1).transpose(0, 3, 1, 2).astype(np.float32) / 255.0)
    b = _Builder(seed, calib)
    E = "/teacher/encoder"
    with torch.no_grad():
        x = b.conv("input", 64, 3, E + "/conv1", stride=2, bn=True, act="relu")
        feat2 = x
        x = b.conv(x, 64, 3, E + "/conv2", stride=2, bn=True, act="relu")
        for j in range(4):
            x = b.bottleneck(x, 64, "%s/layer1/layer1.%d" % (E, j), downsample=(j == 0))
        ys = [x]
        for si in range(1, 4):
            st = _STAGES[si]
            ch, prev = st["channels"], _STAGES[si - 1]["channels"]
            prev_out = [256] if si == 1 else list(prev)
            xs = []
            for i, c in enumerate(ch):
                ts = "%s/transition%d/transition%d.%d" % (E, si, si, i)
                if i < len(prev_out):
                    xs.append(ys[i] if prev_out[i] == c else b.conv(ys[i], c, 3, ts + "/0", bn=True, act="relu"))
                else:
                    xs.append(b.conv(ys[-1], c, 3, ts + "/0/0", stride=2, bn=True, act="relu"))
            for m in range(st["modules"]):
                last_module = si == 3 and m == st["modules"] - 1
                xs = b.hr_module(xs, ch, st["blocks"], "%s/stage%d/stage%d.%d" % (E, si + 1, si + 1, m),
                                 live_outputs=3 if last_module else len(ch))
            ys = xs
        feats = []
        for i in range(3):
            feats.append(b.bottleneck(ys[i], _HEAD_CHANNELS[i], "%s/incre_modules.%d/incre_modules.%d.0" % (E, i, i),
                                      downsample=True))
        # incre_modules.3 (144 -> 1024) is built by timm but its feature (index 4) is not requested
        for cin, cout, k in ((144, 256, 1), (256, 256, 3), (256, 1024, 1), (144, 1024, 1)):
            b.count_only(cin, cout, k, bn=True)
        encx2, encx4, encx8, encx16 = feat2, feats[0], feats[1], feats[2]

        D = "/teacher/decoder"
        hw = tuple(b.val[encx16].shape[2:])
        f1 = b.conv(encx16, 64, 1, D + "/aspp/conv1")
        f2 = b.conv(encx16, 64, 3, D + "/aspp/conv2", dil=2)
        f3 = b.conv(encx16, 64, 3, D + "/aspp/conv3", dil=4)
        fp = b.gap(encx16, D + "/aspp/fm_pool/pool/pool.0")
        fp = b.conv(fp, 64, 1, D + "/aspp/fm_pool/pool/pool.1", bn=True, act="relu")
        fp = b.resize_to(fp, hw, D + "/aspp/fm_pool")
        x = b.concat([f1, f2, f3, fp], D + "/aspp")
        x = b.relu(b.batchnorm(x, D + "/aspp/bn_act/bn_act.0"), D + "/aspp/bn_act/bn_act.1")
        x = b.conv(x, 256, 1, D + "/aspp/project/project.0", bn=True, act="relu")

        def decoder_block(x, skip, cout, scope, attention, second):
            x = b.resize_scale(x, 2, "linear", scope)
            x = b.concat([x, skip], scope)
            c = b.val[x].shape[1]
            x = b.conv(x, c, 3, scope + "/conv1/conv1.0/conv_dw/conv_dw.0", groups=c, bias=True, bn=True)
            x = b.conv(x, cout, 1, scope + "/conv1/conv1.0/conv_pw", bn=True, act="relu")
            if second:
                x = b.conv(x, cout, 3, scope + "/conv2/conv2.0", bias=True, bn=True, act="relu")
            if attention:
                a = scope + "/attention2"
                g = b.gap(x, a + "/cSE/cSE.0")
                g = b.conv(g, cout // 4, 1, a + "/cSE/cSE.1", bias=True, act="relu")
                g = b.conv(g, cout, 1, a + "/cSE/cSE.3", bias=True, act="sigmoid")
                s = b.conv(x, 1, 1, a + "/sSE/sSE.0", bias=True, act="sigmoid")
                x = b.add(b.mul(x, g, a), b.mul(x, s, a + "/Mul_1"), a)
            return x

        x = decoder_block(x, encx8, 256, D + "/upsampler1", attention=True, second=False)
        x = decoder_block(x, encx4, 128, D + "/upsampler2", attention=False, second=True)
        hm = b.conv(x, 98 * 3, 1, "/teacher/hm", bias=True)
        b.count_only(640, 7, 1, bias=True)              # self.fc (pose/cls head, unused at inference)

    # arg-max decode tail (model.py:511-554), traced nodes of the shipped student export, re-rooted on our heat map
    if size != 256:
        from .graph_tools import retarget_input_size
        student_onnx = retarget_input_size(student_onnx, path + ".tail.tmp", size)    # postp constants for size/4 maps
    sg = load_onnx(student_onnx)
    if size != 256:
        os.remove(student_onnx)
    hm_idx = [i for i, n in enumerate(sg.nodes) if n.name == "/student/hm/Conv"][0]
    src = sg.nodes[hm_idx].outputs[0]
    tail = sg.nodes[hm_idx + 1:]
    made = {o for n in tail for o in n.outputs}
    need = {i for n in tail for i in n.inputs if i and i != src and i not in made and i not in sg.weights}
    for n in sg.nodes[:hm_idx]:                      # shared scalar Constants the exporter hoisted to the front
        if n.op == "Constant" and n.outputs[0] in need:
            b.nodes.append(OnnxNode(n.op, n.name, [], list(n.outputs), dict(n.attrs)))
            need.discard(n.outputs[0])
    if need:
        raise ValueError("decode tail depends on non-constant tensors: %s" % sorted(need))
    for n in tail:
        b.nodes.append(OnnxNode(n.op, n.name, [hm if i == src else i for i in n.inputs], list(n.outputs), dict(n.attrs)))
        for i in n.inputs:
            if i in sg.weights:
                b.inits[i] = sg.weights[i]
    save_onnx(path, b.nodes, b.inits, [("input", [1, 3, size, size])],
              [(o, [1, 196] if o == "output" else [1, 98]) for o in sg.outputs], graph_name="teacher_synthetic")
    return dict(params=b.params, macs=b.macs, nodes=len(b.nodes))


def default_teacher_path(size=256, seed=0):
    """Cache location of the generated file (git-ignored; rebuilt on demand, ~46 MB)."""
    d = os.path.join(_HERE, "pretrained", "_generated")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "kps_teacher_synthetic_%d_seed%d.onnx" % (size, seed))


def ensure_teacher_onnx(size=256, seed=0):
    p = default_teacher_path(size, seed)
    if not os.path.exists(p):
        tmp = p + ".tmp%d" % os.getpid()
        build_teacher_onnx(tmp, size, seed)
        os.replace(tmp, p)
    return p
