"""ORACLE (test infrastructure, never on the product path).

numpy restatement of the reference's host-side arithmetic around the two
networks.  Every function cites the reference lines it follows.  numpy >= 2
(NEP 50) promotion rules are what this container runs the reference with, so the
dtypes below are written out explicitly to match them.

Pinned by tests/test_oracle.py against (a) cv2.resize for the fixed-point
bilinear, (b) tests/golden/*.npz produced by running the unmodified reference
(tests/golden/make_golden.py).
"""
import math
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- cv2.resize(INTER_LINEAR) on u8
def _linear_taps(dst, src, clamp_x):
    """Per destination index: source index and the two 11-bit integer weights.
    OpenCV imgproc resize.cpp (third-party; reached from face_detector.py:53 and
    face_landmark.py:97): scale is 1/(dst/src) in double, the fractional offset
    is rounded to float32, weights are rint(w*2048) as int16."""
    inv = float(dst) / float(src)
    scale = 1.0 / inv
    idx = np.empty(dst, np.int32)
    w0 = np.empty(dst, np.int32)
    w1 = np.empty(dst, np.int32)
    for d in range(dst):
        f = F32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = F32(f - F32(s))
        if clamp_x:
            if s < 0:
                s, f = 0, F32(0)
            if s >= src - 1:
                s, f = src - 1, F32(0)
        idx[d] = s
        w0[d] = int(np.rint(F32(F32(1.0) - f) * F32(2048.0)))
        w1[d] = int(np.rint(f * F32(2048.0)))
    return idx, w0, w1


def resize_linear_u8(src, dw, dh):
    """Bit-exact restatement of cv2.resize(src, (dw, dh)) (INTER_LINEAR, uint8, HxWxC)."""
    sh, sw = src.shape[:2]
    xi, a0, a1 = _linear_taps(dw, sw, True)
    yi, b0, b1 = _linear_taps(dh, sh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    s = src.astype(np.int32)
    # horizontal pass: 11-bit weights, int32 rows
    hrow = s[:, xi, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0 = np.clip(yi, 0, sh - 1)
    r1 = np.clip(yi + 1, 0, sh - 1)
    h0 = hrow[r0] >> 4
    h1 = hrow[r1] >> 4
    out = (((b0[:, None, None] * h0) >> 16) + ((b1[:, None, None] * h1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- detector pre/post
def letterbox_geometry(h, w, in_h=384, in_w=640):
    """face_detector.py:51-62: scale, resized size and the four pad widths."""
    scale = min(in_h / h, in_w / w)
    rw, rh = int(w * scale), int(h * scale)
    dh = (in_h - rh) / 2
    dw = (in_w - rw) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return scale, rw, rh, top, bottom, left, right


def letterbox(image_bgr, in_h=384, in_w=640):
    """face_detector.py:45-71.  BGR u8 HxWx3 -> (1,3,in_h,in_w) f32 RGB/255, [scale,left,top]."""
    h, w = image_bgr.shape[:2]
    scale, rw, rh, top, bottom, left, right = letterbox_geometry(h, w, in_h, in_w)
    rgb = image_bgr[:, :, ::-1]
    small = resize_linear_u8(rgb, rw, rh)
    canvas = np.full((rh + top + bottom, rw + left + right, 3), 114, np.uint8)
    canvas[top:top + rh, left:left + rw] = small
    x = canvas.transpose(2, 0, 1).astype(F32)
    x /= F32(255.0)
    return x[None], [scale, left, top]


def xywh2xyxy(x):
    """face_detector.py:73-80 (float32)."""
    y = x.copy()
    half_w = x[:, 2] / F32(2)
    half_h = x[:, 3] / F32(2)
    y[:, 0] = x[:, 0] - half_w
    y[:, 1] = x[:, 1] - half_h
    y[:, 2] = x[:, 0] + half_w
    y[:, 3] = x[:, 1] + half_h
    return y


def nms(rows, iou_thres, score_thres):
    """face_detector.py:95-136.  rows (N,16) xyxy.  Returns (kept_rows, kept_indices
    into the input).  Candidate order is np.argsort(score)[::-1] exactly as the
    reference; survivors of each round are those with iou < thres."""
    cand = np.where(rows[:, 4] > score_thres)[0]
    b = rows[cand]
    order = np.argsort(b[:, 4])[::-1]
    keep = []
    while order.shape[0] > 0:
        cur = order[0]
        keep.append(cur)
        rest = order[1:]
        area = (b[cur, 2] - b[cur, 0]) * (b[cur, 3] - b[cur, 1])
        xx1 = np.maximum(b[cur, 0], b[rest, 0])
        yy1 = np.maximum(b[cur, 1], b[rest, 1])
        xx2 = np.minimum(b[cur, 2], b[rest, 2])
        yy2 = np.minimum(b[cur, 3], b[rest, 3])
        inter = np.maximum(0, yy2 - yy1) * np.maximum(0, xx2 - xx1)
        other = (b[rest, 3] - b[rest, 1]) * (b[rest, 2] - b[rest, 0])
        iou = inter / (area + other - inter)
        order = rest[np.where(iou < iou_thres)[0]]
    keep = np.asarray(keep, dtype=np.int64)
    return b[keep], cand[keep]


def scale_coords(xyxy, recover):
    """face_detector.py:82-93 (in place on float32 (K,4))."""
    scale, dx, dy = recover
    xyxy[:, 0] -= dx
    xyxy[:, 1] -= dy
    xyxy[:, 2] -= dx
    xyxy[:, 3] -= dy
    xyxy /= scale          # python float -> weak scalar -> float32 division
    return xyxy


def detect_post(raw, recover, iou_thres=0.3, score_thres=0.5):
    """face_detector.py:31-37 on the raw (15120,16) network output."""
    out = np.array(raw, dtype=F32).reshape(-1, 16)
    out[:, :4] = xywh2xyxy(out[:, :4])
    kept, idx = nms(out, iou_thres, score_thres)
    kept[:, :4] = scale_coords(kept[:, :4], recover)
    return kept, idx


# --------------------------------------------------------------------------- landmark pre/post
def crop_geometry(bbox, extend0=0.2, min_face=20):
    """face_landmark.py:74-93: float32 box -> (add, x1, y1, x2, y2) ints in the
    zero-padded frame (pad = add on each side).  None if the face is too small."""
    b = np.asarray(bbox[:4], dtype=F32).copy()
    bw = b[2] - b[0]
    bh = b[3] - b[1]
    if bw <= min_face or bh <= min_face:
        return None
    add = int(max(bw, bh))
    b += add                                      # float32 += python int
    face_w = F32((1 + 2 * extend0) * bw)          # python float is weak -> float32 product
    cx = (b[0] + b[2]) // 2
    cy = (b[1] + b[3]) // 2
    half = face_w // 2
    sq = np.array([cx - half, cy - half, cx + half, cy + half], dtype=F32).astype(np.int32)
    return add, int(sq[0]), int(sq[1]), int(sq[2]), int(sq[3])


def crop_face(image, bbox, out_hw=(256, 256), extend0=0.2, min_face=20):
    """face_landmark.py:66-104.  Returns (crop u8 out_h x out_w x 3, [h, w, y1, x1, add])."""
    g = crop_geometry(bbox, extend0, min_face)
    if g is None:
        return None, None
    add, x1, y1, x2, y2 = g
    H, W = image.shape[:2]
    padded = np.zeros((H + 2 * add, W + 2 * add, 3), np.uint8)
    padded[add:add + H, add:add + W] = image
    crop = padded[y1:y2, x1:x2, :]
    h, w = crop.shape[:2]
    out = resize_linear_u8(crop, out_hw[1], out_hw[0])
    return out, [h, w, y1, x1, add]


def landmark_post(xy_norm, detail):
    """face_landmark.py:106-115.  numpy>=2: float32 * python-int stays float32, then
    `+ np.int32 scalar` promotes to float64, result stored back as float32."""
    h, w, y1, x1, add = detail
    lm = np.asarray(xy_norm, dtype=F32).reshape(-1, 2).copy()
    lm[:, 0] = ((lm[:, 0] * F32(w)).astype(np.float64) + float(x1) - float(add)).astype(F32)
    lm[:, 1] = ((lm[:, 1] * F32(h)).astype(np.float64) + float(y1) - float(add)).astype(F32)
    return lm


def heatmap_decode(hm):
    """model.py:511-554 / kps_student.onnx nodes 201-410.  hm (294,64,64) f32 ->
    (xy (98,2) normalised, score (98,))."""
    hm = np.asarray(hm, dtype=F32).reshape(294, -1)
    n = hm.shape[1]
    side = int(round(math.sqrt(n)))
    idx = np.argmax(hm[:98], axis=1)           # first maximum
    ar = np.arange(98)
    score = hm[ar, idx]
    ox = hm[98 + ar, idx]
    oy = hm[196 + ar, idx]
    x = ((idx % side).astype(F32) + ox) / F32(side)
    y = ((idx // side).astype(F32) + oy) / F32(side)
    return np.stack([x, y], 1).astype(F32), score.astype(F32)


# --------------------------------------------------------------------------- FaceAna host logic
def iou_xyxy(r1, r2):
    """facer.py:151-170."""
    s1 = (r1[2] - r1[0]) * (r1[3] - r1[1])
    s2 = (r2[2] - r2[0]) * (r2[3] - r2[1])
    x1 = max(r1[0], r2[0]); y1 = max(r1[1], r2[1])
    x2 = min(r1[2], r2[2]); y2 = min(r1[3], r2[3])
    inter = max(0, x2 - x1) * max(0, y2 - y1)
    return inter / (s1 + s2 - inter)


def judge_boxs(prev, now, iou_thres=0.5, alpha=0.3):
    """facer.py:144-189 + lk.py:155-162.  IoU-match `now` against `prev`; matched
    rows become alpha*now[:4]+(1-alpha)*prev[:4], unmatched rows become now[:4]."""
    if prev is None:
        return now
    res = []
    for i in range(now.shape[0]):
        hit = False
        for j in range(prev.shape[0]):
            if iou_xyxy(now[i], prev[j]) > iou_thres:
                res.append(alpha * now[i][:4] + (1 - alpha) * prev[j][:4])
                hit = True
                break
        if not hit:
            res.append(now[i][0:4])
    return np.array(res)


def sort_and_filter(boxes, min_face=1600, top_k=5):
    """facer.py:120-142."""
    if len(boxes) < 1:
        return []
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    sel = area > min_face
    area = area[sel]
    boxes = boxes[sel, :]
    if boxes.shape[0] > top_k:
        picked = area.argsort()[-top_k:][::-1]
        return np.array([boxes[p] for p in picked])
    return np.array(boxes)


def one_euro(x, x_prev, dx_prev, min_cutoff=0.15, beta=0.8, d_cutoff=1):
    """lk.py:93-149 with t_e = 1."""
    def sf(cutoff):
        r = 2 * math.pi * cutoff * 1
        return r / (r + 1)
    a_d = sf(d_cutoff)
    dx = np.sqrt(np.sum((x - x_prev) ** 2, axis=1))
    dxp = np.sqrt(np.sum(dx_prev ** 2, axis=1))
    dx_hat = a_d * dx + (1 - a_d) * dxp
    cutoff = min_cutoff + beta * np.abs(dx_hat)
    r = 2 * math.pi * cutoff * 1
    a = (r / (r + 1))[:, None]
    a = np.array(a)
    a[dx < 0.002] = 0.01
    return a * x + (1 - a) * x_prev


class GroupTrackRef:
    """lk.py:6-91."""

    def __init__(self, iou_thres=0.5):
        self.prev = None
        self.prev_dx = None
        self.iou_thres = iou_thres

    @staticmethod
    def _rect(p):
        return [np.min(p[:, 0]), np.min(p[:, 1]), np.max(p[:, 0]), np.max(p[:, 1])]

    def calculate(self, img, now):
        h, w = img.shape[:2]
        scale = [w, h]
        if self.prev is None or self.prev.shape[0] == 0:
            self.prev = now
            pdx = np.zeros_like(now)
            result = now
        else:
            result, pdx = [], []
            for i in range(now.shape[0]):
                miss = True
                for j in range(self.prev.shape[0]):
                    if iou_xyxy(self._rect(now[i]), self._rect(self.prev[j])) > self.iou_thres:
                        f = one_euro(now[i] / scale, self.prev[j] / scale, self.prev_dx[j] / scale) * scale
                        result.append(f)
                        pdx.append(self.prev[j] - f)
                        miss = False
                        break
                if miss:
                    result.append(now[i])
                    pdx.append(np.zeros_like(now[i]))
        result = np.array(result)
        self.prev = result
        self.prev_dx = np.array(pdx)
        return result
