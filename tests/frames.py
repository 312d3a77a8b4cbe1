"""Deterministic synthetic inputs shared by the golden generator, the parity tests
and bench.py.  numpy + cv2.imread only; the scaled
faces are produced with the oracle's bit-exact bilinear so the frames are
identical on every machine."""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TEST1 = os.path.join(_HERE, "golden", "test1.jpg")


def load_test1():
    import cv2
    img = cv2.imread(TEST1)
    assert img is not None and img.shape == (273, 410, 3)
    return img


def _resize(img, w, h):
    import sys
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle.host_ref import resize_linear_u8
    return resize_linear_u8(img, w, h)


def canvas_640(img=None):
    """Config 1: test1.jpg pasted at (y=183, x=115) on a 114-grey 640x640 canvas."""
    img = load_test1() if img is None else img
    c = np.full((640, 640, 3), 114, np.uint8)
    c[183:183 + 273, 115:115 + 410] = img
    return c


def _background(h, w):
    if (h, w) not in _BG_CACHE:
        _BG_CACHE[(h, w)] = _background_make(h, w)
    return _BG_CACHE[(h, w)].copy()


_BG_CACHE = {}


def _background_make(h, w):
    yy = (np.arange(h, dtype=np.int32)[:, None] * 40) // h
    xx = (np.arange(w, dtype=np.int32)[None, :] * 40) // w
    g = (94 + yy + xx).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2).copy()


_FACE_CACHE = {}


def multi_face_frame(h, w, grid, face_w, jitter=(0, 0), img=None):
    """`grid`=(rows, cols) copies of test1.jpg scaled to `face_w` wide on a smooth
    grey gradient.  Config 3: (1080,1920,(2,2),440); config 5: (2160,3840,(4,4),600)."""
    fh = int(round(273 * face_w / 410))
    if img is None:
        if face_w not in _FACE_CACHE:                    # the scaled face is the slow part (pure-numpy bilinear)
            _FACE_CACHE[face_w] = _resize(load_test1(), face_w, fh)
        face = _FACE_CACHE[face_w]
    else:
        face = _resize(img, face_w, fh)
    frame = _background(h, w)
    rows, cols = grid
    for r in range(rows):
        for c in range(cols):
            cy = int((r + 0.5) * h / rows) + jitter[1]
            cx = int((c + 0.5) * w / cols) + jitter[0]
            y0, x0 = cy - fh // 2, cx - face_w // 2
            frame[y0:y0 + fh, x0:x0 + face_w] = face
    return frame


def frame_1080p(jitter=(0, 0)):
    return multi_face_frame(1080, 1920, (2, 2), 440, jitter)


def frame_4k(jitter=(0, 0)):
    return multi_face_frame(2160, 3840, (4, 4), 600, jitter)


def crop_variants(n, seed=0):
    """Config 2 'realistic' set: the oracle crop of test1.jpg's face plus seeded
    variants (brightness +-20, shifts +-8 px, flips).  Returns (n,256,256,3) u8 BGR."""
    import sys
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle.host_ref import crop_face
    img = load_test1()
    box = np.array([153.4755, 49.7373, 306.6512, 234.0922], np.float32)
    base, _ = crop_face(img, box.copy())
    rng = np.random.default_rng(seed)
    out = np.empty((n, 256, 256, 3), np.uint8)
    out[0] = base
    for i in range(1, n):
        db = int(rng.integers(-20, 21))
        sx, sy = int(rng.integers(-8, 9)), int(rng.integers(-8, 9))
        flip = bool(rng.integers(0, 2))
        v = np.roll(base, (sy, sx), axis=(0, 1))
        if flip:
            v = v[:, ::-1]
        out[i] = np.clip(v.astype(np.int16) + db, 0, 255).astype(np.uint8)
    return out


def noise_crops(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 256, 256, 3), dtype=np.uint8)
