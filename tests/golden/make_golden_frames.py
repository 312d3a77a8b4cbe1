"""Frame sequences shared by make_golden.py and the tests (no reference import)."""
import numpy as np
import frames


def video_frames():
    f0 = frames.frame_1080p()
    f1 = f0.copy()
    f1[::7, ::5] = np.clip(f1[::7, ::5].astype(np.int16) + 3, 0, 255).astype(np.uint8)
    f2 = frames.frame_1080p(jitter=(6, -4))
    f3 = frames._background(1080, 1920)
    return [f0, f1, f1, f2, f3, f3]
