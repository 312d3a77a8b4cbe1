"""Temporal smoothing of landmarks and boxes — host-side numpy, O(K*98) per frame, stateful.
Same classes and call signatures as /root/reference/Skps/core/smoother/lk.py (GroupTrack :6-91,
OneEuroFilter :105-149, EmaFilter :155-162)."""
import math

import numpy as np


def _alpha(cutoff, t_e=1.0):
    r = 2 * math.pi * cutoff * t_e
    return r / (r + 1)


def _blend(a, x, x_prev):
    return a * x + (1 - a) * x_prev


def _bbox_of(points):
    return [np.min(points[:, 0]), np.min(points[:, 1]), np.max(points[:, 0]), np.max(points[:, 1])]


def _iou(r1, r2):
    a1 = (r1[2] - r1[0]) * (r1[3] - r1[1])
    a2 = (r2[2] - r2[0]) * (r2[3] - r2[1])
    w = max(0, min(r1[2], r2[2]) - max(r1[0], r2[0]))
    h = max(0, min(r1[3], r2[3]) - max(r1[1], r2[1]))
    inter = w * h
    return inter / (a1 + a2 - inter)


class OneEuroFilter:
    """One-Euro filter with unit time step; the derivative is the per-point displacement norm."""

    def __init__(self, dx0=0.0, min_cutoff=0.15, beta=0.8, d_cutoff=1):
        self.min_cutoff, self.beta, self.d_cutoff = min_cutoff, beta, d_cutoff

    def __call__(self, x, x_prev, dx_prev):
        speed = np.sqrt(np.sum((x - x_prev) ** 2, axis=1))
        speed_prev = np.sqrt(np.sum(dx_prev ** 2, axis=1))
        speed_hat = _blend(_alpha(self.d_cutoff), speed, speed_prev)
        a = _alpha(self.min_cutoff + self.beta * np.abs(speed_hat))
        a = np.expand_dims(a, -1)
        a[speed < 0.002] = 0.01            # nearly static points are held almost fixed
        self.dx_prev = speed_hat
        return _blend(a, x, x_prev)


class EmaFilter:
    def __init__(self, alpha):
        self.alpha = alpha

    def __call__(self, p_now, p_previous):
        return _blend(self.alpha, p_now, p_previous)


class GroupTrack:
    """Matches each face's landmark set to last frame's by bounding-box IoU and filters it."""

    def __init__(self, cfg):
        self.old_frame = None
        self.previous_landmarks_set = None
        self.with_landmark = True
        self.thres = cfg['pixel_thres']
        self.iou_thres = cfg['iou_thres']
        self.filter = OneEuroFilter()

    def iou(self, p_set0, p_set1):
        return _iou(_bbox_of(p_set0), _bbox_of(p_set1))

    def smooth(self, now_landmarks, previous_landmarks, previous_df):
        return self.filter(now_landmarks, previous_landmarks, previous_df)

    def calculate(self, img, now_landmarks_set):
        h, w = img.shape[0], img.shape[1]
        scale = [w, h]
        prev = self.previous_landmarks_set
        if prev is None or prev.shape[0] == 0:
            self.previous_landmarks_set = now_landmarks_set
            self.previous_dx = np.zeros_like(now_landmarks_set)
            return now_landmarks_set
        result, deltas = [], []
        for cur in now_landmarks_set:
            match = None
            for j in range(prev.shape[0]):
                if self.iou(cur, prev[j]) > self.iou_thres:
                    match = j
                    break
            if match is None:
                result.append(cur)
                deltas.append(np.zeros_like(cur))
            else:
                f = self.smooth(cur / scale, prev[match] / scale, self.previous_dx[match] / scale) * scale
                result.append(f)
                deltas.append(prev[match] - f)
        result = np.array(result)
        self.previous_landmarks_set = result
        self.previous_dx = np.array(deltas)
        return result
