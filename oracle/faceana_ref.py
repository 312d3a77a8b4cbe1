"""ORACLE (test infrastructure, never on the product path).

End-to-end CPU restatement of `FaceAna.run` (/root/reference/Skps/core/api/facer.py:52-85)
built from oracle.host_ref + oracle.onnx_exec.  Used as the checker for the CUDA
path and as the `cpu_baseline` / `--impl reference` arm of bench.py ("port" kind:
the reference's real engine, onnxruntime, is absent from this image).
"""
import os
import numpy as np

from . import host_ref as H
from .onnx_exec import Session

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                    "peppa_pig_face_landmark_b200")
DET_ONNX = os.path.join(_PKG, "pretrained", "yolov5n-0.5.onnx")
KPS_ONNX = os.path.join(_PKG, "pretrained", "kps_student.onnx")


class DetectorRef:
    """face_detector.py:11-42."""

    def __init__(self, score_thrs=0.5, iou_thrs=0.3, in_hw=(384, 640)):
        self.net = Session(DET_ONNX)
        self.score_thrs, self.iou_thrs, self.in_hw = score_thrs, iou_thrs, in_hw

    def raw(self, image):
        x, recover = H.letterbox(image, *self.in_hw)
        return self.net.run(x)[0], recover, x

    def __call__(self, image, return_indices=False):
        raw, recover, _ = self.raw(image)
        kept, idx = H.detect_post(raw, recover, self.iou_thrs, self.score_thrs)
        return (kept, idx) if return_indices else kept


class LandmarkRef:
    """face_landmark.py:14-115 (batch-1 loop over faces, exactly as the reference)."""

    def __init__(self, extend0=0.2, in_hw=(256, 256)):
        self.net = Session(KPS_ONNX)
        self.extend0, self.in_hw = extend0, in_hw

    def forward_crops(self, crops_u8):
        """crops (N,256,256,3) u8 BGR -> (N,98,2) normalised xy, (N,98) scores."""
        xy, sc = [], []
        for c in crops_u8:
            x = c.transpose(2, 0, 1).astype(np.float32) / np.float32(255.)
            out, score = self.net.run(x[None])
            xy.append(np.asarray(out).reshape(-1)[:196].reshape(98, 2))
            sc.append(np.asarray(score).reshape(-1))
        if not xy:
            return np.zeros((0, 98, 2), np.float32), np.zeros((0, 98), np.float32)
        return np.stack(xy), np.stack(sc)

    def __call__(self, image, bboxes):
        lms, scs = [], []
        for b in bboxes:
            crop, detail = H.crop_face(image, b, self.in_hw, self.extend0)
            xy, sc = self.forward_crops(crop[None])
            lms.append(H.landmark_post(xy[0], detail))
            scs.append(sc[0])
        return np.array(lms), np.array(scs)


class FaceAnaRef:
    """facer.py:25-208."""

    def __init__(self, top_k=5, min_face=1600, iou_thres=0.5, alpha=0.3):
        self.det = DetectorRef()
        self.kps = LandmarkRef()
        self.trace = H.GroupTrackRef(iou_thres)
        self.top_k, self.min_face, self.iou_thres, self.alpha = top_k, min_face, iou_thres, alpha
        self.diff_thres = 5
        self.reset()

    def reset(self):
        self.track_box = None
        self.previous_image = None

    def diff_frames(self, prev, image):
        """facer.py:98-118."""
        if prev is None:
            return True
        d = np.abs(prev.astype(np.int16) - image.astype(np.int16)).astype(np.uint8)
        diff = np.sum(d) / prev.shape[0] / prev.shape[1] / 3.
        return bool(diff > self.diff_thres)

    def run(self, image):
        if self.diff_frames(self.previous_image, image):
            boxes = self.det(image)
            self.previous_image = image
            boxes = H.judge_boxs(self.track_box, boxes, self.iou_thres, self.alpha)
            self.trace.prev = None
        else:
            boxes = self.track_box
            self.previous_image = image
        boxes = H.sort_and_filter(boxes, self.min_face, self.top_k)
        boxes_return = np.array(boxes)
        landmarks, states = self.kps(image, boxes)
        landmarks = self.trace.calculate(image, landmarks)
        track = []
        for i in range(landmarks.shape[0]):
            track.append([np.min(landmarks[i][:, 0]), np.min(landmarks[i][:, 1]),
                          np.max(landmarks[i][:, 0]), np.max(landmarks[i][:, 1])])
        tmp_box = np.array(track)
        self.track_box = H.judge_boxs(boxes_return, tmp_box, self.iou_thres, self.alpha)
        return [{"box": self.track_box[i], "kps": landmarks[i], "scores": states[i]}
                for i in range(len(self.track_box))]
