"""FaceAna — same public surface as /root/reference/Skps/core/api/facer.py:25-208:
`FaceAna(verbose=False)`, `.run(image) -> [{'box','kps','scores'}, ...]`, `.reset()`.

Per frame the GPU does, with no host round trip in between (skps_pipeline_run):
letterbox -> yolov5-face -> NMS -> judge_boxs(track) -> sort_and_filter -> crops -> landmark net
-> de-normalise.  The host keeps what is stateful and tiny: the frame-diff decision, One-Euro
landmark smoothing (GroupTrack) and the EMA of the track boxes (facer.py:71-82)."""
import ctypes as C
import logging
import os
import pathlib

import numpy as np
import yaml

from ... import runtime as rt
from ...logger.logger import logger
from ..smoother.lk import EmaFilter, GroupTrack
from .face_detector import FaceDetector, letterbox_geometry
from .face_landmark import FaceLandmark


def get_cfg():
    root_path = pathlib.Path(__file__).resolve().parents[2]
    cfg_path = os.path.join(root_path, 'config', 'Skps.yml')
    with open(cfg_path, encoding="UTF-8") as f:
        return yaml.load(f, Loader=yaml.FullLoader)


class FaceAna():
    def __init__(self, verbose=False, top_k=None, max_frame_hw=(2160, 3840)):
        if verbose:
            logger.setLevel(logging.DEBUG)
        cfg = get_cfg()
        self.top_k = int(top_k if top_k is not None else cfg['Skps']['Detect']['topk'])
        self.face_detector = FaceDetector(cfg['Skps']['Detect'])
        self.face_landmark = FaceLandmark(cfg['Skps']['Keypoints'], max_faces=self.top_k)
        self.trace = GroupTrack(cfg['Skps']['Trace'])
        logger.info('model init done!')
        self.track_box = None
        self.previous_image = None
        self.previous_box = None
        self.diff_thres = 5
        self.min_face = cfg['Skps']['Detect']['min_face']
        self.iou_thres = cfg['Skps']['Trace']['iou_thres']
        self.alpha = cfg['Skps']['Trace']['smooth_box']
        self.filter = EmaFilter(self.alpha)

        self.lib = rt.load_library()
        det, kps = self.face_detector, self.face_landmark
        pc = rt.PipelineCfg(score_thres=det.score_thrs, iou_thres=det.iou_thrs, min_face=float(self.min_face),
                            top_k=self.top_k, track_iou=float(self.iou_thres), alpha=float(self.alpha),
                            face_scale=kps.face_scale, kps_min_face=float(kps.min_face),
                            max_h=int(max_frame_hw[0]), max_w=int(max_frame_hw[1]))
        h = C.c_void_p()
        rt.check(self.lib.skps_pipeline_create(det.model.handle, kps.model.handle, C.byref(pc), C.byref(h)))
        self._pipe = h
        self._stream = det.model.stream
        K, P = self.top_k, kps.keypoints_num
        self._n = C.c_int32(0)
        self._ndet = C.c_int32(0)
        self._boxes = np.zeros((K, 4), np.float32)
        self._kps = np.zeros((K, P, 2), np.float32)
        self._scores = np.zeros((K, P), np.float32)
        self._det_idx = np.zeros((FaceDetector.MAX_DET,), np.int32)
        self._det_rows = np.zeros((FaceDetector.MAX_DET, 16), np.float32)
        self._have_prev = False
        self.last_det_idx = None       # kept detector rows of the last detector run (parity checks)
        self.last_det_rows = None

    def __del__(self):
        h = getattr(self, "_pipe", None)
        if h is not None and h.value:
            self.lib.skps_pipeline_destroy(h)
            self._pipe = None

    # ------------------------------------------------------------------ facer.py:52-85
    def run(self, image):
        image = np.ascontiguousarray(image)
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("expected an HxWx3 uint8 BGR image, got %s %s" % (image.dtype, image.shape))
        H, W = image.shape[:2]
        run_det = self.diff_frames(self.previous_image, image)     # stages the frame on the device
        self.previous_image = image
        det = self.face_detector
        in_h, in_w = det.input_size[0], det.input_size[1]
        scale, rw, rh, top, left = letterbox_geometry(H, W, in_h, in_w)
        track = self.track_box
        n_track = 0 if track is None else int(len(track))
        track32 = None
        if n_track:
            track32 = np.ascontiguousarray(np.asarray(track)[:, :4], dtype=np.float32)
        if not run_det and n_track == 0:
            # facer.py:61 with an empty/None track: nothing to do (the reference would fail on None)
            boxes_return = np.zeros((0, 4), np.float32)
            landmarks, states = np.array([]), np.array([])
            rt.check(self.lib.skps_pipeline_commit_frame(self._pipe, H, W))     # the staged frame is the new "previous"
        else:
            rt.check(self.lib.skps_pipeline_run(
                self._pipe, None, H, W, 0, 1 if run_det else 0, rw, rh, top, left, float(scale),
                rt.ptr(track32), n_track, C.byref(self._n), self._boxes.ctypes.data, self._kps.ctypes.data,
                self._scores.ctypes.data, C.byref(self._ndet), self._det_idx.ctypes.data,
                self._det_rows.ctypes.data, self._stream.cuda_stream))
            n = self._n.value
            boxes_return = self._boxes[:n].copy()
            landmarks = self._kps[:n].copy() if n else np.array([])
            states = self._scores[:n].copy() if n else np.array([])
            if run_det:
                self.last_det_idx = self._det_idx[:self._ndet.value].astype(np.int64)
                self.last_det_rows = self._det_rows[:self._ndet.value].copy()
        if run_det:
            self.trace.previous_landmarks_set = None

        landmarks = self.trace.calculate(image, landmarks)

        track = []
        for i in range(landmarks.shape[0]):
            track.append([np.min(landmarks[i][:, 0]), np.min(landmarks[i][:, 1]),
                          np.max(landmarks[i][:, 0]), np.max(landmarks[i][:, 1])])
        tmp_box = np.array(track)
        self.track_box = self.judge_boxs(boxes_return, tmp_box)
        return self.to_dict(self.track_box, landmarks, states)

    def to_dict(self, bboxes, kps, states):
        return [{'box': bboxes[i], 'kps': kps[i], 'scores': states[i]} for i in range(len(bboxes))]

    def diff_frames(self, previous_frame, image):
        """facer.py:98-118: mean |prev - cur| > 5 -> run the detector.  The sum is taken on the GPU
        against the previous frame kept in HBM; the frame uploaded here is reused by run()."""
        H, W = image.shape[:2]
        d = C.c_double(0.0)
        rt.check(self.lib.skps_pipeline_frame_diff(self._pipe, image.ctypes.data, H, W, 0, C.byref(d),
                                                   self._stream.cuda_stream))
        if previous_frame is None or d.value < 0:
            return True
        return bool(d.value > self.diff_thres)

    def sort_and_filter(self, bboxes):
        """facer.py:120-142 (host copy of what skps_select_faces does on the device)."""
        if len(bboxes) < 1:
            return []
        area = (bboxes[:, 2] - bboxes[:, 0]) * (bboxes[:, 3] - bboxes[:, 1])
        keep = area > self.min_face
        area, bboxes = area[keep], bboxes[keep, :]
        if bboxes.shape[0] > self.top_k:
            order = area.argsort()[-self.top_k:][::-1]
            return np.array([bboxes[i] for i in order])
        return np.array(bboxes)

    def judge_boxs(self, previuous_bboxs, now_bboxs):
        """facer.py:144-189."""
        if previuous_bboxs is None:
            return now_bboxs
        out = []
        for now in now_bboxs:
            matched = None
            for prev in previuous_bboxs:
                if _iou(now, prev) > self.iou_thres:
                    matched = prev
                    break
            out.append(now[0:4] if matched is None else self.smooth(now, matched))
        return np.array(out)

    def smooth(self, now_box, previous_box):
        return self.filter(now_box[:4], previous_box[:4])

    def reset(self):
        """facer.py:200-208."""
        self.track_box = None
        self.previous_image = None
        self.previous_box = None
        rt.check(self.lib.skps_pipeline_reset(self._pipe))


def _iou(rec1, rec2):
    s1 = (rec1[2] - rec1[0]) * (rec1[3] - rec1[1])
    s2 = (rec2[2] - rec2[0]) * (rec2[3] - rec2[1])
    x1, y1 = max(rec1[0], rec2[0]), max(rec1[1], rec2[1])
    x2, y2 = min(rec1[2], rec2[2]), min(rec1[3], rec2[3])
    inter = max(0, x2 - x1) * max(0, y2 - y1)
    return inter / (s1 + s2 - inter)
