"""FaceAnaStreams — FaceAna.run for many concurrent video streams on one GPU (additive API).

The reference serves one stream per FaceAna instance and keeps the temporal state (previous frame, track boxes,
GroupTrack / One-Euro history; facer.py:28-50, lk.py:6-91) in Python.  Here one object owns S streams: each call takes
one frame per stream and runs ONE detector forward and ONE landmark forward for all of them, the state lives on the
device (csrc/mpipe.cu, csrc/temporal.cu) and two batches can be in flight so that frame uploads overlap compute.

    fa = FaceAnaStreams(n_streams=16, top_k=4)
    results = fa.run(frames)            # list of S lists of {'box','kps','scores'} - what S FaceAna.run calls return
    # or, overlapped:
    fa.submit(frames_t0); fa.submit(frames_t1); r0 = fa.collect(); fa.submit(frames_t2); r1 = fa.collect(); ...
"""
import ctypes as C
import os
import pathlib

import numpy as np

from ... import runtime as rt
from .facer import get_cfg
from .onnx_model_base import ONNXEngine


class FaceAnaStreams:
    def __init__(self, n_streams, top_k=None, max_frame_hw=(2160, 3840), device="cuda"):
        cfg = get_cfg()['Skps']
        det_cfg, kps_cfg, tr_cfg = cfg['Detect'], cfg['Keypoints'], cfg['Trace']
        self.n_streams = int(n_streams)
        self.top_k = int(top_k if top_k is not None else det_cfg['topk'])
        root = pathlib.Path(__file__).resolve().parents[2]
        self.det = ONNXEngine(os.path.join(root, det_cfg['model_path']), device=device, max_batch=self.n_streams)
        self.kps = ONNXEngine(os.path.join(root, kps_cfg['model_path']), device=device,
                              max_batch=self.n_streams * self.top_k)
        self.n_points = int(kps_cfg['num_points'])
        self.lib = rt.load_library()
        pc = rt.PipelineCfg(score_thres=det_cfg['score_thrs'], iou_thres=det_cfg['iou_thrs'],
                            min_face=float(det_cfg['min_face']), top_k=self.top_k, track_iou=float(tr_cfg['iou_thres']),
                            alpha=float(tr_cfg['smooth_box']),
                            face_scale=float(np.float32(1 + 2 * kps_cfg['base_extend_range'][0])), kps_min_face=20.0,
                            max_h=int(max_frame_hw[0]), max_w=int(max_frame_hw[1]))
        h = C.c_void_p()
        rt.check(self.lib.skps_mpipe_create(self.det.handle, self.kps.handle, C.byref(pc), self.n_streams, C.byref(h)))
        self._h = h
        S, K, P = self.n_streams, self.top_k, self.n_points
        self._out = [dict(n=np.zeros(S, np.int32), box=np.zeros((S, K, 4), np.float64), kps=np.zeros((S, K, P, 2), np.float64),
                          sc=np.zeros((S, K, P), np.float32), det=np.zeros(S, np.int32)) for _ in range(2)]
        self._pending = []              # [(slot, n, keep-alive frames)]
        self._next = 0
        self.last_ran_detector = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.skps_mpipe_destroy(h)
            self._h = None

    def reset(self, stream=None):
        """FaceAna.reset (facer.py:200-208) for one stream, or for all of them."""
        while self._pending:
            self.collect()
        rt.check(self.lib.skps_mpipe_reset(self._h, -1 if stream is None else int(stream)))

    def submit(self, frames):
        """Enqueue one HxWx3 uint8 BGR frame per stream (frames[i] -> stream i, len(frames) <= n_streams).  At most two
        batches may be pending; results come back from collect() in submission order."""
        if len(self._pending) == 2:
            raise RuntimeError("FaceAnaStreams: two batches already in flight; call collect() first")
        n = len(frames)
        if not 0 < n <= self.n_streams:
            raise ValueError("expected 1..%d frames, got %d" % (self.n_streams, n))
        keep = []
        for f in frames:
            f = np.ascontiguousarray(f)
            if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise ValueError("expected HxWx3 uint8 BGR images, got %s %s" % (f.dtype, f.shape))
            keep.append(f)
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in keep])
        hw = np.array([[f.shape[0], f.shape[1]] for f in keep], np.int32)
        slot = self._next
        rt.check(self.lib.skps_mpipe_submit(self._h, slot, ptrs, hw.ctypes.data, n, 0))
        self._pending.append((slot, n, keep))
        self._next ^= 1

    def collect(self):
        """Results of the oldest pending batch: a list (one entry per stream) of lists of {'box','kps','scores'}."""
        if not self._pending:
            raise RuntimeError("FaceAnaStreams: nothing submitted")
        slot, n, _keep = self._pending.pop(0)
        o = self._out[slot]
        rt.check(self.lib.skps_mpipe_wait(self._h, slot, o["n"].ctypes.data, o["box"].ctypes.data, o["kps"].ctypes.data,
                                          o["sc"].ctypes.data, o["det"].ctypes.data))
        self.last_ran_detector = o["det"][:n].astype(bool)
        res = []
        for s in range(n):
            k = int(o["n"][s])
            res.append([{'box': o["box"][s, i].copy(), 'kps': o["kps"][s, i].copy(), 'scores': o["sc"][s, i].copy()}
                        for i in range(k)])
        return res

    def run(self, frames):
        """One frame per stream in, per-stream results out (blocking)."""
        self.submit(frames)
        return self.collect()
