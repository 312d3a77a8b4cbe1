"""Generates tests/golden/*.npz by running the UNMODIFIED reference package from
/root/reference (read-only) with oracle/ort_shim standing in for onnxruntime.
Run once in the build container:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle.py) and the CUDA path
(tests/test_parity_gpu.py).  /root/reference does not exist on the GPU box, so
nothing at test time imports this file."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ort_shim"))
sys.path.insert(0, "/root/reference")
sys.path.append(ROOT)
sys.path.append(os.path.join(ROOT, "tests"))
sys.path.append(HERE)

from Skps import FaceAna  # noqa: E402  (the reference's own package)
import Skps as _ref_pkg  # noqa: E402
assert _ref_pkg.__file__.startswith("/root/reference"), _ref_pkg.__file__
import frames  # noqa: E402  (tests/frames.py)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Recorder:
    """Wraps the reference objects' methods to capture per-stage tensors."""

    def __init__(self, facer):
        self.f = facer
        self.log = {}
        det, kps = facer.face_detector, facer.face_landmark
        self._det_pre, self._det_model = det.preprocess, det.model
        self._det_nms = det.py_nms
        self._kps_pre, self._kps_model = kps.preprocess, kps.model
        det.preprocess = self.det_pre
        det.model = self.det_model
        det.py_nms = self.det_nms
        kps.preprocess = self.kps_pre
        kps.model = self.kps_model
        self.crops, self.details, self.kps_raw, self.kps_score = [], [], [], []

    def det_pre(self, image):
        x, rec = self._det_pre(image)
        self.log["letterbox"] = x.copy()
        self.log["recover"] = np.array(rec, np.float64)
        return x, rec

    def det_model(self, x):
        y = self._det_model(x)
        self.log["det_raw"] = np.array(y[0]).reshape(15120, 16).copy()
        return y

    def det_nms(self, bboxes, iou, score):
        kept = self._det_nms(bboxes, iou, score)
        # recover the kept row indices (rows are unique in practice)
        idx = [int(np.where((bboxes == k).all(axis=1))[0][0]) for k in kept]
        self.log["det_keep_idx"] = np.array(idx, np.int64)
        return kept

    def kps_pre(self, img, bbox, i):
        crop, detail = self._kps_pre(img, bbox, i)
        self.crops.append(crop.copy())
        self.details.append([int(d) for d in detail])
        return crop, detail

    def kps_model(self, x):
        out = self._kps_model(x)
        self.kps_raw.append(np.array(out[0]).reshape(-1).copy())
        self.kps_score.append(np.array(out[1]).reshape(-1).copy())
        return out

    def flush(self):
        d = dict(self.log)
        if self.crops:
            d["crops"] = np.stack(self.crops)
            d["details"] = np.array(self.details, np.int64)
            d["kps_raw"] = np.stack(self.kps_raw)
            d["kps_score"] = np.stack(self.kps_score)
        self.log = {}
        self.crops, self.details, self.kps_raw, self.kps_score = [], [], [], []
        return d


def pack_result(res):
    if len(res) == 0:
        return {"n": np.array(0)}
    return {"n": np.array(len(res)),
            "box": np.stack([np.asarray(r["box"], np.float64) for r in res]),
            "kps": np.stack([r["kps"] for r in res]).astype(np.float32),
            "scores": np.stack([r["scores"] for r in res]).astype(np.float32)}


def run_case(name, frames_list, top_k=None):
    facer = FaceAna()
    if top_k is not None:
        facer.top_k = top_k
    rec = Recorder(facer)
    out = {}
    for t, fr in enumerate(frames_list):
        res = facer.run(fr.copy())
        st = rec.flush()
        if "letterbox" in st:
            lb = st.pop("letterbox")
            st["letterbox_sha"] = np.array(sha(lb))
            st["letterbox_sum"] = np.array(float(lb.astype(np.float64).sum()))
            raw = st.pop("det_raw")
            st["det_raw_sha"] = np.array(sha(raw))
            cand = np.where(raw[:, 4] > 0.25)[0]
            st["det_cand_idx"] = cand
            st["det_cand_rows"] = raw[cand]
        st["frame_sha"] = np.array(sha(fr))
        for k, v in pack_result(res).items():
            st["res_" + k] = v
        for k, v in st.items():
            out["f%d_%s" % (t, k)] = v
    out["n_frames"] = np.array(len(frames_list))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, [int(out["f%d_res_n" % t]) for t in range(len(frames_list))], os.path.getsize(path))


from make_golden_frames import video_frames  # noqa: E402


def main():
    run_case("test1", [frames.load_test1()])
    run_case("canvas640", [frames.canvas_640()])
    run_case("video1080", video_frames())
    run_case("uhd4k_top5", [frames.frame_4k()])
    run_case("uhd4k_top16", [frames.frame_4k()], top_k=16)


if __name__ == "__main__":
    main()
