"""WFLW evaluation harness (peppa_pig_face_landmark_b200/eval/wflw.py) on a synthetic WFLW-format directory against the
reference's own procedure (TRAIN/face_landmark/tools/eval_WFLW.py:38-142 restated with the same OpenCV/numpy calls):
crops bit-exact, per-sample NME within float32 rounding.  The real dataset is not in the image; the harness is what is
under test, not the published accuracy."""
import os

import numpy as np
import pytest

import frames

pytestmark = pytest.mark.gpu


def _reference_crop(img, bbox, joints, S):
    """augmentationCropImage(is_training=False) + cv2.resize, eval_WFLW.py:38-80,113-124."""
    import cv2
    bbox = np.array(bbox).reshape(4, ).astype(np.float32)
    add = int(max(bbox[2] - bbox[0], bbox[3] - bbox[1]))
    bimg = cv2.copyMakeBorder(img, add, add, add, add, borderType=cv2.BORDER_CONSTANT)
    objcenter = np.array([(bbox[0] + bbox[2]) / 2., (bbox[1] + bbox[3]) / 2.])
    bbox += add
    objcenter += add
    joints = joints.copy()
    joints[:, :2] += add
    cwh = (bbox[2] - bbox[0]) * (1 + 0.2 * 2) // 2
    chh = (bbox[3] - bbox[1]) * (1 + 0.2 * 2) // 2
    min_x, max_x = int(objcenter[0] - cwh), int(objcenter[0] + cwh)
    min_y, max_y = int(objcenter[1] - chh), int(objcenter[1] + chh)
    joints[:, 0] -= min_x
    joints[:, 1] -= min_y
    crop = bimg[min_y:max_y, min_x:max_x, :]
    h, w, _ = crop.shape
    joints[:, 0] /= w
    joints[:, 1] /= h
    return cv2.resize(crop, (S, S)), joints


def test_wflw_harness_matches_reference_procedure(tmp_path):
    import cv2
    from peppa_pig_face_landmark_b200.eval import wflw
    from oracle.faceana_ref import LandmarkRef
    rng = np.random.default_rng(0)
    root = tmp_path
    os.makedirs(root / "WFLW_images" / "0--x")
    os.makedirs(root / "WFLW_annotations" / "list_98pt_test")
    base = frames.load_test1()
    lines = {"test": [], "pose": []}
    samples = []
    for i in range(6):
        img = np.clip(base.astype(np.int16) + int(rng.integers(-15, 15)), 0, 255).astype(np.uint8)
        if i % 2:
            img = np.ascontiguousarray(img[:, ::-1])
        fn = "0--x/im%d.png" % i
        cv2.imwrite(str(root / "WFLW_images" / fn), img)
        # "ground truth": a plausible landmark cloud over the face region of test1.jpg, near the image border for some
        cx, cy = (250 if not i % 2 else 160) + rng.uniform(-6, 6), 140 + rng.uniform(-6, 6)
        kps = np.stack([cx + 70 * rng.uniform(-1, 1, 98), cy + 85 * rng.uniform(-1, 1, 98)], 1).astype(np.float32)
        kps[60], kps[72] = (cx - 30, cy - 20), (cx + 30, cy - 20)
        line = " ".join("%.4f" % v for v in kps.reshape(-1)) + " 0 0 0 0 0 0 0 0 0 0 " + fn + "\n"
        lines["test" if i < 4 else "pose"].append(line)
        samples.append((img, kps))
    for k, v in lines.items():
        with open(root / "WFLW_annotations" / "list_98pt_test" / ("list_98pt_test_%s.txt" % k), "w") as f:
            f.writelines(v)
    ev = wflw.WFLWEvaluator(batch=4)                      # 6 samples: one full batch + a partial one
    got = ev.eval_lines(lines["test"] + lines["pose"], str(root / "WFLW_images"))
    ref_net = LandmarkRef()
    for i, (img, kps) in enumerate(samples):
        kq = np.array(["%.4f" % v for v in kps.reshape(-1)], dtype=np.float32).reshape(-1, 2)   # as parsed from the file
        bbox = [float(np.min(kq[:, 0])), float(np.min(kq[:, 1])), float(np.max(kq[:, 0])), float(np.max(kq[:, 1]))]
        crop, label = _reference_crop(img, bbox, kq, 256)
        x0, y0, w, h = wflw.eval_crop_rect(bbox)
        import torch
        from peppa_pig_face_landmark_b200 import runtime as rt
        d = torch.from_numpy(img).cuda()
        o = torch.zeros((256, 256, 3), dtype=torch.uint8, device="cuda")
        rt.check(rt.load_library().skps_crop_rect(d.data_ptr(), img.shape[0], img.shape[1], img.shape[1] * 3, x0, y0, w, h,
                                                  o.data_ptr(), 256, None))
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy(), crop), i                   # bit-exact with copyMakeBorder + slice + resize
        xy, _ = ref_net.forward_crops(crop[None])
        want = wflw.nme(label, np.asarray(xy).reshape(1, 98, 2))
        assert abs(got[i] - want) <= 2e-5 * max(1.0, abs(want)), (i, got[i], want)
    res = ev.do_eval(str(root))
    assert set(res) == {"test", "pose"} and abs(res["test"] - np.mean(got[:4])) < 1e-6
