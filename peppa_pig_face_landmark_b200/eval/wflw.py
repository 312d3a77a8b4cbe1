"""WFLW evaluation on the GPU — what /root/reference/TRAIN/face_landmark/tools/eval_WFLW.py:19-142 computes (NME per
test subset, inter-ocular normalisation) for any Skps landmark .onnx export, batched:

    python -m peppa_pig_face_landmark_b200.eval.wflw --data_dir <WFLW root> [--onnx kps_student.onnx] [--img_size 256]

data_dir holds WFLW_images/ and WFLW_annotations/list_98pt_test/*.txt exactly as the reference expects.  Per annotation
line: crop box from the ground-truth landmarks (augmentationCropImage, is_training=False, base_extend_range 0.2 from
train_config.py), zero-bordered crop + resize on the GPU (skps_crop_rect, bit-exact with the reference's OpenCV calls),
network forward in batches (ONNXEngine), NME on the GPU (skps_nme).  The reference normalises the labels by the crop size
and compares them with the network's normalised output; so does this."""
import argparse
import os

import numpy as np

from .. import runtime as rt

BASE_EXTEND = (0.2, 0.2)           # TRAIN/face_landmark/train_config.py: config.DATA.base_extend_range


def load_test_f(data_dir):
    """eval_WFLW.py:19-35: {subset name: [annotation lines]} from the list_98pt_test directory."""
    df = {}
    for txt in sorted(x for x in os.listdir(data_dir) if 'txt' in x):
        cls = txt.rsplit('.')[0].rsplit('_')[-1]
        with open(os.path.join(data_dir, txt)) as f:
            df[cls] = f.readlines()
    return df


def eval_crop_rect(bbox):
    """eval_WFLW.py:38-72 with is_training=False -> (x0, y0, w, h) of the crop in ORIGINAL image coordinates (the
    reference crops the zero-bordered image; the border offset `add` cancels out)."""
    bbox = np.array(bbox).reshape(4, ).astype(np.float32)
    add = int(max(bbox[2] - bbox[0], bbox[3] - bbox[1]))
    objcenter = np.array([(bbox[0] + bbox[2]) / 2., (bbox[1] + bbox[3]) / 2.])
    bbox = bbox + add
    objcenter = objcenter + add
    gt_width, gt_height = bbox[2] - bbox[0], bbox[3] - bbox[1]
    cwh = gt_width * (1 + BASE_EXTEND[0] * 2) // 2
    chh = gt_height * (1 + BASE_EXTEND[1] * 2) // 2
    min_x, max_x = int(objcenter[0] - cwh), int(objcenter[0] + cwh)
    min_y, max_y = int(objcenter[1] - chh), int(objcenter[1] + chh)
    return min_x - add, min_y - add, max_x - min_x, max_y - min_y


def nme(target, preds):
    """eval_WFLW.py:84-95 (host restatement, used by the tests)."""
    target = np.reshape(target, [-1, 98, 2])
    preds = np.reshape(preds, [-1, 98, 2])
    norm = np.linalg.norm(target[:, 60, :] - target[:, 72, :], axis=-1)
    return np.mean(np.mean(np.linalg.norm(preds - target, axis=-1), axis=-1) / norm)


class WFLWEvaluator:
    def __init__(self, onnx_path=None, input_size=256, batch=256):
        from ..core.api.onnx_model_base import ONNXEngine
        torch = rt.require_cuda()
        if onnx_path is None:
            onnx_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pretrained", "kps_student.onnx")
        self.engine = ONNXEngine(onnx_path, max_batch=batch)
        if self.engine.in_hw != (input_size, input_size):
            raise ValueError("%s takes %dx%d inputs; re-target it with graph_tools.retarget_input_size for --img_size %d"
                             % (onnx_path, *self.engine.in_hw, input_size))
        self.S, self.batch, self.lib, self.torch = input_size, batch, rt.load_library(), torch
        self.crops = torch.zeros((batch, input_size, input_size, 3), dtype=torch.uint8, device=self.engine.device)
        self.targets = torch.zeros((batch, 98, 2), dtype=torch.float32, device=self.engine.device)
        self.nme = torch.zeros((batch,), dtype=torch.float32, device=self.engine.device)

    def _flush(self, n, scores):
        torch, s = self.torch, self.engine.stream
        outs = self.engine.forward_device(self.crops[:n], stream=s)
        rt.check(self.lib.skps_nme(self.targets.data_ptr(), outs[0].data_ptr(), n, 98, 60, 72, self.nme.data_ptr(), s.cuda_stream))
        s.synchronize()
        scores.extend(self.nme[:n].cpu().numpy().tolist())

    def eval_lines(self, lines, image_dir, imread=None):
        """NME of each annotation line (196 landmark floats ... file name) -> list of float."""
        import cv2
        imread = imread or cv2.imread
        torch = self.torch
        scores, k = [], 0
        for dp in lines:
            dp = dp.split()
            kps = np.array(dp[:98 * 2], dtype=np.float32).reshape([-1, 2])
            image = imread(os.path.join(image_dir, dp[-1]))
            bbox = [float(np.min(kps[:, 0])), float(np.min(kps[:, 1])), float(np.max(kps[:, 0])), float(np.max(kps[:, 1]))]
            x0, y0, w, h = eval_crop_rect(bbox)
            frame = torch.from_numpy(np.ascontiguousarray(image)).to(self.engine.device)
            with torch.cuda.stream(self.engine.stream):
                rt.check(self.lib.skps_crop_rect(frame.data_ptr(), image.shape[0], image.shape[1], image.shape[1] * 3, x0, y0, w, h,
                                                 self.crops[k].data_ptr(), self.S, self.engine.stream.cuda_stream))
                label = (kps - np.array([x0, y0], np.float32)) / np.array([w, h], np.float32)     # eval_WFLW.py:74-75, 117-118
                self.targets[k].copy_(torch.from_numpy(label.astype(np.float32)), non_blocking=False)
            self.engine.stream.synchronize()          # `frame` may be freed after this point
            k += 1
            if k == self.batch:
                self._flush(k, scores)
                k = 0
        if k:
            self._flush(k, scores)
        return scores

    def do_eval(self, data_dir):
        """eval_WFLW.py:97-142: {subset: mean NME}."""
        df = load_test_f(os.path.join(data_dir, 'WFLW_annotations/list_98pt_test'))
        out = {}
        for name, lines in df.items():
            out[name] = float(np.mean(self.eval_lines(lines, os.path.join(data_dir, 'WFLW_images'))))
            print('for cls:', name, ' nme:', out[name])
        return out


def main():
    ap = argparse.ArgumentParser(description='WFLW NME on the GPU.')
    ap.add_argument('--data_dir', required=True)
    ap.add_argument('--onnx', default=None, help='Skps landmark export (default: the shipped student)')
    ap.add_argument('--img_size', type=int, default=256)
    a = ap.parse_args()
    WFLWEvaluator(a.onnx, a.img_size).do_eval(a.data_dir)


if __name__ == '__main__':
    main()
