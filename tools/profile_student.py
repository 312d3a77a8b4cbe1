"""One forward of a network (for ncu captures): python tools/profile_student.py [batch] [reps] [student|teacher|detector]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import frames
from peppa_pig_face_landmark_b200 import ONNXEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model = sys.argv[3] if len(sys.argv) > 3 else "student"
if model == "teacher":
    from peppa_pig_face_landmark_b200 import teacher_graph
    path = teacher_graph.ensure_teacher_onnx()
elif model == "detector":
    path = os.path.join(ROOT, "peppa_pig_face_landmark_b200", "pretrained", "yolov5n-0.5.onnx")
else:
    path = os.path.join(ROOT, "peppa_pig_face_landmark_b200", "pretrained", "kps_student.onnx")
eng = ONNXEngine(path, max_batch=B)
x = frames.noise_crops(B, seed=1) if model != "detector" else \
    np.random.default_rng(1).integers(0, 256, (B, 384, 640, 3), dtype=np.uint8)
for _ in range(n):
    out = eng.run_u8(x)
print("ok", out[0].shape)
