"""Run ONE plan op under the CUDA profiler range (for `ncu --profile-from-start off --set full ...`):

    python tools/profile_op.py <name-substring | #index | dominant> [batch] [reps] [student|teacher|detector]

One full forward runs first (outside the range) so the op's inputs hold real activations; then the selected op is
launched `reps` times between cudaProfilerStart/Stop.  `dominant` = the dense conv with the most MACs (what
bench.py's roofline times).  Prints the op index/name so the capture can be tied to the plan."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frames  # noqa: E402
from peppa_pig_face_landmark_b200 import ONNXEngine, plan as P, runtime as rt  # noqa: E402

sel = sys.argv[1] if len(sys.argv) > 1 else "dominant"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = sys.argv[4] if len(sys.argv) > 4 else "student"
pre = os.path.join(ROOT, "peppa_pig_face_landmark_b200", "pretrained")
if model == "teacher":
    from peppa_pig_face_landmark_b200 import teacher_graph
    path = teacher_graph.ensure_teacher_onnx()
elif model == "detector":
    path = os.path.join(pre, "yolov5n-0.5.onnx")
else:
    path = os.path.join(pre, "kps_student.onnx")
eng = ONNXEngine(path, max_batch=B)
ops = eng.plan.ops


def macs(op):
    o = op.outs[0]
    return o.C * o.H * o.W * op.ins[0].C * op.k[0] * op.k[1] if op.type == P.OP_CONV else 0


if sel == "dominant":
    idxs = [max(range(len(ops)), key=lambda i: macs(ops[i]))]
elif sel.startswith("#"):
    idxs = [int(v) for v in sel[1:].split(",")]
else:
    idxs = [i for i, op in enumerate(ops) if sel in op.name]
assert idxs, "no op matches %r" % sel
ih, iw = eng.in_hw
x = frames.noise_crops(B, seed=1) if (ih, iw) == (256, 256) else \
    np.random.default_rng(1).integers(0, 256, (B, ih, iw, 3), dtype=np.uint8)
eng.run_u8(x)
lib = rt.load_library()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in idxs:
    for _ in range(reps):
        rt.check(lib.skps_engine_run_op(eng.handle, i, B, eng.stream.cuda_stream))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
for i in idxs:
    op = ops[i]
    print("profiled op #%d %s %s cin=%d cout=%d %dx%d k=%d batch=%d" % (
        i, P.OP_NAMES[op.type], op.name, op.ins[0].C, op.outs[0].C, op.outs[0].H, op.outs[0].W, op.k[0], B))
