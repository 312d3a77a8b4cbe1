"""CPU checks of the host logic: the ONNX->plan lowering (executed by the torch plan interpreter)
against the oracle graph executor, plan serialisation, and the C-ABI export list."""
import os
import re

import numpy as np
import pytest

import frames

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PRE = os.path.join(ROOT, "peppa_pig_face_landmark_b200", "pretrained")


def test_student_plan_matches_oracle_graph():
    from peppa_pig_face_landmark_b200 import lowering
    from oracle.plan_interp import PlanInterp
    from oracle.onnx_exec import Session
    plan = lowering.lower(os.path.join(PRE, "kps_student.onnx"), (256, 256))
    assert plan.macs == 1482829696            # 1.4828e9 MAC per face
    crops = frames.crop_variants(2)
    xy, sc = PlanInterp(plan).run(crops)
    sess = Session(os.path.join(PRE, "kps_student.onnx"))
    for i in range(2):
        x = crops[i].transpose(2, 0, 1).astype(np.float32)[None] / np.float32(255)
        o, s = sess.run(x)
        assert np.abs(xy[i] - o.reshape(-1)).max() * 256 < 1e-3
        assert np.abs(sc[i] - s.reshape(-1)).max() < 1e-4


def test_detector_plan_matches_oracle_graph():
    from peppa_pig_face_landmark_b200 import lowering, plan as P
    from oracle.plan_interp import PlanInterp
    from oracle.onnx_exec import Session
    from oracle import host_ref as H
    plan = lowering.lower(os.path.join(PRE, "yolov5n-0.5.onnx"), (384, 640))
    assert plan.macs == 441169920
    assert sum(o.type == P.OP_COPY for o in plan.ops) == 13      # only the shuffle pass-through halves move
    x, _ = H.letterbox(frames.load_test1())
    u8 = np.round(x[0].transpose(1, 2, 0) * 255).astype(np.uint8)[None]
    out = PlanInterp(plan).run(u8)[0][0]
    ref = Session(os.path.join(PRE, "yolov5n-0.5.onnx")).run(x)[0].reshape(15120, 16)
    assert np.array_equal(np.where(out[:, 4] > 0.5)[0], np.where(ref[:, 4] > 0.5)[0])
    assert np.abs(out - ref).max() < 5e-3


def test_plan_serialisation_layout():
    from peppa_pig_face_landmark_b200 import lowering, plan as P
    plan = lowering.lower(os.path.join(PRE, "kps_student.onnx"), (256, 256))
    words, blob = plan.serialize()
    assert words[0] == 0x534B5053 and words[2] == len(plan.bufs) and words[3] == len(plan.ops)
    body = 8 + 4 * len(plan.bufs) + P.OP_WORDS * len(plan.ops)
    assert words.size == body + 1 + 3 * words[body]          # trailer: L2 chunking segments
    segs = words[body + 1:].reshape(-1, 3)
    assert segs[0, 0] == 0 and segs[-1, 1] == len(plan.ops) and (segs[1:, 0] == segs[:-1, 1]).all()
    assert blob.dtype == np.float32 and all(op.w_off % 4 == 0 for op in plan.ops if op.w_off >= 0)


def test_c_abi_exports_every_declared_symbol():
    """include/skps_b200.h vs the built library vs the ctypes table (no compute calls)."""
    from peppa_pig_face_landmark_b200 import build, runtime
    build.build()
    hdr = open(os.path.join(ROOT, "include", "skps_b200.h")).read()
    declared = set(re.findall(r"SKPS_API [\w\s\*]+?(skps_\w+)\(", hdr))
    assert len(declared) >= 25
    assert declared == set(runtime.SIGNATURES), declared ^ set(runtime.SIGNATURES)
    lib = runtime.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.skps_version() == 1


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from peppa_pig_face_landmark_b200 import FaceAna
    with pytest.raises(RuntimeError):
        FaceAna()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "peppa_pig_face_landmark_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
