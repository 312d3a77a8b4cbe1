"""Head pose on the GPU (csrc/headpose.cu) against the reference's own OpenCV calls (Skps/core/headpose/pose.py:48-77):
cv2.solvePnP + cv2.projectPoints + cv2.Rodrigues + cv2.decomposeProjectionMatrix on the same points.  Tolerances: the two
solvers stop at slightly different points of the same minimum (OpenCV: 20 LM iterations in its own parametrisation), so
Euler angles agree to 1e-3 degrees and re-projected cube corners to 1e-2 px."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_pose(shape, img_hw):
    """pose.py:48-77 verbatim in behaviour (the reference function, restated around the same cv2 calls)."""
    import cv2
    from peppa_pig_face_landmark_b200.core.headpose.pose import object_pts, reprojectsrc
    h, w = img_hw
    K = [w, 0.0, w // 2, 0.0, w, h // 2, 0.0, 0.0, 1.0]
    cam = np.array(K).reshape(3, 3).astype(np.float32)
    dist = np.zeros((5, 1), np.float32)
    image_pts = np.float32([shape[17], shape[21], shape[22], shape[26], shape[36], shape[39], shape[42], shape[45],
                            shape[31], shape[35]])
    _, rvec, tvec = cv2.solvePnP(object_pts, image_pts, cam, dist)
    dst, _ = cv2.projectPoints(reprojectsrc, rvec, tvec, cam, dist)
    rot, _ = cv2.Rodrigues(rvec)
    euler = cv2.decomposeProjectionMatrix(cv2.hconcat((rot, tvec)))[6]
    return dst.reshape(8, 2), euler.reshape(3), rvec.reshape(3), tvec.reshape(3)


def _synthetic_shapes(n, img_hw, seed=0):
    import cv2
    from peppa_pig_face_landmark_b200.core.headpose.pose import object_pts, POSE_POINTS
    rng = np.random.default_rng(seed)
    h, w = img_hw
    shapes = rng.uniform(0, w, (n, 68, 2)).astype(np.float32)           # the other 58 points are not read
    for i in range(n):
        rv = np.array([math.pi, 0, 0]) + rng.uniform(-0.5, 0.5, 3)       # a face looking at the camera, +- 30 degrees
        R = cv2.Rodrigues(rv)[0]
        X = object_pts.astype(np.float64) @ R.T + np.array([rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(50, 140)])
        uv = np.stack([w * X[:, 0] / X[:, 2] + w // 2, w * X[:, 1] / X[:, 2] + h // 2], 1) + rng.normal(0, 0.8, (10, 2))
        shapes[i, POSE_POINTS] = uv
    return shapes


@pytest.mark.parametrize("img_hw", [(480, 640), (1080, 1920)])
def test_head_pose_matches_opencv(img_hw):
    from peppa_pig_face_landmark_b200.core.headpose.pose import head_poses
    shapes = _synthetic_shapes(64, img_hw, seed=img_hw[0])
    got = head_poses(shapes, img_hw)
    for i in range(len(shapes)):
        dst, euler, rvec, tvec = _reference_pose(shapes[i], img_hw)
        d = np.abs(got["euler"][i] - euler)
        assert np.minimum(d, 360 - d).max() < 1e-3, (i, got["euler"][i], euler)
        assert np.abs(got["reproject"][i] - dst).max() < 1e-2, i
        import cv2
        # the rotation itself (cv2 may return the same rotation as a vector with |r| > pi; compare the matrices)
        assert np.abs(cv2.Rodrigues(got["rvec"][i])[0] - cv2.Rodrigues(rvec)[0]).max() < 1e-5, i
        assert np.abs(got["tvec"][i] - tvec).max() < 1e-2, i


def test_get_head_pose_signature_matches_reference():
    from Skps.core.headpose.pose import get_head_pose
    img = np.zeros((480, 640, 3), np.uint8)
    shape = _synthetic_shapes(1, (480, 640), seed=3)[0]
    reproject, euler = get_head_pose(shape, img)
    assert isinstance(reproject, tuple) and len(reproject) == 8 and len(reproject[0]) == 2 and euler.shape == (3, 1)
    dst, e_ref, _, _ = _reference_pose(shape, (480, 640))
    assert np.abs(np.array(reproject) - dst).max() < 1e-2 and np.abs(euler.reshape(3) - e_ref).max() < 1e-3
