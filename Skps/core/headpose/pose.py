"""Drop-in import path of the reference: `from Skps.core.headpose.pose import get_head_pose` (Skps/core/headpose/pose.py)."""
from peppa_pig_face_landmark_b200.core.headpose.pose import (get_head_pose, head_poses, line_pairs, object_pts,  # noqa: F401
                                                              reprojectsrc)
