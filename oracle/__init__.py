"""ORACLE — test infrastructure only.

CPU restatement of the reference's FaceAna inference path, used solely as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing in peppa_pig_face_landmark_b200/ may import it.

Parity pinning: the reference has no tests or golden vectors, so
this oracle is pinned against the reference's own Python host code executed
unchanged from /root/reference with `oracle/ort_shim` standing in for the absent
onnxruntime (tests/golden/make_golden.py; fixtures in tests/golden/*.npz).
"""
