"""Summarise an ncu launch list (gpu__time_duration.sum, one forward) per plan op:
python tools/launch_table.py launches.csv [top] [student|teacher|detector]"""
import csv, sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from collections import defaultdict
from peppa_pig_face_landmark_b200 import lowering, plan as P
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
model = sys.argv[3] if len(sys.argv) > 3 else "student"
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
if model == 'teacher':
    from peppa_pig_face_landmark_b200 import teacher_graph
    pl = lowering.lower(teacher_graph.ensure_teacher_onnx(), (256, 256))
elif model == 'detector':
    pl = lowering.lower(os.path.join(ROOT, 'peppa_pig_face_landmark_b200/pretrained/yolov5n-0.5.onnx'), (384, 640))
else:
    pl = lowering.lower(os.path.join(ROOT, 'peppa_pig_face_landmark_b200/pretrained/kps_student.onnx'), (256, 256))
ops = []
for op in pl.ops:
    ops.append(op)
    if op.type == P.OP_UPCAT_DW and len(rows) != len(pl.ops):
        ops.append(op)          # TMA variant = up-sampled part + skip part, two launches
assert len(rows) == len(ops), (len(rows), len(ops))
out, tot, d = [], 0.0, defaultdict(float)
for i, (row, op) in enumerate(zip(rows, ops)):
    t = float(row['Metric Value']) / 1e3
    tot += t
    o = op.outs[0]
    kind = P.OP_NAMES[op.type] + ('/XF' if op.flags & P.FLAG_XF else '/TC' if op.flags & 2 else '')
    d[kind] += t
    cin = op.w.shape[1] if op.type == P.OP_DWPW else op.ins[0].C
    out.append((t, i, kind, cin, o.C, o.H, op.k[0], op.name[-44:]))
print('total %.1f us over %d launches' % (tot, len(rows)))
for x in sorted(out, reverse=True)[:top]:
    print('%9.1f us  #%-3d %-22s cin=%-4d cout=%-4d H=%-4d k=%d %s' % x)
print({k: round(v, 1) for k, v in sorted(d.items(), key=lambda kv: -kv[1])})
