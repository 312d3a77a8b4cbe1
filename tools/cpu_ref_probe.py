"""Why does the oracle port run ~1.5x faster inside the GPU bench process than in the `--impl reference` process?  Measures
CpuReference.rate(64) twice under one of the process set-ups named on the command line."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mode = sys.argv[1]
import torch
if mode == "cuda":
    torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
if mode == "lib":
    from peppa_pig_face_landmark_b200 import runtime as rt
    rt.load_library()
if mode == "flush":
    torch.set_flush_denormal(True)
if mode == "pinned":
    x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
import bench
ref = bench.CpuReference()
r = [ref.rate(64, seed=s)[0] for s in (1, 2)]
print("%-7s threads=%d  %s faces/s" % (mode, torch.get_num_threads(), ["%.1f" % v for v in r]))
