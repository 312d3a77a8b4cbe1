#!/usr/bin/env python
"""BASELINE config 4: Teacher@256 (HRNet-w18 + Decoder, synthetic weights) landmark-only, batch sweep
1..1024 on one B200.  Prints one JSON line per batch size: faces/s with crops resident in HBM (CUDA events)
and the achieved fraction of 2 * 5.757 GMAC per face against the measured bf16 peak (and peak/3, the
ceiling of the 3-MMA fp16 hi/lo scheme).  Usage: python tools/bench_teacher.py [--batches 1,2,...] [--steps K]"""
import argparse
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from peppa_pig_face_landmark_b200 import ONNXEngine, teacher_graph as T
    path = T.ensure_teacher_onnx()
    peak = 1590.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("bf16_tflops", peak)
    lines = []
    for B in [int(b) for b in args.batches.split(",")]:
        free, _ = torch.cuda.mem_get_info()
        eng = None
        try:
            eng = ONNXEngine(path, max_batch=B)
        except RuntimeError as e:
            print(json.dumps({"batch": B, "error": str(e)[:200]}))
            continue
        sets = [torch.from_numpy(T.synthetic_crops(min(B, 8), 256, 50 + i)).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().cuda()
                for i in range(2)]
        outs = [torch.empty((B, e), dtype=torch.float32, device="cuda") for e in eng.out_elems]
        st = eng.stream
        steps = args.steps if B >= 16 else args.steps * 4
        with torch.cuda.stream(st):
            for i in range(args.warmup):
                eng.forward_device(sets[i & 1], outs, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                eng.forward_device(sets[i & 1], outs, st)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        fps = B / ms * 1e3
        tf = fps * 2 * eng.macs_per_sample / 1e12
        line = {"workload": "teacher256_landmark_only", "batch": B, "faces_per_s": fps, "ms_per_step": ms,
                "gmac_per_face": eng.macs_per_sample / 1e9, "tflops_2mac": tf, "frac_of_bf16_peak": tf / peak,
                "frac_of_split_ceiling": tf / (peak / 3), "launches": eng.launches, "weights": "synthetic seed 0"}
        print(json.dumps(line))
        sys.stdout.flush()
        lines.append(line)
        del eng, sets, outs
        torch.cuda.empty_cache()
    if args.out:
        json.dump(lines, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
