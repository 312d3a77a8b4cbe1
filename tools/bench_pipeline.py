"""Full-pipeline throughput (BASELINE configs 3 and 5): FaceAna.run on synthetic 1080p (4 faces) and 4K
(16 faces) streams whose faces jitter every frame so the detector runs on every frame.  Host frames in,
host results out (H2D of the frame and D2H of the results inside the timed region).  Prints one JSON line
per config, plus the oracle CPU path on a few frames."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import frames


def stream(maker, n):
    rng = np.random.default_rng(0)
    return [maker(jitter=(int(rng.integers(-2, 3)) * 4, int(rng.integers(-2, 3)) * 4)) for _ in range(n)]


def main():
    import torch
    from Skps import FaceAna
    from oracle.faceana_ref import FaceAnaRef
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    for name, maker, topk in (("1080p_4faces", frames.frame_1080p, 4), ("4k_16faces", frames.frame_4k, 16)):
        fr = stream(maker, 8)
        if os.environ.get("PIN_FRAMES", "1") != "0":
            # frame buffers in page-locked memory (what a capture/decoder ring would hand over): no staging copy
            fr = [torch.from_numpy(f).pin_memory().numpy() for f in fr]
        facer = FaceAna(top_k=topk)
        for f in fr[:3]:
            res = facer.run(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nf = 0
        for i in range(n_frames):
            res = facer.run(fr[i % len(fr)])
            nf += len(res)
        dt = time.perf_counter() - t0
        ref = FaceAnaRef(top_k=topk)
        ref.run(fr[0])
        t1 = time.perf_counter()
        nr = 0
        for f in fr[1:4]:
            nr += len(ref.run(f))
        dtr = time.perf_counter() - t1
        print(json.dumps({"config": name, "frames_per_s": n_frames / dt, "faces_per_s": nf / dt,
                          "faces_per_frame": nf / n_frames, "ms_per_frame": 1e3 * dt / n_frames,
                          "cpu_oracle_frames_per_s": 3 / dtr, "cpu_oracle_faces_per_frame": nr / 3,
                          "h2d_bytes_per_frame": int(fr[0].nbytes),
                          "frames_pinned": os.environ.get("PIN_FRAMES", "1") != "0"}))


if __name__ == "__main__":
    main()
