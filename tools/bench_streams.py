"""Full-pipeline throughput with FaceAnaStreams (BASELINE configs 3 and 5): S concurrent synthetic video streams per GPU,
one frame per stream per call, two calls in flight (frame uploads overlap compute).  Faces jitter every frame so the
frame-difference gate re-runs the detector on every frame (the worst case for the pipeline).  Host (pinned) frames in,
host results out: every H2D / D2H is inside the timed region.

    python tools/bench_streams.py [--streams 16] [--batches 12] [--configs 1080p_4faces,4k_16faces] [--gather]

Under torchrun every rank drives its own S streams on its own GPU (streams shard across GPUs, no collective on the data
path); time = max over ranks.  --gather adds one NCCL all_gather of the packed (box, landmarks, scores) rows per call."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

CONFIGS = {"1080p_4faces": ("frame_1080p", 4), "4k_16faces": ("frame_4k", 16)}


def make_streams(torch, frames, maker, n_streams, length=6, seed0=0, pin=True):
    """`length` distinct frames per rank (faces jittered by multiples of 4 px), every stream walks them with its own phase:
    consecutive frames of a stream differ, so the frame-difference gate re-runs the detector, and the set-up cost does not
    grow with the number of streams (a 4K frame takes ~0.3 s to synthesise on a host core)."""
    rng = np.random.default_rng(seed0)
    jit = []
    while len(jit) < length:                      # no two consecutive frames alike (the walk is cyclic)
        j = (int(rng.integers(-2, 3)) * 4, int(rng.integers(-2, 3)) * 4)
        if not jit or (j != jit[-1] and (len(jit) < length - 1 or j != jit[0])):
            jit.append(j)
    base = [maker(jitter=j) for j in jit]
    if pin:
        base = [torch.from_numpy(f).pin_memory().numpy() for f in base]       # what a capture / decoder ring hands over
    return [[base[(t + s) % length] for t in range(length)] for s in range(n_streams)]


def run_config(name, n_streams=16, batches=12, warmup=3, gather=False, dist=None, rank=0, world=1, length=6):
    """Returns a dict with whole-job frames/s and faces/s (all ranks), or None on ranks != 0."""
    import torch
    import frames
    from Skps import FaceAnaStreams
    maker, topk = getattr(frames, CONFIGS[name][0]), CONFIGS[name][1]
    seqs = make_streams(torch, frames, maker, n_streams, length=length, seed0=1000 * rank)
    H, W = seqs[0][0].shape[:2]
    fa = FaceAnaStreams(n_streams=n_streams, top_k=topk, max_frame_hw=(H, W))
    L = len(seqs[0])
    rows = torch.zeros((n_streams * topk, 298), device="cuda") if (gather and dist is not None) else None
    allrows = [torch.zeros_like(rows) for _ in range(world)] if rows is not None else None

    def batch(t):
        return [seqs[s][t % L] for s in range(n_streams)]

    def consume(res):
        n = sum(len(r) for r in res)
        if rows is not None:
            host = np.zeros((n_streams * topk, 298), np.float32)
            for s, faces in enumerate(res):
                for i, r in enumerate(faces):
                    host[s * topk + i, :4] = r["box"]
                    host[s * topk + i, 4:200] = np.asarray(r["kps"], np.float32).reshape(-1)
                    host[s * topk + i, 200:] = r["scores"]
            rows.copy_(torch.from_numpy(host))
            dist.all_gather(allrows, rows)
        return n

    for t in range(warmup):
        fa.run(batch(t))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    faces = 0
    fa.submit(batch(0))
    for t in range(1, batches):
        fa.submit(batch(t))
        faces += consume(fa.collect())
    faces += consume(fa.collect())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    det_share = float(np.mean(fa.last_ran_detector))
    if dist is not None:
        t = torch.tensor([dt, float(faces)], device="cuda", dtype=torch.float64)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, faces = float(tmax[0]), float(tsum[1])
    del fa
    if rank != 0:
        return None
    frames_total = world * n_streams * batches
    return {"config": name, "n_gpus": world, "streams_per_gpu": n_streams, "calls": batches,
            "frames_per_s": frames_total / dt, "faces_per_s": faces / dt, "faces_per_frame": faces / frames_total,
            "ms_per_call": 1e3 * dt / batches, "h2d_bytes_per_frame": int(H * W * 3),
            "d2h_bytes_per_frame": int(topk * (32 + 98 * 2 * 8 + 98 * 4)), "detector_runs_per_frame": det_share,
            "api": "FaceAnaStreams.submit/collect, pinned host frames, 2 calls in flight, temporal layer on the device",
            "gather_to_rank0": bool(rows is not None)}


def main():
    import torch
    a = sys.argv[1:]

    def opt(name, default):
        return a[a.index(name) + 1] if name in a else default
    n_streams, batches = int(opt("--streams", 16)), int(opt("--batches", 12))
    names = opt("--configs", ",".join(CONFIGS)).split(",")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    for name in names:
        r = run_config(name, n_streams, batches, gather="--gather" in a, dist=dist, rank=rank, world=world)
        if r is not None:
            print(json.dumps(r))
            sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
