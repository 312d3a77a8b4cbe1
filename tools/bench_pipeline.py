"""Full-pipeline throughput (BASELINE configs 3 and 5): FaceAna.run on synthetic 1080p (4 faces) and 4K
(16 faces) streams whose faces jitter every frame so the detector runs on every frame.  Host frames in,
host results out (H2D of the frame and D2H of the results inside the timed region).  Prints one JSON line
per config, plus the oracle CPU path on a few frames.

Multi-GPU (config 5): launch under torchrun; frames are dealt round-robin to the ranks (one
FaceAna per GPU, no collective on the data path), the timed region is bracketed by barriers and the time is
the max over ranks.  With --gather every rank's (box, landmarks, scores) rows of the last frame are collected
on rank 0 with one NCCL all_gather (the optional collection step north_star names); it is outside the data path
but inside the timed region.

    python tools/bench_pipeline.py [n_frames] [--gather] [--configs 4k_16faces]
"""
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


def deal(items, rank, world):
    """Frame-level round-robin sharding (same rule as bench.shard_frames)."""
    return list(items[rank::world])


def pack_results(res, topk):
    """[{box,kps,scores}] -> float32 (topk, 4+196+98) rows, zero padded, for the rank-0 collection."""
    out = np.zeros((topk, 4 + 196 + 98), np.float32)
    for i, r in enumerate(res[:topk]):
        out[i, :4] = np.asarray(r["box"], np.float32)[:4]
        out[i, 4:200] = np.asarray(r["kps"], np.float32).reshape(-1)
        out[i, 200:] = np.asarray(r["scores"], np.float32)
    return out


def main():
    import torch
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_frames = int(args[0]) if args else 40
    gather = "--gather" in sys.argv
    only = None
    if "--configs" in sys.argv:
        only = sys.argv[sys.argv.index("--configs") + 1].split(",")
        args = [a for a in args if a not in only]
        n_frames = int(args[0]) if args else 40
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from Skps import FaceAna

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for name, maker, topk in (("1080p_4faces", frames.frame_1080p, 4), ("4k_16faces", frames.frame_4k, 16)):
        if only and name not in only:
            continue
        fr = deal(stream(maker, 8 * world), rank, world)        # this rank's share of the stream
        if os.environ.get("PIN_FRAMES", "1") != "0":
            # frame buffers in page-locked memory (what a capture/decoder ring would hand over): no staging copy
            fr = [torch.from_numpy(f).pin_memory().numpy() for f in fr]
        facer = FaceAna(top_k=topk)
        for f in fr[:3]:
            res = facer.run(f)
        rows = torch.zeros((topk, 298), device="cuda")
        allrows = [torch.zeros_like(rows) for _ in range(world)] if (gather and dist is not None) else None
        barrier()
        t0 = time.perf_counter()
        nf = 0
        for i in range(n_frames):
            res = facer.run(fr[i % len(fr)])
            nf += len(res)
            if allrows is not None:
                rows.copy_(torch.from_numpy(pack_results(res, topk)))
                dist.all_gather(allrows, rows)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt, float(nf)], device="cuda", dtype=torch.float64)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dt, nf = float(tmax[0]), float(tsum[1])
        if rank != 0:
            continue
        line = {"config": name, "n_gpus": world, "frames_per_s": world * n_frames / dt, "faces_per_s": nf / dt,
                "faces_per_frame": nf / (world * n_frames), "ms_per_frame_per_gpu": 1e3 * dt / n_frames,
                "h2d_bytes_per_frame": int(fr[0].nbytes), "frames_pinned": os.environ.get("PIN_FRAMES", "1") != "0",
                "sharding": "frame round-robin, one FaceAna per GPU, no data-path collective",
                "gather_to_rank0": bool(allrows is not None)}
        if world == 1 and "--no-cpu" not in sys.argv:
            from oracle.faceana_ref import FaceAnaRef
            ref = FaceAnaRef(top_k=topk)
            ref.run(fr[0])
            t1 = time.perf_counter()
            nr = 0
            for f in fr[1:4]:
                nr += len(ref.run(f))
            dtr = time.perf_counter() - t1
            line.update({"cpu_oracle_frames_per_s": 3 / dtr, "cpu_oracle_faces_per_frame": nr / 3})
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
