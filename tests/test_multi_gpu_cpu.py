"""N>1 host logic on CPU (gloo, world size 2): the bench's sharding is "same per-GPU batch on every rank, no
collective on the data path, max-over-ranks time, whole-job throughput = world * batch * steps / time"."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = textwrap.dedent('''
    import os, sys, json
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from bench import shard_frames, whole_job_rate
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard_frames(list(range(10)), rank, world)          # frame-level round robin (SURVEY 8e)
    allf = [None] * world
    dist.all_gather_object(allf, mine)
    t = torch.tensor([1.0 + rank])                              # rank 1 is slower
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"frames": allf, "t": float(t.item()), "rate": whole_job_rate(world, 256, 10, float(t.item()))}))
    dist.destroy_process_group()
''') % ROOT


def test_gloo_world2_sharding_and_timing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["frames"] == [[0, 2, 4, 6, 8], [1, 3, 5, 7, 9]]
    assert d["t"] == 2.0                                         # max over ranks
    assert d["rate"] == 2 * 256 * 10 / 2.0


WORKER2 = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tools"))
    from bench_pipeline import deal, pack_results
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    frames = deal(list(range(7)), rank, world)
    # each rank "detects" rank+1 faces in its frame; rows are collected on every rank (config 5's optional gather)
    res = [dict(box=np.full(4, 10 * rank + i, np.float32), kps=np.full((98, 2), i, np.float32),
                scores=np.full(98, rank, np.float32)) for i in range(rank + 1)]
    rows = torch.from_numpy(pack_results(res, 4))
    allrows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(allrows, rows)
    if rank == 0:
        print(json.dumps({"frames": frames, "n": [int((r.abs().sum(1) > 0).sum()) for r in allrows],
                          "box1": allrows[1][1, :4].tolist(), "shape": list(allrows[1].shape)}))
    dist.destroy_process_group()
''') % (ROOT, ROOT)


def test_gloo_world2_pipeline_result_gather(tmp_path):
    """tools/bench_pipeline.py's frame dealing and rank-0 collection of (box, landmarks, scores) rows."""
    script = tmp_path / "w2.py"
    script.write_text(WORKER2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29612", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["frames"] == [0, 2, 4, 6] and d["shape"] == [4, 298]
    assert d["n"] == [0, 2] or d["n"] == [1, 2]          # rank 0's single face has box 0 / kps 0 / score 0: all-zero row
    assert d["box1"] == [11.0, 11.0, 11.0, 11.0]


def test_bench_streams_frame_walk_is_cyclic_and_always_changes():
    """tools/bench_streams.make_streams (the pipeline leg of bench.py): every stream walks the rank's `length` frames with its
    own phase; consecutive frames of a stream always differ (so the frame-difference gate re-runs the detector on every
    frame), also across the wrap-around, and ranks get different jitters."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import bench_streams
    import frames

    def maker(jitter=(0, 0)):                                  # a small stand-in for frames.frame_1080p
        return frames.multi_face_frame(120, 160, (1, 1), 60, jitter)
    for seed in (0, 1000, 2000):
        seqs = bench_streams.make_streams(None, frames, maker, 5, length=4, seed0=seed, pin=False)
        assert len(seqs) == 5 and all(len(s) == 4 for s in seqs)
        for s in seqs:
            for t in range(4):
                assert (s[t] != s[(t + 1) % 4]).any()
        assert seqs[1][0] is seqs[0][1] and seqs[4][3] is seqs[0][(3 + 4) % 4]       # phase-shifted views of the same frames
    a = bench_streams.make_streams(None, frames, maker, 1, length=4, seed0=0, pin=False)
    b = bench_streams.make_streams(None, frames, maker, 1, length=4, seed0=1000, pin=False)
    assert any((x != y).any() for x, y in zip(a[0], b[0]))
