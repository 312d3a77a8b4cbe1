"""FaceAnaStreams (csrc/mpipe.cu + csrc/temporal.cu): many video streams per GPU with the temporal layer on the device.
Each stream must return what its own FaceAna instance returns (facer.py:52-85 incl. GroupTrack / One-Euro / track-box
EMA, lk.py:6-162): checked against the golden video fixture made from the unmodified reference, and against the
single-stream FaceAna (host-side temporal layer) on other sequences, frame sizes and stream/slot interleavings."""
import numpy as np
import pytest

import frames
from golden.make_golden_frames import video_frames
from test_parity_gpu import KPS_TOL_PX, SCORE_TOL, _check_result

pytestmark = pytest.mark.gpu


def _same(a, b, tol=1e-6):
    assert len(a) == len(b), (len(a), len(b))
    for x, y in zip(a, b):
        assert np.abs(np.asarray(x["kps"], np.float64) - np.asarray(y["kps"], np.float64)).max() <= tol
        assert np.abs(np.asarray(x["box"], np.float64) - np.asarray(y["box"], np.float64)).max() <= tol
        assert np.array_equal(x["scores"], y["scores"])


def _sequences():
    v = video_frames()
    c = frames.canvas_640()
    c2 = c.copy()
    c2[::9, ::4] = np.clip(c2[::9, ::4].astype(np.int16) + 2, 0, 255).astype(np.uint8)       # under the diff threshold
    t1 = frames.load_test1()
    return [v,                                          # the golden clip: detect, static x2, moved, empty x2
            [v[3], v[0], v[1], v[4], v[0], v[2]],       # other order: re-detections, a face-less frame in between
            [c, c2, c, c2, c2, c],                      # 640x640: tracker path + One-Euro on tiny motion
            [t1, t1, v[4], t1, t1, t1]]                 # frame size changes mid-stream (no previous frame of that size)


def test_streams_match_golden_video_and_single_stream_faceana(golden):
    from Skps import FaceAna, FaceAnaStreams
    seqs = _sequences()
    fa = FaceAnaStreams(n_streams=len(seqs))
    singles = [FaceAna() for _ in seqs]
    g = golden("video1080")
    for t in range(6):
        res = fa.run([s[t] for s in seqs])
        _check_result(res[0], g, t, "video1080 via FaceAnaStreams")
        for k, s in enumerate(seqs):
            _same(res[k], singles[k].run(s[t]))


def test_streams_two_batches_in_flight_equal_blocking_runs():
    from Skps import FaceAnaStreams
    seqs = _sequences()[:3]
    a, b = FaceAnaStreams(n_streams=3), FaceAnaStreams(n_streams=3)
    want = [a.run([s[t] for s in seqs]) for t in range(6)]
    got = []
    b.submit([s[0] for s in seqs])
    for t in range(1, 6):
        b.submit([s[t] for s in seqs])
        got.append(b.collect())
    got.append(b.collect())
    for w, g_ in zip(want, got):
        for x, y in zip(w, g_):
            _same(x, y, tol=0.0)


def test_streams_reset_one_stream_and_partial_batches():
    from Skps import FaceAna, FaceAnaStreams
    v = video_frames()
    fa = FaceAnaStreams(n_streams=4)
    one = FaceAna()
    r0 = fa.run([v[0], v[0]])                      # only streams 0 and 1 are fed
    ref0 = one.run(v[0])
    _same(r0[0], ref0); _same(r0[1], ref0)
    fa.reset(1)                                    # stream 1 forgets its previous frame: detector path again
    r1 = fa.run([v[1], v[1]])
    assert list(fa.last_ran_detector) == [False, True]
    _same(r1[0], one.run(v[1]))
    fresh = FaceAna()
    _same(r1[1], fresh.run(v[1]))
    with pytest.raises(ValueError):
        fa.run([v[0]] * 5)
