"""Host-side rows SURVEY.md 8(f) ranks next: the pose queue in front of the batch entry point (8f-1) and the
typed-atom `.gninatypes` reader (8f-2).  CPU: logic with a stub scorer; GPU: through the library."""
import os
import subprocess
import numpy as np
import pytest
from gnina_b200 import gninatypes
from gnina_b200.batching import PoseQueue


def test_gninatypes_round_trip_and_errors(tmp_path):
    rs = np.random.RandomState(0)
    xyz = rs.randn(17, 3).astype(np.float32) * 10
    t = rs.randint(0, 28, 17).astype(np.int32)
    for name in ("lig.gninatypes", "lig.gninatypes.gz"):
        p = tmp_path / name
        gninatypes.write_gninatypes(p, xyz, t)
        x2, t2 = gninatypes.read_gninatypes(p)
        assert x2.dtype == np.float32 and t2.dtype == np.int32
        assert np.array_equal(x2, xyz) and np.array_equal(t2, t)
    raw = (tmp_path / "lig.gninatypes").read_bytes()
    assert len(raw) == 16 * 17                        # gninatyper.cpp:30-36: 16-byte records, no header
    assert np.frombuffer(raw[:12], "<f4").tolist() == xyz[0].tolist() and int(np.frombuffer(raw[12:16], "<i4")[0]) == t[0]
    (tmp_path / "bad.gninatypes").write_bytes(raw[:-3])
    with pytest.raises(ValueError, match="Truncated"):
        gninatypes.read_gninatypes(tmp_path / "bad.gninatypes")
    (tmp_path / "empty.gninatypes").write_bytes(b"")
    x0, t0 = gninatypes.read_gninatypes(tmp_path / "empty.gninatypes")
    assert x0.shape == (0, 3) and t0.shape == (0,)
    xs, ts, offs = gninatypes.read_many([tmp_path / "lig.gninatypes", tmp_path / "empty.gninatypes", tmp_path / "lig.gninatypes.gz"])
    assert offs.tolist() == [0, 17, 17, 34] and len(ts) == 34


class StubScorer:
    def __init__(self):
        self.calls = []

    def score_batch(self, xyz, types, offs, centers=None):
        n = len(offs) - 1
        self.calls.append(n)
        na = np.diff(offs).astype(np.float32)
        first_x = np.array([xyz[offs[i], 0] if offs[i + 1] > offs[i] else -1 for i in range(n)], np.float32)
        c = np.full(n, -1, np.float32) if centers is None else np.asarray(centers, np.float32)[:, 0]
        return na, first_x, c, np.full(n, n, np.float32)


def test_pose_queue_orders_batches_and_flushes():
    s = StubScorer()
    got = []
    q = PoseQueue(s, capacity=3, deliver=lambda t, *r: got.append((t, r)))
    for p in range(7):
        assert q.add(np.full((p + 1, 3), 10 + p, np.float32), np.full(p + 1, 2, np.int32)) == p
    assert s.calls == [3, 3] and len(q) == 1 and len(got) == 6
    q.flush(); q.flush()
    assert s.calls == [3, 3, 1] and q.batches_run == 3
    assert [t for t, _ in got] == list(range(7))
    for t, r in got:
        assert r[0] == t + 1 and r[1] == 10 + t and r[2] == -1 and r[3] == (3 if t < 6 else 1)
    # fixed centre (cnn_center), results dictionary, context manager flush, empty pose
    s2 = StubScorer()
    with PoseQueue(s2, capacity=100, fixed_center=[4, 5, 6]) as q2:
        q2.add(np.zeros((0, 3), np.float32), np.zeros(0, np.int32))
        q2.add(np.ones((2, 3), np.float32), [2, 3])
    assert s2.calls == [2] and q2.results[0][2] == 4.0 and q2.results[1][0] == 2.0 and q2.results[0][0] == 0.0
    with pytest.raises(ValueError):
        q2.add(np.ones((2, 3), np.float32), [2])


def test_failed_batch_is_not_delivered_twice():
    class Boom(StubScorer):
        def score_batch(self, *a, **k):
            if not self.calls:
                self.calls.append(-1)
                raise RuntimeError("device error")
            return super().score_batch(*a, **k)
    s = Boom()
    q = PoseQueue(s, capacity=2)
    q.add(np.ones((1, 3)), [2])
    with pytest.raises(RuntimeError):
        q.add(np.ones((1, 3)), [2])
    assert len(q) == 0
    q.add(np.ones((1, 3)), [2]); q.flush()
    assert list(q.results) == [2]


def test_cpp_pose_batcher_and_gninatypes(tmp_path):
    from test_cpp_host import build_exe
    out = subprocess.check_output([build_exe(), "--host", str(tmp_path / "t.gninatypes")], text=True).strip().splitlines()
    assert out[0] == "gninatypes ok"
    assert out[1] == "queued 1 batches 2 delivered 6"
    for p in range(7):
        assert out[2 + p] == "ticket %d n %d x %d c -1 batch %d" % (p, p + 1, 10 + p, 3 if p < 6 else 1)
    assert out[9] == "fixed centre 4 deliveries 1"
    assert out[10] == "merge 2 1"       # chain 1's better copy (flat index 2) replaces chain 0's; the distinct pose stays


@pytest.mark.gpu
def test_queue_and_gninatypes_through_the_library(golden_dir, tmp_path):
    from gnina_b200 import CNNScorer
    kat = np.load(os.path.join(golden_dir, "cnn_kat.npz"))
    offs = kat["pose_offsets"]
    n = len(offs) - 1
    s = CNNScorer(["crossdock_default2018"])
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    direct = s.score_batch(kat["lig_xyz"], kat["lig_types"], offs)
    paths = []
    for p in range(n):
        paths.append(tmp_path / ("pose%d.gninatypes" % p))
        gninatypes.write_gninatypes(paths[-1], kat["lig_xyz"][offs[p]:offs[p + 1]], kat["lig_types"][offs[p]:offs[p + 1]])
    q = PoseQueue(s, capacity=4)
    for p in paths:
        q.add(*gninatypes.read_gninatypes(p))
    q.flush()
    assert q.batches_run == 2
    for p in range(n):
        assert abs(q.results[p][0] - direct[0][p]) < 1e-6 and abs(q.results[p][1] - direct[1][p]) < 1e-5
    x, t, o = gninatypes.read_many(paths)
    again = s.score_batch(x, t, o)
    assert np.array_equal(again[0], direct[0])
