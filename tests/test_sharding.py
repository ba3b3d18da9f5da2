"""World-size-2 gloo test of the pose-sharding + final gather logic (the N>1 path of bench.py / SURVEY.md §8e)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from gnina_b200 import sharding, synth


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 10, 10000):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _fake_score(xyz, types, offs, centers):
    # deterministic per-pose function of the pose's atoms (stands in for the CUDA scorer on the CPU box)
    n = len(offs) - 1
    s = np.array([xyz[offs[i]:offs[i + 1]].sum() for i in range(n)], np.float32)
    t = np.array([types[offs[i]:offs[i + 1]].sum() for i in range(n)], np.float32)
    return s, t


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lx, lt, offs = synth.make_screen(11, seed=9)
    full = sharding.score_sharded(_fake_score, lx, lt, offs)
    q.put((rank, full[0].tolist(), full[1].tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    lx, lt, offs = synth.make_screen(11, seed=9)
    want = _fake_score(lx, lt, offs, None)
    for rank, a, b in res:
        assert np.allclose(a, want[0]) and np.allclose(b, want[1])
