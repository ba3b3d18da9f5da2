"""Pose queue in front of the batch entry point: gnina's ligand loop (main/main.cpp:749-771, 233-269,
324-346) scores one pose per call; `PoseQueue.add` collects poses and every `capacity` poses (or on `flush`) ONE
`score_batch` call scores them; results come back in submission order.  Mirrors gb::PoseBatcher (gnina_b200.hpp)."""
import numpy as np


class PoseQueue:
    def __init__(self, scorer, capacity=1024, deliver=None, fixed_center=None):
        """scorer: anything with score_batch(xyz, types, offsets, centers) -> (score, affinity, loss, variance)."""
        self._scorer = scorer
        self.capacity = max(1, int(capacity))
        self._deliver = deliver
        self._center = None if fixed_center is None else np.asarray(fixed_center, np.float32).reshape(3)
        self._xyz, self._types, self._offs = [], [], [0]
        self._next = 0
        self.batches_run = 0
        self.results = {}            # ticket -> (score, affinity, loss, variance) when no deliver callback is given

    def __len__(self):
        return len(self._offs) - 1

    def add(self, xyz, types):
        xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
        types = np.asarray(types, np.int32).reshape(-1)
        if len(xyz) != len(types):
            raise ValueError("xyz and types disagree")
        self._xyz.append(xyz); self._types.append(types); self._offs.append(self._offs[-1] + len(types))
        ticket = self._next
        self._next += 1
        if len(self) >= self.capacity:
            self.flush()
        return ticket

    def flush(self):
        n = len(self)
        if n == 0:
            return
        xyz = np.concatenate(self._xyz) if self._offs[-1] else np.zeros((0, 3), np.float32)
        types = np.concatenate(self._types) if self._offs[-1] else np.zeros(0, np.int32)
        offs = np.asarray(self._offs, np.int32)
        # take the queue first: a failing batch must not be delivered twice
        self._xyz, self._types, self._offs = [], [], [0]
        first = self._next - n
        centers = None if self._center is None else np.tile(self._center, (n, 1))
        out = self._scorer.score_batch(xyz, types, offs, centers)
        self.batches_run += 1
        for i in range(n):
            r = tuple(float(o[i]) for o in out)
            if self._deliver:
                self._deliver(first + i, *r)
            else:
                self.results[first + i] = r

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.flush()
        return False
