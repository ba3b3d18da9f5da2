"""Typed-atom records: `.gninatypes` = flat array of {float32 x, y, z; int32 smina_type}
(gninasrc/gninatyper/gninatyper.cpp:30-36, one molecule per file, :72-77; optionally gzipped).  The smina type is the
int32 the C ABI takes, so pre-typed ligands go straight to `CNNScorer.score_batch` without OpenBabel."""
import gzip
import numpy as np

RECORD = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("type", "<i4")])


def read_gninatypes(path):
    """-> (xyz float32 [n,3], types int32 [n]); raises ValueError on a truncated file."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) % RECORD.itemsize:
        raise ValueError("Truncated gninatypes file %s" % path)
    rec = np.frombuffer(raw, RECORD)
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float32) if len(rec) else np.zeros((0, 3), np.float32)
    return xyz, rec["type"].astype(np.int32)


def write_gninatypes(path, xyz, types):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    types = np.asarray(types, np.int32).reshape(-1)
    if len(xyz) != len(types):
        raise ValueError("xyz and types disagree")
    rec = np.empty(len(types), RECORD)
    rec["x"], rec["y"], rec["z"], rec["type"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], types
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(rec.tobytes())


def read_many(paths):
    """Concatenate one-molecule files into the (xyz, types, pose_offsets) triple of the batch entry point."""
    xs, ts, offs = [], [], [0]
    for p in paths:
        x, t = read_gninatypes(p)
        xs.append(x); ts.append(t); offs.append(offs[-1] + len(t))
    xyz = np.concatenate(xs) if xs else np.zeros((0, 3), np.float32)
    types = np.concatenate(ts) if ts else np.zeros(0, np.int32)
    return xyz, types, np.asarray(offs, np.int32)
