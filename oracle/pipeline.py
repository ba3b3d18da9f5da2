"""End-to-end CPU oracle of one CNN scoring call (TEST INFRASTRUCTURE ONLY).

Mirrors TorchModel::forward (gninasrc/lib/torch_model.cpp:153-224) + CNNTorchScorer::score
(gninasrc/lib/cnn_torch_scorer.cpp:105-198) with rotations off (default --cnn_rotation 0).
"""
import numpy as np
import torch
from . import gridmaker as gm
from . import cnn_ref


class OracleModel:
    def __init__(self, blob):
        self.blob = blob
        self.n_rec, self.rec_t2c = gm.parse_typemap(blob.recmap)
        self.n_lig, self.lig_t2c = gm.parse_typemap(blob.ligmap)
        self.n_channels = self.n_rec + self.n_lig

    def grids(self, rec_xyz, rec_types, lig_xyz, lig_types, pose_offsets, centers=None, n_threads=1):
        b = self.blob
        rc, rr = gm.type_atoms(rec_types, self.rec_t2c, 0)
        lc, lr = gm.type_atoms(lig_types, self.lig_t2c, self.n_rec)
        return gm.grid_forward_batch(rec_xyz, rc, rr, lig_xyz, lc, lr, pose_offsets, self.n_channels, centers,
                                     b.resolution, b.dimension, b.radius_scaling, n_threads)

    def score(self, rec_xyz, rec_types, lig_xyz, lig_types, pose_offsets, centers=None, dtype=torch.float32,
              batch=8, n_threads=1):
        """-> (pose[n], affinity[n], loss[n]) float arrays for this single model."""
        n = len(pose_offsets) - 1
        P, A, L = [], [], []
        for s in range(0, n, batch):
            e = min(n, s + batch)
            off = np.asarray(pose_offsets[s:e + 1])
            g = self.grids(rec_xyz, rec_types, lig_xyz[off[0]:off[-1]], lig_types[off[0]:off[-1]], off - off[0],
                           None if centers is None else centers[s:e], n_threads)
            p, a, l = cnn_ref.score_grid(self.blob, g, dtype)
            P.append(p); A.append(a); L.append(l)
        return np.concatenate(P), np.concatenate(A), np.concatenate(L)


def score_ensemble(models, *args, **kw):
    """-> (score[n], affinity[n], loss[n], variance[n]) with the reference's ensemble arithmetic."""
    per = [m.score(*args, **kw) for m in models]
    n = len(per[0][0])
    out = np.zeros((4, n))
    for i in range(n):
        out[:, i] = cnn_ref.ensemble([p[0][i] for p in per], [p[1][i] for p in per], [p[2][i] for p in per])
    return out[0], out[1], out[2], out[3]
