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


def score_grad(models, rec_xyz, rec_types, lig_xyz, lig_types, pose_offsets, centers=None, dtype=torch.float64,
               receptor=False):
    """CNNTorchScorer::score(m, compute_gradient=True): ensemble outputs + d(mean loss)/d(ligand atoms) [n_atoms,3]
    (TorchModel::forward autograd + GridMaker::backward, torch_model.cpp:197-221; 1/cnt scaling
    cnn_torch_scorer.cpp:176-179)."""
    n = len(pose_offsets) - 1
    # the reference accumulates the models' atom gradients in the model's float minus_forces (model::add_minus_forces) and scales the
    # sum once by fl(1.0 / cnt) (scale_minus_forces, cnn_torch_scorer.cpp:176-179): the same association here
    grad = np.zeros((len(lig_types), 3), np.float32)
    rgrad = np.zeros((len(rec_types), 3), np.float64)   # getReceptorGradient (torch_model.cpp:226-232), summed over poses
    per = []
    for m in models:
        b = m.blob
        P, A, L = [], [], []
        rc, rr = gm.type_atoms(rec_types, m.rec_t2c, 0)
        lc, lr = gm.type_atoms(lig_types, m.lig_t2c, m.n_rec)
        for p in range(n):
            sl = slice(pose_offsets[p], pose_offsets[p + 1])
            c = gm.center_of(lig_xyz[sl]) if centers is None else np.asarray(centers[p], np.float32)
            xyz = np.concatenate([rec_xyz, lig_xyz[sl]])
            ch = np.concatenate([rc, lc[sl]])
            rad = np.concatenate([rr, lr[sl]])
            g = gm.grid_forward(c, xyz, ch, rad, m.n_channels, b.resolution, b.dimension, b.radius_scaling)
            pose, aff, loss, dg = cnn_ref.loss_grid_gradient(b, g[None], dtype)
            ag = gm.grid_backward(c, xyz, ch, rad, dg[0].astype(np.float32), b.resolution, b.dimension, b.radius_scaling)
            grad[sl] += ag[len(rec_xyz):].astype(np.float32)
            rgrad += ag[:len(rec_xyz)] / len(models)
            P.append(pose[0]); A.append(aff[0]); L.append(loss[0])
        per.append((P, A, L))
    if len(models) > 1:
        grad *= np.float32(1.0 / len(models))
    out = np.zeros((4, n))
    for i in range(n):
        out[:, i] = cnn_ref.ensemble([q[0][i] for q in per], [q[1][i] for q in per], [q[2][i] for q in per])
    if receptor:
        return out[0], out[1], out[2], out[3], grad, rgrad
    return out[0], out[1], out[2], out[3], grad


def score_ensemble(models, *args, **kw):
    """-> (score[n], affinity[n], loss[n], variance[n]) with the reference's ensemble arithmetic."""
    per = [m.score(*args, **kw) for m in models]
    n = len(per[0][0])
    out = np.zeros((4, n))
    for i in range(n):
        out[:, i] = cnn_ref.ensemble([p[0][i] for p in per], [p[1][i] for p in per], [p[2][i] for p in per])
    return out[0], out[1], out[2], out[3]
