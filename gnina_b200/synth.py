"""Deterministic synthetic receptor + ligand-pose generator (SURVEY.md §8d / BASELINE.md §4).

No OpenBabel: atoms are carried as `.gninatypes`-style records (float x,y,z; int32 smina type;
gninasrc/gninatyper/gninatyper.cpp:30-36).  numpy's MT19937 RandomState is the seeded source.
"""
import numpy as np

# receptor-side smina-type mix, (smina type id, probability)
REC_TYPE_MIX = [(2, .22), (3, .22), (4, .06), (5, .03), (6, .03), (7, .12), (9, .02), (13, .20), (12, .06),
                (10, .01), (14, .02), (23, .01)]
LIG_HEAVY_TYPES = [2, 3, 4, 5, 6, 7, 9, 10, 12, 13, 14, 17, 18]


def make_receptor(n_atoms=3000, box=60.0, min_sep=1.2, seed=20240229):
    rs = np.random.RandomState(seed)
    pts = np.empty((0, 3), np.float32)
    cell = {}
    out = []
    inv = 1.0 / min_sep
    while len(out) < n_atoms:
        p = (rs.rand(3) - 0.5) * box
        key = tuple(np.floor(p * inv).astype(int))
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for q in cell.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                        if np.sum((q - p) ** 2) < min_sep * min_sep:
                            ok = False
        if ok:
            cell.setdefault(key, []).append(p)
            out.append(p)
    xyz = np.asarray(out, np.float32)
    ids = np.array([t for t, _ in REC_TYPE_MIX])
    pr = np.array([w for _, w in REC_TYPE_MIX]); pr = pr / pr.sum()
    types = ids[rs.choice(len(ids), size=n_atoms, p=pr)].astype(np.int32)
    return xyz, types


def make_ligand(n_heavy=30, n_polar_h=4, seed=7):
    """Random-walk chain with 1.5 Å bonds (+ polar H of smina type 1 at 1.0 Å from random heavy atoms)."""
    rs = np.random.RandomState(seed)
    xyz = [np.zeros(3)]
    while len(xyz) < n_heavy:
        d = rs.randn(3); d /= np.linalg.norm(d)
        cand = xyz[-1] + 1.5 * d
        if all(np.linalg.norm(cand - q) > 1.3 for q in xyz):
            xyz.append(cand)
    types = list(rs.choice(LIG_HEAVY_TYPES, size=n_heavy, p=None))
    for _ in range(n_polar_h):
        host = rs.randint(n_heavy)
        d = rs.randn(3); d /= np.linalg.norm(d)
        xyz.append(xyz[host] + 1.0 * d)
        types.append(1)
    xyz = np.asarray(xyz, np.float32)
    xyz -= xyz.mean(0, keepdims=True)
    return xyz, np.asarray(types, np.int32)


def random_rotations(rs, n):
    q = rs.randn(n, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(n, 3, 3)
    return R


def make_poses(lig_xyz, n_poses, trans_box=16.0, seed=1):
    """n rigid poses: uniform random rotation + translation uniform in a cube at the receptor centre.
    -> (xyz [n_poses*n_atoms,3] float32, pose_offsets int32[n_poses+1])"""
    rs = np.random.RandomState(seed)
    R = random_rotations(rs, n_poses)
    t = (rs.rand(n_poses, 3) - 0.5) * trans_box
    xyz = np.einsum("pij,aj->pai", R, lig_xyz.astype(np.float64)) + t[:, None, :]
    na = len(lig_xyz)
    return xyz.reshape(-1, 3).astype(np.float32), (np.arange(n_poses + 1) * na).astype(np.int32)


def make_screen(n_ligands, seed=3, trans_box=16.0, nmin=15, nmax=45):
    """Config-4 style: many different ligands (N_heavy ~ U[nmin,nmax]) x 1 pose each (ragged)."""
    rs = np.random.RandomState(seed)
    xs, ts, offs = [], [], [0]
    for i in range(n_ligands):
        nh = int(rs.randint(nmin, nmax + 1))
        lx, lt = make_ligand(nh, int(rs.randint(0, 5)), seed=int(rs.randint(1 << 30)))
        px, _ = make_poses(lx, 1, trans_box, seed=int(rs.randint(1 << 30)))
        xs.append(px); ts.append(lt); offs.append(offs[-1] + len(lt))
    return np.concatenate(xs), np.concatenate(ts), np.asarray(offs, np.int32)
