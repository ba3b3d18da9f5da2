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


def make_flexible_ligand(n_heavy=24, n_tors=5, n_branch=3, seed=11):
    """Synthetic ligand with a torsion tree in the reference's representation (lib/tree.h): segment 0 is the rigid
    root, segments 1.. are torsion segments in DFS pre-order (= torsion order of `conf`).  Atoms are stored in the
    local frame of their segment (identity orientation): coords = segment origin + local.  A linear chain with
    `n_tors` rotatable bonds plus one side branch of `n_branch` atoms hanging off the root.
    -> dict(local_xyz, types, seg_parent, seg_begin, seg_end, seg_rel_origin, seg_rel_axis, pair_a, pair_b,
            conf0 [7+T] reproducing the generated coordinates, xyz0, gyration_radius, axis_root = per segment the atom at the
            begin of its rotation axis)"""
    rs = np.random.RandomState(seed)
    xyz = [np.zeros(3)]
    while len(xyz) < n_heavy:
        d = rs.randn(3); d /= np.linalg.norm(d)
        cand = xyz[-1] + 1.5 * d
        if all(np.linalg.norm(cand - q) > 1.3 for q in xyz):
            xyz.append(cand)
    breaks = sorted(rs.choice(np.arange(3, n_heavy - 2, 2), size=n_tors, replace=False).tolist())
    # side branch off atom 1 (root segment)
    branch = []
    cur = xyz[1]
    while len(branch) < n_branch:
        d = rs.randn(3); d /= np.linalg.norm(d)
        cand = cur + 1.5 * d
        if all(np.linalg.norm(cand - q) > 1.3 for q in xyz + branch):
            branch.append(cand); cur = cand
    pos = np.array(xyz + branch, np.float64)
    n = len(pos)
    types = rs.choice(LIG_HEAVY_TYPES, size=n).astype(np.int32)
    seg_begin = [0] + breaks + [n_heavy]
    seg_end = breaks + [n_heavy] + [n]
    seg_parent = [-1] + list(range(0, n_tors)) + [0]
    axis_root = [None] + [b - 1 for b in breaks] + [1]
    ns = len(seg_begin)
    origin = np.array([pos[seg_begin[s]] for s in range(ns)])
    rel_origin = np.zeros((ns, 3)); rel_axis = np.zeros((ns, 3))
    for s in range(1, ns):
        rel_origin[s] = origin[s] - origin[seg_parent[s]]
        a = origin[s] - pos[axis_root[s]]
        rel_axis[s] = a / np.linalg.norm(a)
    seg_of = np.zeros(n, int)
    for s in range(ns):
        seg_of[seg_begin[s]:seg_end[s]] = s
    local = pos - origin[seg_of]
    # chain neighbours for the 1-4 exclusion: consecutive chain atoms, branch hangs off atom 1
    nbr = {i: set() for i in range(n)}
    for i in range(n_heavy - 1):
        nbr[i].add(i + 1); nbr[i + 1].add(i)
    prev = 1
    for b in range(n_heavy, n):
        nbr[prev].add(b); nbr[b].add(prev); prev = b

    def within3(i):
        seen, front = {i}, {i}
        for _ in range(3):
            front = set(q for f in front for q in nbr[f]) - seen
            seen |= front
        return seen
    pa, pb = [], []
    for i in range(n):
        w = within3(i)
        for j2 in range(i + 1, n):
            if seg_of[i] != seg_of[j2] and j2 not in w:
                pa.append(i); pb.append(j2)
    conf0 = np.zeros(7 + ns - 1, np.float32)
    conf0[:3] = origin[0]; conf0[3] = 1.0
    gr = gyration_radius(pos.astype(np.float32), types, origin[0].astype(np.float32))
    return dict(local_xyz=local.astype(np.float32), types=types, seg_parent=np.array(seg_parent, np.int32),
                seg_begin=np.array(seg_begin, np.int32), seg_end=np.array(seg_end, np.int32),
                seg_rel_origin=rel_origin.astype(np.float32), seg_rel_axis=rel_axis.astype(np.float32),
                pair_a=np.array(pa, np.int32), pair_b=np.array(pb, np.int32), conf0=conf0, xyz0=pos.astype(np.float32),
                gyration_radius=gr, axis_root=np.array([0] + axis_root[1:], np.int32))


def make_tree_ligand(seed, max_children=3, n_seg_target=7):
    """A synthetic ligand with a RANDOM torsion tree (nodes with several children, nested branches) in the same representation as
    make_flexible_ligand: segments in DFS pre-order, every segment a short chain of 2-5 heavy atoms, children attached to a random atom
    of their parent (that atom is the begin point of the child's rotation axis).  Exercises the child-ordering rules of
    heterotree::set_conf / derivative (lib/tree.h:300-310,361-382) beyond a chain."""
    rs = np.random.RandomState(seed)
    pos, seg_of, bonds = [], [], []
    seg_parent, seg_begin, seg_end, axis_root = [], [], [], []
    budget = [n_seg_target - 1]

    def place(near):
        for _ in range(200):
            d = rs.randn(3); d /= np.linalg.norm(d)
            cand = near + 1.5 * d
            if all(np.linalg.norm(cand - q) > 1.25 for q in pos):
                return cand
        raise RuntimeError("could not place an atom")

    def build(parent, attach_atom):
        s = len(seg_parent)
        seg_parent.append(parent); seg_begin.append(len(pos)); axis_root.append(attach_atom if attach_atom is not None else 0)
        first = np.zeros(3) if attach_atom is None else place(pos[attach_atom])
        pos.append(first); seg_of.append(s)
        if attach_atom is not None:
            bonds.append((attach_atom, len(pos) - 1))
        for _ in range(rs.randint(1, 5)):
            pos.append(place(pos[-1])); seg_of.append(s); bonds.append((len(pos) - 2, len(pos) - 1))
        seg_end.append(len(pos))
        own = list(range(seg_begin[s], seg_end[s]))
        kids = min(budget[0], rs.randint(0, max_children + 1) if parent >= 0 else rs.randint(1, max_children + 1))
        budget[0] -= kids
        for _ in range(kids):
            build(s, int(rs.choice(own)))
    build(-1, None)
    pos = np.array(pos, np.float64)
    n, ns = len(pos), len(seg_parent)
    types = rs.choice(LIG_HEAVY_TYPES, size=n).astype(np.int32)
    origin = np.array([pos[seg_begin[s]] for s in range(ns)])
    rel_origin, rel_axis = np.zeros((ns, 3)), np.zeros((ns, 3))
    for s in range(1, ns):
        rel_origin[s] = origin[s] - origin[seg_parent[s]]
        a = origin[s] - pos[axis_root[s]]
        rel_axis[s] = a / np.linalg.norm(a)
    seg_of = np.array(seg_of)
    nbr = {i: set() for i in range(n)}
    for a, b in bonds:
        nbr[a].add(b); nbr[b].add(a)
    pa, pb = [], []
    for i in range(n):
        seen, front = {i}, {i}
        for _ in range(3):
            front = set(q for f in front for q in nbr[f]) - seen
            seen |= front
        for j in range(i + 1, n):
            if seg_of[i] != seg_of[j] and j not in seen:
                pa.append(i); pb.append(j)
    conf0 = np.zeros(7 + ns - 1, np.float32)
    conf0[:3] = origin[0]; conf0[3] = 1.0
    return dict(local_xyz=(pos - origin[seg_of]).astype(np.float32), types=types, seg_parent=np.array(seg_parent, np.int32),
                seg_begin=np.array(seg_begin, np.int32), seg_end=np.array(seg_end, np.int32),
                seg_rel_origin=rel_origin.astype(np.float32), seg_rel_axis=rel_axis.astype(np.float32),
                pair_a=np.array(pa, np.int32), pair_b=np.array(pb, np.int32), conf0=conf0, xyz0=pos.astype(np.float32),
                gyration_radius=gyration_radius(pos.astype(np.float32), types, origin[0].astype(np.float32)),
                axis_root=np.array(axis_root, np.int32))


def gyration_radius(xyz, types, origin):
    """model::gyration_radius (lib/model.cpp:1002-1014): root-mean-square distance of the HEAVY atoms (smina type >= 2) from the
    root origin, accumulated in float32 in atom order.  gb_ligand_topology.gyration_radius wants this number for the pose the
    search starts from (mutate_conf's first rotation uses it; afterwards the kernels follow the conformation the model holds)."""
    acc, cnt = np.float32(0), 0
    o = np.asarray(origin, np.float32)
    for p, t in zip(np.asarray(xyz, np.float32), types):
        if t >= 2:
            d = p - o
            acc = np.float32(acc + np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])))
            cnt += 1
    return float(np.sqrt(np.float32(acc / np.float32(cnt)))) if cnt else 0.0


def make_gninacheck_mol(rs, natoms=0, min_atoms=200, max_atoms=500, max_x=25.0, max_y=25.0, max_z=25.0):
    """The random-molecule generator of the reference's own `gninacheck` tests (test/gnina/test_utils.cpp:13-44,
    defaults test_utils.h:21-24): atom count uniform in [min_atoms, max_atoms + 1] unless given, coordinates uniform in
    [-max, max] per axis, smina type uniform over ALL 28 types (hydrogens and metals included), nothing prevents overlaps.
    `rs` is a numpy RandomState (MT19937 like the reference's std::mt19937; the draw algorithms of libstdc++'s
    distributions are not reproduced, so molecules match in distribution, not draw by draw).
    -> (xyz float32 [n,3], smina types int32 [n])"""
    if not natoms:
        natoms = int(rs.randint(min_atoms, max_atoms + 2))
    xyz = np.stack([rs.uniform(-m, m, size=natoms) for m in (max_x, max_y, max_z)], axis=1).astype(np.float32)
    types = rs.randint(0, 28, size=natoms).astype(np.int32)
    return xyz, types
