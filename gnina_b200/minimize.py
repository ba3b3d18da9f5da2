"""Lock-step minimisation of MANY poses (BASELINE config 5: `--minimize --cnn_scoring all`, 1 k poses).

The reference minimises one pose at a time: quasi_newton (lib/quasi_newton.cpp:49-83 -> bfgs.h:358-502) calls
non_cache_cnn::eval_deriv once per function evaluation, i.e. one CNN forward + backward for ONE pose per call
(main/main.cpp:264-268).  A B200 wants thousands of poses per call, so here the same BFGS runs for all poses at once as a
vectorised state machine: every round each unfinished pose has exactly one pending function evaluation (the first evaluation, or a
line-search trial), the pending conformations of ALL poses go to the energy function in one batch (-> one gb_cnn_score_grad call),
and every pose then advances by its own rules -- accept / backtrack / Hessian update / stop -- exactly as bfgs.h prescribes
(fast_line_search :73-91, accurate_line_search :107-180 with its float / double mix, bfgs_update :52-66, --minimize_early_term
:455-462, the restore of x_orig :494-498).  Per pose the sequence of evaluations is the reference's.

Host side only (numpy, float32 in the reference's operation order): the torsion-tree kinematics (tree.h set_conf / derivative,
quaternion.h) and the quasi-Newton bookkeeping; the energy function is the device's.  Checked on the CPU against the reference's own
quasi_newton + non_cache_cnn (oracle/_ref) with a test double for the network: tests/test_oracle_vs_reference_build.py."""
import numpy as np

F = np.float32
PI = F(3.1415926535897931)
EPS = F(1.1920929e-07)


# sin / cos / acos of float32 arguments: correctly rounded (evaluated in double, rounded once) like the device kernels.  Tests that
# compare with the reference compiled on this host swap in the host's sinf / cosf / acosf (set_transcendentals), which is what that build
# executes, to demand identical trajectories.
_fn = {"sin": lambda a: np.sin(a.astype(np.float64)).astype(F), "cos": lambda a: np.cos(a.astype(np.float64)).astype(F),
       "acos": lambda a: np.arccos(a.astype(np.float64)).astype(F)}


def set_transcendentals(sin=None, cos=None, acos=None):
    """replace (or, with None, restore) the float32 -> float32 sin / cos / acos used by the kinematics; test hook"""
    _fn["sin"] = sin or (lambda a: np.sin(a.astype(np.float64)).astype(F))
    _fn["cos"] = cos or (lambda a: np.cos(a.astype(np.float64)).astype(F))
    _fn["acos"] = acos or (lambda a: np.arccos(a.astype(np.float64)).astype(F))


def _normalize_angle(x):
    """normalize_angle (lib/common.h): into [-pi, pi]"""
    x = x.astype(F).copy()
    for _ in range(2):
        big, small = x > 3 * PI, x < -3 * PI
        if big.any():
            nn = (x[big] - PI) / (2 * PI)
            x[big] = x[big] - 2 * PI * np.ceil(nn)
        if small.any():
            nn = (-x[small] - PI) / (2 * PI)
            x[small] = x[small] + 2 * PI * np.ceil(nn)
    x = np.where(x > PI, x - 2 * PI, x)
    x = np.where(x < -PI, x + 2 * PI, x)
    return x.astype(F)


def _sincos_half(angle):
    """sin / cos of angle / 2"""
    h = (angle / F(2)).astype(F)
    return _fn["sin"](h), _fn["cos"](h)


def _angle_to_q(axis, angle):
    angle = _normalize_angle(angle)
    s, c = _sincos_half(angle)
    return np.stack([c, s * axis[:, 0], s * axis[:, 1], s * axis[:, 2]], axis=1).astype(F)


def _qmul(l, r):
    a, b, c, d = l[:, 0], l[:, 1], l[:, 2], l[:, 3]
    return np.stack([a * r[:, 0] - b * r[:, 1] - c * r[:, 2] - d * r[:, 3],
                     a * r[:, 1] + b * r[:, 0] + c * r[:, 3] - d * r[:, 2],
                     a * r[:, 2] - b * r[:, 3] + c * r[:, 0] + d * r[:, 1],
                     a * r[:, 3] + b * r[:, 2] - c * r[:, 1] + d * r[:, 0]], axis=1).astype(F)


def _qnorm_approx(q):
    """quaternion_normalize_approx (quaternion.h:243-257): leave it alone when |q|^2 is within 1e-6 of 1"""
    s = q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3]
    fix = ~(np.abs(s - F(1)) < F(1e-6))
    if fix.any():
        q = q.copy()
        q[fix] = q[fix] * (F(1) / np.sqrt(s[fix]))[:, None]
    return q.astype(F)


def _q_to_r3(q):
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    aa, ab, ac, ad, bb, bc, bd, cc, cd, dd = a * a, a * b, a * c, a * d, b * b, b * c, b * d, c * c, c * d, d * d
    m = np.empty((len(q), 3, 3), F)
    m[:, 0, 0] = aa + bb - cc - dd; m[:, 0, 1] = 2 * (-ad + bc); m[:, 0, 2] = 2 * (ac + bd)
    m[:, 1, 0] = 2 * (ad + bc); m[:, 1, 1] = aa - bb + cc - dd; m[:, 1, 2] = 2 * (-ab + cd)
    m[:, 2, 0] = 2 * (-ac + bd); m[:, 2, 1] = 2 * (ab + cd); m[:, 2, 2] = aa - bb - cc + dd
    return m


def _mv(m, v):
    """m [n,3,3] times v [n,3] or [3], summed left to right"""
    v = np.broadcast_to(v, (len(m), 3))
    return (m[:, :, 0] * v[:, None, 0] + m[:, :, 1] * v[:, None, 1] + m[:, :, 2] * v[:, None, 2]).astype(F)


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]],
                    axis=1).astype(F)


def _quaternion_increment(q, rot):
    """quaternion_increment (quaternion.cu:32-43,96-100): q <- normalize_approx(angle_to_quaternion(rot) * q)"""
    angle = np.sqrt(rot[:, 0] * rot[:, 0] + rot[:, 1] * rot[:, 1] + rot[:, 2] * rot[:, 2]).astype(F)
    r = np.zeros((len(q), 4), F); r[:, 0] = 1
    big = angle > EPS
    if big.any():
        axis = (F(1) / angle[big])[:, None] * rot[big]
        r[big] = _angle_to_q(axis.astype(F), angle[big])
    return _qnorm_approx(_qmul(r, q))


def _dot(a, b):
    """scalar_product: sequential float32 sum over the degrees of freedom, for every pose"""
    s = np.zeros(len(a), F)
    for i in range(a.shape[1]):
        s = s + a[:, i] * b[:, i]
    return s


class TorsionTree:
    """heterotree<rigid_body> kinematics for a batch of conformations (lib/tree.h:218-233,300-310,361-382).  `lig` = the ligand
    description gb_ligand_topology takes (local_xyz, seg_parent / seg_begin / seg_end in DFS pre-order, seg_rel_origin, seg_rel_axis)."""

    def __init__(self, lig):
        self.local = np.ascontiguousarray(lig["local_xyz"], F)
        self.parent = [int(p) for p in lig["seg_parent"]]
        self.begin = [int(b) for b in lig["seg_begin"]]
        self.end = [int(e) for e in lig["seg_end"]]
        self.rel_o = np.ascontiguousarray(lig["seg_rel_origin"], F)
        self.rel_a = np.ascontiguousarray(lig["seg_rel_axis"], F)
        self.ns, self.na = len(self.parent), len(self.local)
        self.T = self.ns - 1
        self.children = [[c for c in range(self.ns) if self.parent[c] == s] for s in range(self.ns)]

    def set_conf(self, X):
        """model::set: X [n, 7+T] -> (coords [n,na,3], segment origins [n,ns,3], segment axes [n,ns,3])"""
        X = np.ascontiguousarray(X, F)
        n = len(X)
        so, sa = np.zeros((n, self.ns, 3), F), np.zeros((n, self.ns, 3), F)
        q, M = [None] * self.ns, [None] * self.ns
        coords = np.empty((n, self.na, 3), F)
        for s in range(self.ns):
            if s == 0:
                so[:, 0] = X[:, :3]
                q[0] = X[:, 3:7].copy()
            else:
                p = self.parent[s]
                so[:, s] = so[:, p] + _mv(M[p], self.rel_o[s])
                sa[:, s] = _mv(M[p], self.rel_a[s])
                q[s] = _qnorm_approx(_qmul(_angle_to_q(sa[:, s], X[:, 7 + s - 1]), q[p]))
            M[s] = _q_to_r3(q[s])
            for i in range(self.begin[s], self.end[s]):
                coords[:, i] = so[:, s] + _mv(M[s], self.local[i])
        return coords, so, sa

    def derivative(self, coords, forces, so, sa):
        """heterotree::derivative: minus_forces [n,na,3] -> change [n, 6+T] (force, torque about the root origin, torsion derivatives)"""
        n = len(coords)
        ft = np.zeros((n, self.ns, 6), F)
        change = np.zeros((n, 6 + self.T), F)
        for s in range(self.ns - 1, -1, -1):
            f, t = np.zeros((n, 3), F), np.zeros((n, 3), F)
            for i in range(self.begin[s], self.end[s]):                     # sum_force_and_torque, atoms in order
                f = f + forces[:, i]
                t = t + _cross(coords[:, i] - so[:, s], forces[:, i])
            for c in self.children[s]:                                       # branches_derivative, children ascending
                f = f + ft[:, c, :3]
                t = t + (_cross(so[:, c] - so[:, s], ft[:, c, :3]) + ft[:, c, 3:])
            ft[:, s, :3], ft[:, s, 3:] = f, t
            if s == 0:
                change[:, :3], change[:, 3:6] = f, t
            else:
                change[:, 6 + s - 1] = t[:, 0] * sa[:, s, 0] + t[:, 1] * sa[:, s, 1] + t[:, 2] * sa[:, s, 2]
        return change


def conf_increment(X, p, alpha, T):
    """conf::increment (lib/conf.h:54-59,113-118,385-393): position, orientation (rotation vector), normalised torsions"""
    X = X.copy()
    a = alpha[:, None].astype(F)
    X[:, :3] = X[:, :3] + a * p[:, :3]
    X[:, 3:7] = _quaternion_increment(X[:, 3:7], (a * p[:, 3:6]).astype(F))
    for t in range(T):
        X[:, 7 + t] = _normalize_angle(X[:, 7 + t] + _normalize_angle(alpha * p[:, 6 + t]))
    return X.astype(F)


def _conf_as_change_coordinates(X, T):
    """conf::operator()(i) (lib/conf.h:459-473): position, quaternion_to_angle(orientation) (quaternion.cu:46-62), torsions"""
    n = len(X)
    out = np.zeros((n, 6 + T), F)
    out[:, :3] = X[:, :3]
    c = X[:, 3]
    inside = (c > -1) & (c < 1)
    if inside.any():
        angle = (2 * _fn["acos"](c[inside].astype(F))).astype(F)
        angle = np.where(angle > PI, angle - 2 * PI, angle).astype(F)
        s = _fn["sin"]((angle / F(2)).astype(F))
        ok = ~(np.abs(s) < EPS)
        f = np.zeros_like(angle); f[ok] = angle[ok] / s[ok]
        out[inside, 3:6] = X[inside, 4:7] * f[:, None]
    out[:, 6:] = X[:, 7:]
    return out


def minimize_poses(tree, energy_and_forces, X0, maxiters=10000, accurate=True, early_term=False):
    """quasi_newton for every row of X0 [n, 7+T] in lock step.
    energy_and_forces(coords [k,na,3], idx [k]) -> (e [k], minus_forces [k,na,3]) is called once per round with the pending
    conformations of all unfinished poses; idx = which rows of X0 they belong to.  accurate = BFGSAccurateLineSearch (what --minimize selects, main/main.cpp:1160), maxiters 10000 is gnina's
    default for it (:1157-1158).  -> (e [n], X [n, 7+T], function evaluations per pose [n], rounds)"""
    X0 = np.ascontiguousarray(X0, F)
    n, T = len(X0), tree.T
    m = 6 + T

    def f(X, idx):
        coords, so, sa = tree.set_conf(X)
        e, mf = energy_and_forces(coords, idx)
        return np.asarray(e, F), tree.derivative(coords, np.asarray(mf, F), so, sa)

    x = X0.copy()
    f0, g = f(x, np.arange(n))
    evals = np.ones(n, np.int64)
    f_orig, x_orig, g_orig = f0.copy(), x.copy(), g.copy()
    H = np.zeros((n, m, m), F)
    H[:, np.arange(m), np.arange(m)] = 1
    p = np.zeros((n, m), F)
    pg, alpha, alpha2, f2, alamin = (np.zeros(n, F) for _ in range(5))
    trial, step = np.zeros(n, np.int64), np.zeros(n, np.int64)
    active = np.ones(n, bool) if maxiters > 0 else np.zeros(n, bool)

    def start_iteration(mask):
        """p = -H g (minus_mat_vec_product), slope, first trial step"""
        idx = np.flatnonzero(mask)
        if not len(idx):
            return
        s = np.zeros((len(idx), m), F)
        for j in range(m):
            s = s + H[idx, :, j] * g[idx, j][:, None]
        p[idx] = -s
        pg[idx] = _dot(p[idx], g[idx])
        alpha[idx], trial[idx], alpha2[idx], f2[idx] = 1, 0, 0, 0
        if accurate:
            stop = pg[idx] >= 0                                              # not a descent direction: the search returns 0, bfgs gives up
            active[idx[stop]] = False
            xc = _conf_as_change_coordinates(x[idx], T)
            ax = np.abs(xc)
            test = (np.abs(p[idx]) / np.where(ax < 1, F(1), ax)).max(axis=1)  # compute_lambdamin
            with np.errstate(divide="ignore"):
                alamin[idx] = EPS / test

    start_iteration(active)
    rounds = 0
    while active.any():
        rounds += 1
        A = np.flatnonzero(active)
        x_new = conf_increment(x[A], p[A], alpha[A], T)
        f1, g_new = f(x_new, A)
        evals[A] += 1
        al, sl, f0A = alpha[A], pg[A], f0[A]
        over = np.zeros(len(A), bool)
        if accurate:
            too_small = (al < alamin[A]) | ~np.isfinite(al)
            enough = ~too_small & (f1 <= f0A + F(1.0e-4) * al * sl)
            over = too_small | enough
            al = np.where(too_small, F(0), al)
            back = ~over
            if back.any():
                b_ = np.flatnonzero(back)
                a_, s_, f1_, f0_, a2_, f2_ = (v[b_].astype(F) for v in (al, sl, f1, f0A, alpha2[A], f2[A]))
                first = a_ == 1
                with np.errstate(all="ignore"):
                    t_first = (-s_.astype(np.float64) / (2.0 * (f1_ - f0_ - s_).astype(np.float64))).astype(F)
                    rhs1 = (f1_ - f0_ - a_ * s_).astype(F)
                    rhs2 = (f2_ - f0_ - a2_ * s_).astype(F)
                    ca = ((rhs1 / (a_ * a_) - rhs2 / (a2_ * a2_)) / (a_ - a2_)).astype(F)
                    cb = ((-a2_ * rhs1 / (a_ * a_) + a_ * rhs2 / (a2_ * a2_)) / (a_ - a2_)).astype(F)
                    disc = ((cb * cb).astype(np.float64) - 3.0 * ca.astype(np.float64) * s_.astype(np.float64)).astype(F)
                    sq = np.sqrt(np.where(disc < 0, F(0), disc)).astype(F)
                    t_lin = (-s_.astype(np.float64) / (2.0 * cb.astype(np.float64))).astype(F)
                    t_neg = ((-cb + sq).astype(np.float64) / (3.0 * ca.astype(np.float64))).astype(F)
                    t_pos = (-s_ / (cb + sq)).astype(F)
                    t_cub = np.where(disc < 0, (0.5 * a_.astype(np.float64)).astype(F), np.where(cb <= 0, t_neg, t_pos))
                    t_later = np.where(ca == 0, t_lin, t_cub)
                    half = (0.5 * a_.astype(np.float64))
                    t_later = np.where(t_later.astype(np.float64) > half, half.astype(F), t_later).astype(F)
                    tmplam = np.where(first, t_first, t_later).astype(F)
                tenth = (F(0.1) * a_).astype(F)
                new_alpha = np.where(tmplam < tenth, tenth, tmplam).astype(F)   # std::max(tmplam, 0.1 alpha)
                alpha2[A[b_]], f2[A[b_]] = a_, f1_
                al = al.copy(); al[b_] = new_alpha
        else:
            accepted = f1 - f0A < F(0.0001) * al * sl
            al = np.where(accepted, al, al * F(0.5)).astype(F)
            trial[A] += ~accepted
            over = accepted | (trial[A] >= 10)
        alpha[A] = al
        O = np.flatnonzero(over)
        if not len(O):
            continue
        gO, lO = A[O], np.arange(len(A))[O]
        gave_up = alpha[gO] == 0                                                # the line search found nothing: bfgs breaks
        active[gO[gave_up]] = False
        keep = ~gave_up
        gK, lK = gO[keep], lO[keep]
        if not len(gK):
            continue
        y = (g_new[lK] - g[gK]).astype(F)
        prev = f0[gK].copy()
        f0[gK] = f1[lK]
        x[gK] = x_new[lK]
        early = np.zeros(len(gK), bool)
        if early_term:
            early = np.abs((prev - f0[gK]).astype(np.float64)) < 1e-5            # before g is replaced (bfgs.h:455-464)
        g[gK[~early]] = g_new[lK[~early]]
        gn = _dot(g[gK], g[gK])
        stop = early | ~(gn >= F(1e-4))
        active[gK[stop]] = False
        go = ~stop
        gG = gK[go]
        if not len(gG):
            continue
        yG, pG, aG = y[go], p[gG], alpha[gG]
        yp = _dot(yG, pG)
        firsts = step[gG] == 0
        if firsts.any():                                                        # set_diagonal(h, alpha y.p / y.y) on the first iteration
            yy = _dot(yG, yG)
            ok = firsts & (np.abs(yy) > EPS)
            if ok.any():
                with np.errstate(all="ignore"):
                    dval = (aG * yp / yy).astype(F)
                ii = gG[ok]
                H[ii[:, None], np.arange(m)[None, :], np.arange(m)[None, :]] = dval[ok][:, None]
        upd = ~(aG * yp < EPS)                                                  # bfgs_update :52-66
        if upd.any():
            iu = gG[upd]
            yu, pu, au, ypu = yG[upd], pG[upd], aG[upd], yp[upd]
            s = np.zeros((len(iu), m), F)
            for j in range(m):
                s = s + H[iu, :, j] * yu[:, j][:, None]
            mhy = -s
            yhy = -_dot(yu, mhy)
            r = (F(1) / (au * ypu)).astype(F)
            c1 = (au * r)[:, None, None]
            c2 = (au * au * (r * r * yhy + r))[:, None, None]
            upd_m = (c1 * (mhy[:, :, None] * pu[:, None, :] + mhy[:, None, :] * pu[:, :, None])
                     + c2 * pu[:, :, None] * pu[:, None, :]).astype(F)
            # the reference keeps the upper triangle only (h(i, j), i <= j: triangular_matrix_index.h): entry (j, i) IS entry (i, j), and
            # ((c2 p_i) p_j) does not round like ((c2 p_j) p_i)
            up = np.triu(upd_m)
            H[iu] = (H[iu] + (up + np.transpose(np.triu(upd_m, 1), (0, 2, 1)))).astype(F)
        step[gG] += 1
        done = step[gG] >= maxiters
        active[gG[done]] = False
        nxt = np.zeros(n, bool); nxt[gG[~done]] = True
        start_iteration(nxt)
    worse = ~(f0 <= f_orig)                                                     # succeeds for NaNs too (bfgs.h:494-498)
    f0[worse], x[worse], g[worse] = f_orig[worse], x_orig[worse], g_orig[worse]
    return f0, x, evals, rounds


MAX_FL = F(3.4028234663852886e+38)


def refine_structure_poses(tree, make_energy, within, X0, maxiters, accurate=False, early_term=False):
    """refine_structure (main/main.cpp:131-171) for many poses in lock step: up to five quasi-Newton runs with the out-of-box slope 10,
    100, ... ; after each run the poses that are `within` the box are finished, the others go on with the slope x 10; a pose that never
    gets inside ends with e = max_fl (:163-164).  make_energy(slope) -> energy_and_forces(coords, idx) (e.g. cnn_energy with that
    slope: --cnn_scoring refinement), within(coords [k,na,3], idx) -> bool [k] (non_cache_cnn::within, within_boxes below).
    -> (e [n], X [n, 7+T], inside [n], function evaluations [n])"""
    x = np.ascontiguousarray(X0, F).copy()
    n = len(x)
    e, evals = np.zeros(n, F), np.zeros(n, np.int64)
    todo = np.arange(n)
    slope = 10.0
    for _ in range(5):
        energy = make_energy(slope)
        ee, xx, ev, _ = minimize_poses(tree, lambda c, i, t=todo: energy(c, t[i]), x[todo], maxiters, accurate, early_term)
        e[todo], x[todo] = ee, xx
        evals[todo] += ev
        ok = within(tree.set_conf(xx)[0], todo)
        todo = todo[~ok]
        if not len(todo):
            break
        slope *= 10
    inside = np.ones(n, bool); inside[todo] = False
    e[todo] = MAX_FL
    return e, x, inside, evals


def within_boxes(heavy, boxes, margin=1e-4):
    """non_cache_cnn::within (lib/non_cache_cnn.cpp:73-76) = inside ANY of the boxes (the CNN grid or the search box); every box is
    (begin, end) with arrays broadcastable to [k,1,3] after indexing by idx (per-pose CNN boxes) or plain [3]"""
    def f(coords, idx):
        res = np.zeros(len(coords), bool)
        for b, e in boxes:
            b, e = np.asarray(b, F), np.asarray(e, F)
            if b.ndim == 2:
                b, e = b[idx][:, None, :], e[idx][:, None, :]
            inside = ((coords >= b - F(margin)) & (coords <= e + F(margin))).all(axis=2)      # gd_within: every heavy atom
            res |= (inside | ~heavy[None, :]).all(axis=1)
        return res
    return f


def box_penalty(coords, heavy, begin, end, slope):
    """non_cache::check_bounds_deriv (lib/non_cache.cpp:102-123) for every atom of every pose -> (penalty [n, na], derivative [n,na,3]);
    zero for hydrogens"""
    b, e = np.asarray(begin, F), np.asarray(end, F)
    lo, hi = coords < b, coords > e
    d = ((np.where(lo, F(-1), F(0)) + np.where(hi, F(1), F(0))).astype(F) * F(slope)).astype(F)
    dist = (np.where(lo, np.abs(coords - b), F(0)) + np.where(hi, np.abs(coords - e), F(0))).astype(F)
    pen = (((dist[:, :, 0] + dist[:, :, 1]) + dist[:, :, 2]) * F(slope)).astype(F)
    d[:, ~heavy] = 0
    pen[:, ~heavy] = 0
    return pen, d


def with_box_penalties(loss, grad, coords, heavy, search_box, cnn_box, slope):
    """the part of non_cache_cnn::eval_deriv (lib/non_cache_cnn.cpp:79-169) around the network, in its order of operations: e = loss,
    then atom by atom e += penalty(search box) + penalty(CNN box); minus_forces = gradient + (d search box + d CNN box); hydrogens 0"""
    p1, d1 = box_penalty(coords, heavy, search_box[0], search_box[1], slope)
    p2, d2 = box_penalty(coords, heavy, cnn_box[0], cnn_box[1], slope)
    e = np.asarray(loss, F).copy()
    for a in np.flatnonzero(heavy):
        e = (e + (p1[:, a] + p2[:, a])).astype(F)
    g = np.asarray(grad, F).copy()
    g[:, ~heavy] = 0
    return e, (g + (d1 + d2)).astype(F)


def route_forces_like_the_reference(grad, heavy):
    """What CNNTorchScorer::score leaves in the model: getGradient builds a list indexed by movable atom (cnn_torch_scorer.cpp:209-227)
    and model::add_minus_forces consumes it COMPACTLY over the non-hydrogen atoms (lib/model.cu:247-259) -- the j-th heavy atom receives
    entry j.  With no hydrogens among the movable atoms this is the identity; with hydrogens the heavy atoms after the first hydrogen
    receive their predecessors' gradients.  Confirmed on the reference's own code with real networks
    (tests/test_oracle_cnn_vs_reference_build.py); grad [k, na, 3] by atom -> [k, na, 3] as the reference's minimiser sees it"""
    hv = np.flatnonzero(heavy)
    if len(hv) == len(heavy):
        return grad
    out = np.zeros_like(grad)
    out[:, hv] = grad[:, :len(hv)]
    return out


def cnn_energy(scorer, types, search_box, slope=10.0, cnn_center=None, reference_force_routing=True):
    """non_cache_cnn::eval_deriv for a batch of poses of ONE ligand: the CNN loss and its atom gradients from ONE gb_cnn_score_grad call,
    plus the out-of-box penalties of the search box and of the CNN's cubic grid.  The grid centre of every pose is the mean of ITS
    heavy atoms in its start conformation -- adjust_center sets it once before a pose's minimisation (lib/non_cache_cnn.cpp:57-68,
    lib/dl_scorer.cpp:196-217) -- unless cnn_centers [n,3] are given; the network itself centres its grid on the pose (--cnn_center
    unset), as TorchModel::forward does.  reference_force_routing: hand the minimiser the forces the reference's scorer leaves in the
    model (route_forces_like_the_reference) so that ligands WITH hydrogens move exactly as under gnina; False = the true per-atom
    gradient.  -> energy_and_forces(coords, idx) for minimize_poses"""
    types = np.ascontiguousarray(types, np.int32)
    heavy = types >= 2
    half = F(scorer.model_info(0).dimension) / F(2)
    state = {"centers": None if cnn_center is None else np.asarray(cnn_center, F).reshape(-1, 3)}

    def energy_and_forces(coords, idx):
        k, na = coords.shape[0], coords.shape[1]
        if state["centers"] is None:                      # first call: every pose in its start conformation
            state["centers"] = heavy_centers(coords, heavy)
        offs = (np.arange(k + 1) * na).astype(np.int32)
        out = scorer.score_grad_batch(coords.reshape(-1, 3), np.tile(types, k), offs)
        loss, grad = np.asarray(out[2], F), np.asarray(out[4], F).reshape(k, na, 3)
        if reference_force_routing:
            grad = route_forces_like_the_reference(grad, heavy)
        c = state["centers"][idx][:, None, :]
        return with_box_penalties(loss, grad, coords, heavy, search_box, (c - half, c + half), slope)
    return energy_and_forces


def heavy_centers(coords, heavy):
    """DLScorer::set_center_from_model (lib/dl_scorer.cpp:196-217): the mean of the heavy movable atoms, summed in atom order"""
    c = np.zeros((len(coords), 3), F)
    for a in np.flatnonzero(heavy):
        c = (c + coords[:, a]).astype(F)
    return (c / F(int(heavy.sum()))).astype(F)
