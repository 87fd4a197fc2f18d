"""TEST INFRASTRUCTURE — golden-vector generator for Optimizer::PoseOptimization ("next" row N4): point edges and the
structural-line (LIL) edges of add_inc/EdgeLIL.h:210-374 (Optimizer.cc:619-693, :976-1007).

Independent numpy restatement (written from src/Optimizer.cc:239-1023 and the vendored g2o sources it calls, not from
oracle/c/orc_pose.cpp): rotations as matrices re-projected through a unit quaternion after every update (what
SE3Quat::normalizeRotation does), numpy.linalg.solve for the damped 6x6 system.  g2o / Eigen are not available in this
image, so neither this file nor the C oracle is pinned by the reference; they pin each other to ~1e-9.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def _quat(R):
    """Eigen::Quaterniond(Matrix3d) + SE3Quat::normalizeRotation -> (w, x, y, z)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        t = math.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q = [w, (R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t]
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        v = [0.0, 0.0, 0.0]
        v[i] = 0.5 * t
        t = 0.5 / t
        w = (R[k, j] - R[j, k]) * t
        v[j] = (R[j, i] + R[i, j]) * t
        v[k] = (R[k, i] + R[i, k]) * t
        q = [w] + v
    q = np.array(q)
    if q[0] < 0:
        q = -q
    return q / math.sqrt(float(q @ q))


def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def _exp_times(x, q, t):
    om, up = x[:3], x[3:]
    th = math.sqrt(float(om @ om))
    O = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    O2 = O @ O
    if th < 0.00001:
        R = np.eye(3) + O + O2
        V = R
    else:
        R = np.eye(3) + math.sin(th) / th * O + (1 - math.cos(th)) / (th * th) * O2
        V = np.eye(3) + (1 - math.cos(th)) / (th * th) * O + (th - math.sin(th)) / th ** 3 * O2
    qe = _quat(R)
    qn = _qmul(qe, q)
    if qn[0] < 0:
        qn = -qn
    qn = qn / math.sqrt(float(qn @ qn))
    return qn, V @ up + _rot(qe) @ t


class _Lil:
    """The LIL edges of a frame as arrays: world points P [m, 5, 3] (segment 1 start / end, segment 2 start / end, cross
    point), observed line equations l1, l2 [m, 3], observed cross point ins [m, 2]."""

    def __init__(self, lils, fx, fy, cx, cy):
        self.sel = np.nonzero(lils["flags"] & 1)[0] if lils is not None and len(lils) else np.zeros(0, int)
        a = lils[self.sel] if len(self.sel) else None
        m = len(self.sel)
        self.P = np.concatenate([a["line1"].reshape(m, 2, 3), a["line2"].reshape(m, 2, 3), a["cross"].reshape(m, 1, 3)], 1) \
            if m else np.zeros((0, 5, 3))
        self.l1 = a["obs1"].astype(np.float64) if m else np.zeros((0, 3))
        self.l2 = a["obs2"].astype(np.float64) if m else np.zeros((0, 3))
        self.ins = a["ins"].astype(np.float64) if m else np.zeros((0, 2))
        self.K = (fx, fy, cx, cy)
        self.level = np.zeros(m, int)
        self.robust = np.ones(m, bool)
        self.err = np.zeros((m, 6))
        self.delta = float(F32(math.sqrt(11.07)))           # float deltaLJL = sqrt(11.07)

    def errors(self, q, t, mask):
        fx, fy, cx, cy = self.K
        X = self.P[mask] @ _rot(q).T + t
        u = X[..., 0] / X[..., 2] * fx + cx
        v = X[..., 1] / X[..., 2] * fy + cy
        l1, l2 = self.l1[mask], self.l2[mask]
        e = np.zeros((int(mask.sum()), 6))
        for r in range(2):
            e[:, r] = u[:, r] * l1[:, 0] + v[:, r] * l1[:, 1] + l1[:, 2]
            e[:, 2 + r] = u[:, 2 + r] * l2[:, 0] + v[:, 2 + r] * l2[:, 1] + l2[:, 2]
        e[:, 4] = self.ins[mask][:, 0] - u[:, 4]
        e[:, 5] = self.ins[mask][:, 1] - v[:, 4]
        return e

    def chi2(self, mask=slice(None)):
        return (self.err[mask] ** 2).sum(1)

    def jacobians(self, q, t, a):
        """d error / d pose, [len(a), 6, 6].  Rows 2 and 3 both take the END point of the second segment: the reference
        reads estimate().segment<3>(9) for its start as well (EdgeLIL.h:276-279)."""
        fx, fy, cx, cy = self.K
        X = self.P[a] @ _rot(q).T + t
        J = np.zeros((len(a), 6, 6))
        for r, (k, l) in enumerate(((0, self.l1), (1, self.l1), (3, self.l2), (3, self.l2))):
            x, y, iz = X[:, k, 0], X[:, k, 1], 1.0 / X[:, k, 2]
            iz2 = iz * iz
            l0, l1 = l[a][:, 0], l[a][:, 1]
            J[:, r] = np.stack([-fx * x * y * iz2 * l0 - fy * (1 + y * y * iz2) * l1,
                                fx * (1 + x * x * iz2) * l0 + fy * x * y * iz2 * l1,
                                -fx * y * iz * l0 + fy * x * iz * l1, fx * iz * l0, fy * iz * l1,
                                (-fx * x * l0 - fy * y * l1) * iz2], 1)
        x, y, iz = X[:, 4, 0], X[:, 4, 1], 1.0 / X[:, 4, 2]
        iz2 = iz * iz
        J[:, 4] = np.stack([x * y * iz2 * fx, -(1 + x * x * iz2) * fx, y * iz * fx, -fx * iz, 0 * x, x * iz2 * fx], 1)
        J[:, 5] = np.stack([(1 + y * y * iz2) * fy, -fy * x * y * iz2, -fy * x * iz, 0 * x, -fy * iz, fy * y * iz2], 1)
        return J


def _huber(c, d):
    big = c > d * d
    s = np.sqrt(np.where(big, c, 1.0))
    return np.where(big, 2 * s * d - d * d, c), np.where(big, d / s, 1.0)


def pose_optimization(Tcw, pts, fx, fy, cx, cy, bf, lils=None):
    """pts: structured (u, v, u_right, inv_sigma2, xw, yw, zw, flags); lils: structured psl_pose_lil records or None.
    Returns (Tcw [4,4] f32, outlier [n] u8, count) and, with lils, additionally lil_outlier [n_lil] u8."""
    fx, fy, cx, cy, bf = (float(F32(v)) for v in (fx, fy, cx, cy, bf))
    T0 = np.asarray(Tcw, F32).astype(np.float64)
    sel = np.nonzero(pts["flags"] & 1)[0]
    n = len(pts)
    outlier = np.zeros(n, np.uint8)
    LL = _Lil(lils, fx, fy, cx, cy)
    lil_outlier = np.zeros(0 if lils is None else len(lils), np.uint8)
    ret = (lambda T, o, c: (T, o, c)) if lils is None else (lambda T, o, c: (T, o, c, lil_outlier))
    n_initial = len(sel) + len(LL.sel)
    if n_initial < 3:
        return ret(np.asarray(Tcw, F32).copy(), outlier, 0)
    stereo = ~(pts["u_right"][sel] < 0)
    obs = np.stack([pts["u"][sel], pts["v"][sel], np.where(stereo, pts["u_right"][sel], 0)], 1).astype(np.float64)
    Xw = np.stack([pts["xw"][sel], pts["yw"][sel], pts["zw"][sel]], 1).astype(np.float64)
    info = pts["inv_sigma2"][sel].astype(np.float64)
    delta = np.where(stereo, float(F32(math.sqrt(7.815))), float(F32(math.sqrt(5.991))))   # const float delta = sqrt(5.991)
    level = np.zeros(len(sel), int)
    robust = np.ones(len(sel), bool)
    err = np.zeros((len(sel), 3))
    q0, t0 = _quat(T0[:3, :3]), T0[:3, 3].copy()

    def errors(q, t, mask):
        X = Xw[mask] @ _rot(q).T + t
        e = np.zeros((mask.sum(), 3))
        st = stereo[mask]
        invz_f = (F32(1.0) / X[:, 2]).astype(F32).astype(np.float64)     # const float invz = 1.0f / z
        r0s, r1s = X[:, 0] * invz_f * fx + cx, X[:, 1] * invz_f * fy + cy
        r0m, r1m = X[:, 0] / X[:, 2] * fx + cx, X[:, 1] / X[:, 2] * fy + cy
        o = obs[mask]
        e[:, 0] = np.where(st, o[:, 0] - r0s, o[:, 0] - r0m)
        e[:, 1] = np.where(st, o[:, 1] - r1s, o[:, 1] - r1m)
        e[:, 2] = np.where(st, o[:, 2] - (r0s - bf * invz_f), 0.0)
        return e

    def chi2(mask=slice(None)):
        return (err[mask] ** 2).sum(1) * info[mask]

    def rho_of(c, d):
        big = c > d * d
        s = np.sqrt(np.where(big, c, 1.0))
        return np.where(big, 2 * s * d - d * d, c), np.where(big, d / s, 1.0)

    def robust_chi():
        a = level == 0
        c = chi2(a)
        r0, _ = rho_of(c, delta[a])
        al = LL.level == 0
        cl = LL.chi2(al)
        rl, _ = _huber(cl, LL.delta)
        return float(np.where(robust[a], r0, c).sum()) + float(np.where(LL.robust[al], rl, cl).sum())

    def system(q, t):
        a = np.nonzero(level == 0)[0]
        X = Xw[a] @ _rot(q).T + t
        x, y, iz = X[:, 0], X[:, 1], 1.0 / X[:, 2]
        iz2 = iz * iz
        J = np.zeros((len(a), 3, 6))
        J[:, 0] = np.stack([x * y * iz2 * fx, -(1 + x * x * iz2) * fx, y * iz * fx, -iz * fx, 0 * x, x * iz2 * fx], 1)
        J[:, 1] = np.stack([(1 + y * y * iz2) * fy, -x * y * iz2 * fy, -x * iz * fy, 0 * x, -iz * fy, y * iz2 * fy], 1)
        J2 = np.stack([J[:, 0, 0] - bf * y * iz2, J[:, 0, 1] + bf * x * iz2, J[:, 0, 2], J[:, 0, 3], 0 * x, J[:, 0, 5] - bf * iz2], 1)
        J[:, 2] = np.where(stereo[a][:, None], J2, 0.0)
        c = chi2(a)
        _, w = rho_of(c, delta[a])
        w = np.where(robust[a], w, 1.0)
        H = np.einsum("n,nij,nik->jk", w * info[a], J, J)
        b = -np.einsum("n,nij,ni->j", w * info[a], J, err[a])
        al = np.nonzero(LL.level == 0)[0]
        if len(al):
            Jl = LL.jacobians(q, t, al)
            _, wl = _huber(LL.chi2(al), LL.delta)
            wl = np.where(LL.robust[al], wl, 1.0)
            H = H + np.einsum("n,nij,nik->jk", wl, Jl, Jl)
            b = b - np.einsum("n,nij,ni->j", wl, Jl, LL.err[al])
        return H, b

    nbad = 0
    q, t = q0, t0
    for it in range(4):
        q, t = q0.copy(), t0.copy()
        lam, ni, bad_steps = 0.0, 2.0, 0
        for iteration in range(10):
            act = level == 0
            err[act] = errors(q, t, act)
            actl = LL.level == 0
            LL.err[actl] = LL.errors(q, t, actl)
            cur = robust_chi()
            ini = cur
            H, b = system(q, t)
            if iteration == 0:
                lam, ni, bad_steps = 1e-5 * float(np.abs(np.diag(H)).max()), 2.0, 0
            rho, qmax = 0.0, 0
            while True:
                Hl = H + lam * np.eye(6)
                ok = bool((np.linalg.eigvalsh(Hl) > 0).all())
                x = np.linalg.solve(Hl, b) if ok else np.zeros(6)
                qn, tn = _exp_times(x, q, t)
                err[act] = errors(qn, tn, act)
                LL.err[actl] = LL.errors(qn, tn, actl)
                tmp = robust_chi() if ok else float(np.finfo(np.float64).max)
                rho = (cur - tmp) / (float(x @ (lam * x + b)) + 1e-3)
                if rho > 0 and math.isfinite(tmp):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    ni = 2.0
                    cur = tmp
                    q, t = qn, tn
                else:
                    lam *= ni
                    ni *= 2
                qmax += 1
                if not (rho < 0 and qmax < 10):
                    break
            if qmax == 10 or rho == 0:
                break
            bad_steps = bad_steps + 1 if (ini - cur) * 1e3 < ini else 0
            if bad_steps >= 3:
                break
        lvl1 = level == 1
        if lvl1.any():
            err[lvl1] = errors(q, t, lvl1)
        c = chi2().astype(F32)
        thr = np.where(stereo, F32(7.815), F32(5.991))
        bad = c > thr
        level = bad.astype(int)
        nbad = int(bad.sum())
        l1 = LL.level == 1
        if l1.any():
            LL.err[l1] = LL.errors(q, t, l1)
        badl = LL.chi2().astype(F32) > F32(11.07)
        LL.level = badl.astype(int)
        if it == 2:
            robust[:] = False
            LL.robust[:] = False
        if n_initial < 10:
            break
    outlier[sel] = bad.astype(np.uint8)
    if len(LL.sel):
        lil_outlier[LL.sel] = badl.astype(np.uint8)
    T = np.eye(4, dtype=F32)
    T[:3, :3] = _rot(q).astype(F32)
    T[:3, 3] = t.astype(F32)
    return ret(T, outlier, n_initial - nbad)
