"""Independent float64 check of the plane-to-plane refinement (LidarPlaneFactor FCCF.cpp:178-208, ceres_refine 210-249):
the same cost minimised by scipy.optimize.least_squares over a 6-parameter tangent (rotation vector, translation).
Used by the parity tests to judge hypotheses on which the GPU's and the oracle's Levenberg-Marquardt trajectories end
more than the 0.01 degree / 1 mm bar apart."""
import numpy as np


def _rot(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], float)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def problem(T, planes1, planes2, pairs):
    """Rows (n1, p1, n2', p2', w) of the refinement started from hypothesis T (quick_verify FCCF.cpp:680-783: the
    target planes are moved by T in float32 first; importance = 2 min(size) / (int-truncated size sums))."""
    T = np.asarray(T, np.float32).reshape(4, 4)
    p1 = np.asarray(planes1, np.float32).reshape(-1, 7)
    p2 = np.asarray(planes2, np.float32).reshape(-1, 7)
    fs1 = 0
    for s in p1[:, 6]:
        fs1 = int(np.float32(fs1) + s)
    fs2 = 0
    for s in p2[:, 6]:
        fs2 = int(np.float32(fs2) + s)
    rows = []
    for i1, i2 in np.asarray(pairs).reshape(-1, 2):
        c2 = (T[:3, :3] @ p2[i2, :3] + T[:3, 3]).astype(np.float32)
        n2 = (T[:3, :3] @ p2[i2, 3:6]).astype(np.float32)
        w = np.float32(2 * min(p1[i1, 6], p2[i2, 6])) / np.float32(fs1 + fs2)
        rows.append((p1[i1, 3:6].astype(float), p1[i1, :3].astype(float), n2.astype(float), c2.astype(float), float(w)))
    return rows


def residuals(x, rows):
    R, t = _rot(x[:3]), x[3:]
    r = []
    for n1, q1, n2, q2, w in rows:
        n2r = R @ n2
        r.append(w * np.linalg.norm(np.cross(n1, n2r)))
        r.append(w * abs(n1 @ q1 - n2r @ (R @ q2 + t)))
    return np.array(r)


def cost_of(Tref, T0, rows):
    """0.5 * sum r^2 of the increment D with Tref = D * T0."""
    D = np.asarray(Tref, float).reshape(4, 4) @ np.linalg.inv(np.asarray(T0, float).reshape(4, 4))
    R, t = D[:3, :3], D[:3, 3]
    r = []
    for n1, q1, n2, q2, w in rows:
        n2r = R @ n2
        r.append(w * np.linalg.norm(np.cross(n1, n2r)))
        r.append(w * abs(n1 @ q1 - n2r @ (R @ q2 + t)))
    return 0.5 * float(np.sum(np.square(r)))


def solve(T0, rows):
    """(T_scipy, cost, smallest / largest eigenvalue of J^T J at the solution)."""
    from scipy.optimize import least_squares

    res = least_squares(residuals, np.zeros(6), args=(rows,), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    D = np.eye(4)
    D[:3, :3] = _rot(res.x[:3]); D[:3, 3] = res.x[3:]
    ev = np.linalg.eigvalsh(res.jac.T @ res.jac)
    return D @ np.asarray(T0, float).reshape(4, 4), float(res.cost), float(ev[0] / max(ev[-1], 1e-300))
