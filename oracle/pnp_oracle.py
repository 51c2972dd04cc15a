"""TEST INFRASTRUCTURE - CPU restatement (numpy, fp64) of the absolute-pose stage that follows matching
(nerf_loc/models/nerf_pose_estimator.py:557-583: `pycolmap.absolute_pose_estimation(p2d, p3d, PINHOLE camera, thresh)`).

PARITY UNPINNED: pycolmap / COLMAP is a third-party dependency that is neither vendored under /root/reference nor installable
here (requirements.txt:11, `pycolmap>=0.1.0`, unpinned), and the reference holds no golden vectors for it.  This file restates
COLMAP's published algorithm shape - P3P minimal solver inside RANSAC, scoring by squared reprojection error against the
threshold, local optimisation / final refinement of the pose on the inliers by Levenberg-Marquardt - and the tests validate
both this oracle and the CUDA path against the KNOWN synthetic ground-truth pose, not against COLMAP.

Conventions: p2d [M,2] pixels, p3d [M,3] world, camera (fx, fy, cx, cy) PINHOLE; the pose is world->camera (R [3,3], t [3]);
`estimate_pose` returns the camera-to-world 4x4 like the reference (nerf_pose_estimator.py:577-583).
"""
import numpy as np


def _poly_mul(a, b):
    return np.convolve(a, b)


def p3p_grunert(j, P):
    """Grunert's three-point pose: j [3,3] unit bearings (camera frame), P [3,3] world points.
    Returns a list of (R, t).  With s2 = u s1, s3 = v s1 the law-of-cosines system reduces to a quartic in v whose
    coefficients are formed here by polynomial arithmetic (no closed-form coefficient tables)."""
    a2 = np.sum((P[1] - P[2]) ** 2)
    b2 = np.sum((P[0] - P[2]) ** 2)
    c2 = np.sum((P[0] - P[1]) ** 2)
    if min(a2, b2, c2) < 1e-18:
        return []
    ca, cb, cg = j[1] @ j[2], j[0] @ j[2], j[0] @ j[1]
    Kq = (a2 - c2) / b2
    # u = N(v) / D(v);  polynomials are stored highest power first
    N = np.array([Kq - 1.0, -2.0 * Kq * cb, 1.0 + Kq])
    D = np.array([-2.0 * ca, 2.0 * cg])
    Q = np.array([-c2 / b2, 2.0 * c2 / b2 * cb, 1.0 - c2 / b2])      # 1 - (c^2/b^2)(1 + v^2 - 2 v cos(beta))
    # (3'): D^2 Q + N^2 - 2 cos(gamma) N D = 0
    quartic = _poly_mul(_poly_mul(D, D), Q) + _poly_mul(N, N) - 2.0 * cg * np.concatenate([[0.0], _poly_mul(N, D)])
    if abs(quartic[0]) < 1e-14:
        return []
    out = []
    for v in np.roots(quartic):
        if abs(v.imag) > 1e-7 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = v.real
        den = 2.0 * (cg - v * ca)
        if abs(den) < 1e-12:
            continue
        u = ((Kq - 1.0) * v * v - 2.0 * Kq * cb * v + 1.0 + Kq) / den
        w = 1.0 + v * v - 2.0 * v * cb
        if u <= 0 or w <= 0:
            continue
        s1 = np.sqrt(b2 / w)
        X = np.stack([s1 * j[0], u * s1 * j[1], v * s1 * j[2]])
        out.append(_align3(P, X))
    return [rt for rt in out if rt is not None]


def _triad(A):
    e1 = A[1] - A[0]
    n1 = np.linalg.norm(e1)
    e3 = np.cross(e1, A[2] - A[0])
    n3 = np.linalg.norm(e3)
    if n1 < 1e-12 or n3 < 1e-12:
        return None
    e1, e3 = e1 / n1, e3 / n3
    return np.stack([e1, np.cross(e3, e1), e3], 1)


def _align3(P, X):
    """Rigid transform with X_i = R P_i + t from three correspondences (orthonormal triads)."""
    Fw, Fc = _triad(P), _triad(X)
    if Fw is None or Fc is None:
        return None
    R = Fc @ Fw.T
    return R, X[0] - R @ P[0]


def reproj_err2(R, t, p2d, p3d, cam):
    fx, fy, cx, cy = cam
    Xc = p3d @ R.T + t
    z = Xc[:, 2]
    ok = z > 1e-9
    zs = np.where(ok, z, 1.0)
    e = (fx * Xc[:, 0] / zs + cx - p2d[:, 0]) ** 2 + (fy * Xc[:, 1] / zs + cy - p2d[:, 1]) ** 2
    return np.where(ok, e, np.inf)


def _expm_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def refine(R, t, p2d, p3d, cam, iters=10):
    """Levenberg-Marquardt on the reprojection error over the given correspondences; left-multiplicative update
    X' <- Exp(w) X' + tau."""
    fx, fy, cx, cy = cam
    lam = 1e-3

    def cost(R, t):
        return float(np.sum(np.minimum(reproj_err2(R, t, p2d, p3d, cam), 1e12)))

    c0 = cost(R, t)
    for _ in range(iters):
        Xc = p3d @ R.T + t
        x, y, z = Xc[:, 0], Xc[:, 1], np.maximum(Xc[:, 2], 1e-9)
        r = np.stack([fx * x / z + cx - p2d[:, 0], fy * y / z + cy - p2d[:, 1]], 1).reshape(-1)
        J = np.zeros((len(p3d), 2, 6))
        # d proj / d X'
        dudX = np.stack([fx / z, np.zeros_like(z), -fx * x / z ** 2], 1)
        dvdX = np.stack([np.zeros_like(z), fy / z, -fy * y / z ** 2], 1)
        # d X' / d w = -[X']x ; d X' / d tau = I
        for row, d in ((0, dudX), (1, dvdX)):
            J[:, row, 0] = d[:, 2] * y - d[:, 1] * z
            J[:, row, 1] = d[:, 0] * z - d[:, 2] * x
            J[:, row, 2] = d[:, 1] * x - d[:, 0] * y
            J[:, row, 3:] = d
        J = J.reshape(-1, 6)
        H, g = J.T @ J, J.T @ r
        for _try in range(8):
            try:
                dlt = -np.linalg.solve(H + lam * np.diag(np.diag(H)), g)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            E = _expm_so3(dlt[:3])
            R2, t2 = E @ R, E @ t + dlt[3:]
            c1 = cost(R2, t2)
            if c1 < c0:
                R, t, c0, lam = R2, t2, c1, max(lam * 0.1, 1e-9)
                break
            lam *= 10
        else:
            break
    return R, t


def absolute_pose_ransac(p2d, p3d, cam, thresh=8.0, iters=2048, seed=0, lo_rounds=3):
    """Returns dict(success, R, t, inliers [M] bool, num_inliers)."""
    p2d, p3d = np.asarray(p2d, np.float64), np.asarray(p3d, np.float64)
    M = len(p2d)
    if M < 4:
        return dict(success=False)
    fx, fy, cx, cy = cam
    bear = np.stack([(p2d[:, 0] - cx) / fx, (p2d[:, 1] - cy) / fy, np.ones(M)], 1)
    bear /= np.linalg.norm(bear, axis=1, keepdims=True)
    rng = np.random.default_rng(seed)
    thr2 = thresh * thresh
    best = (np.inf, None, None)
    for _ in range(iters):
        ids = rng.choice(M, 3, replace=False)
        for R, t in p3p_grunert(bear[ids], p3d[ids]):
            score = float(np.sum(np.minimum(reproj_err2(R, t, p2d, p3d, cam), thr2)))   # MSAC
            if score < best[0]:
                best = (score, R, t)
    if best[1] is None:
        return dict(success=False)
    _, R, t = best
    inl = reproj_err2(R, t, p2d, p3d, cam) < thr2
    for _ in range(lo_rounds):
        if inl.sum() < 4:
            break
        R, t = refine(R, t, p2d[inl], p3d[inl], cam)
        inl = reproj_err2(R, t, p2d, p3d, cam) < thr2
    return dict(success=bool(inl.sum() >= 4), R=R, t=t, inliers=inl, num_inliers=int(inl.sum()))


def estimate_pose(p2d, p3d, K, thresh=8.0, **kw):
    """nerf_pose_estimator.py:557-583: camera-to-world 4x4 and the inlier mask, or None."""
    ret = absolute_pose_ransac(p2d, p3d, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), thresh, **kw)
    if not ret["success"]:
        return None
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = ret["R"], ret["t"]
    return np.linalg.inv(T), ret["inliers"]


def synthetic_correspondences(M=2048, outlier_frac=0.25, noise_px=1.0, seed=0, W=640, H=480):
    """planted pairs with exact projections + 1 px Gaussian noise, 25 % outliers, known GT pose."""
    rng = np.random.default_rng(seed)
    cam = (525.0, 525.0, W / 2.0, H / 2.0)
    ang = np.deg2rad(rng.uniform(-10, 10, 3))
    Rgt = _expm_so3(ang)
    tgt = rng.uniform(-0.3, 0.3, 3)
    # points in front of the camera
    uv = np.stack([rng.uniform(0, W, M), rng.uniform(0, H, M)], 1)
    z = rng.uniform(1.0, 4.0, M)
    Xc = np.stack([(uv[:, 0] - cam[2]) / cam[0] * z, (uv[:, 1] - cam[3]) / cam[1] * z, z], 1)
    p3d = (Xc - tgt) @ Rgt          # X_w = R^T (X_c - t)
    p2d = uv + rng.normal(0, noise_px, (M, 2))
    n_out = int(outlier_frac * M)
    out_ids = rng.choice(M, n_out, replace=False)
    p2d[out_ids] = np.stack([rng.uniform(0, W, n_out), rng.uniform(0, H, n_out)], 1)
    gt_inl = np.ones(M, bool)
    gt_inl[out_ids] = False
    return p2d.astype(np.float32), p3d.astype(np.float32), cam, Rgt, tgt, gt_inl


def pose_error(R, t, Rgt, tgt):
    """(rotation error in degrees, camera-centre error in scene units)"""
    cosv = np.clip((np.trace(R @ Rgt.T) - 1) / 2, -1, 1)
    return float(np.degrees(np.arccos(cosv))), float(np.linalg.norm(-R.T @ t + Rgt.T @ tgt))
