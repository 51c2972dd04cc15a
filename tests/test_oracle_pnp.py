"""Absolute-pose oracle (oracle/pnp_oracle.py).  pycolmap is not available (parity unpinned): the restatement is validated
against the KNOWN pose of the synthetic correspondences of SURVEY.md section 8(d), and against OpenCV's P3P-RANSAC when cv2
is importable."""
import numpy as np
import pytest

from oracle import pnp_oracle as P


def test_p3p_contains_the_true_pose():
    hits = 0
    for trial in range(40):
        p2d, p3d, cam, Rgt, tgt, _ = P.synthetic_correspondences(M=8, outlier_frac=0, noise_px=0, seed=trial)
        fx, fy, cx, cy = cam
        b = np.stack([(p2d[:3, 0] - cx) / fx, (p2d[:3, 1] - cy) / fy, np.ones(3)], 1).astype(np.float64)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        sols = P.p3p_grunert(b, p3d[:3].astype(np.float64))
        assert 1 <= len(sols) <= 4
        hits += min(P.pose_error(R, t, Rgt, tgt)[0] for R, t in sols) < 1e-2
    assert hits == 40


@pytest.mark.parametrize("seed", [0, 1])
def test_ransac_recovers_pose_with_outliers(seed):
    p2d, p3d, cam, Rgt, tgt, gt_inl = P.synthetic_correspondences(M=1024, seed=seed)
    r = P.absolute_pose_ransac(p2d, p3d, cam, 8.0, iters=128, seed=seed)
    assert r["success"]
    rot, pos = P.pose_error(r["R"], r["t"], Rgt, tgt)
    assert rot < 0.1 and pos < 5e-3, (rot, pos)
    # every planted inlier within 8 px is found; random outliers that happen to land within 8 px are legitimate inliers too
    assert (r["inliers"] & gt_inl).sum() >= 0.99 * gt_inl.sum()
    assert (r["inliers"] & ~gt_inl).sum() <= 0.02 * len(gt_inl)


def test_failure_and_reference_return_shape():
    assert P.absolute_pose_ransac(np.zeros((3, 2)), np.zeros((3, 3)), (1, 1, 0, 0))["success"] is False
    p2d, p3d, cam, Rgt, tgt, _ = P.synthetic_correspondences(M=256, seed=5)
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1.0]])
    c2w, inl = P.estimate_pose(p2d, p3d, K, 8.0, iters=64)
    w2c = np.linalg.inv(c2w)
    assert P.pose_error(w2c[:3, :3], w2c[:3, 3], Rgt, tgt)[0] < 0.2 and inl.dtype == bool and inl.shape == (256,)


def test_against_opencv_when_available():
    cv2 = pytest.importorskip("cv2")
    p2d, p3d, cam, Rgt, tgt, _ = P.synthetic_correspondences(M=512, seed=3)
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1.0]])
    ok, rvec, tvec, inl = cv2.solvePnPRansac(p3d.astype(np.float64), p2d.astype(np.float64), K, None, reprojectionError=8.0,
                                             iterationsCount=500, flags=cv2.SOLVEPNP_P3P)
    assert ok
    Rcv = cv2.Rodrigues(rvec)[0]
    r = P.absolute_pose_ransac(p2d, p3d, cam, 8.0, iters=128, seed=1)
    rot_cv, pos_cv = P.pose_error(Rcv, tvec.reshape(3), Rgt, tgt)
    rot, pos = P.pose_error(r["R"], r["t"], Rgt, tgt)
    # the refined oracle pose is at least as close to the truth as OpenCV's un-refined consensus pose (up to noise)
    assert rot <= rot_cv + 0.05 and pos <= pos_cv + 2e-3, ((rot, pos), (rot_cv, pos_cv))
