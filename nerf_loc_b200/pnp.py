"""Absolute pose from the 2D-3D matches - host mirror of `NerfPoseEstimator.estimate_pose`
(nerf_loc/models/nerf_pose_estimator.py:557-583), which hands the matches to `pycolmap.absolute_pose_estimation`.

Here the whole RANSAC (P3P hypotheses, MSAC scoring, LM local optimisation) runs on the device through
`nlb_pnp_ransac`; only the 12 pose numbers, the success flag and (on request) the inlier mask come back to the host.
COLMAP itself is not available (parity unpinned): the tests check the recovered pose against the known one.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def absolute_pose_estimation(p2d, p3d, camera, ransac_thresh, iters=2048, seed=0, lo_rounds=3):
    """p2d [M,2], p3d [M,3] CUDA tensors; camera = (fx, fy, cx, cy).  Returns a dict shaped like pycolmap's:
    success, R [3,3] / tvec [3] (world -> camera, numpy fp64), inliers [M] bool (CUDA tensor), num_inliers."""
    lib = _lib.load()
    p2d, p3d = _lib.f32(p2d), _lib.f32(p3d)
    M = p2d.shape[0]
    if p3d.shape[0] != M or p2d.shape[1] != 2 or p3d.shape[1] != 3:
        raise ValueError("absolute_pose_estimation: p2d must be [M,2] and p3d [M,3]")
    if M < 4:
        return {"success": False}
    dev = p2d.device
    cam = (ctypes.c_float * 4)(*[float(c) for c in camera])
    pose = torch.empty(12, dtype=torch.float64, device=dev)
    inl = torch.empty(M, dtype=torch.uint8, device=dev)
    res = torch.empty(2, dtype=torch.int32, device=dev)
    nbytes = lib.nlb_pnp_scratch_bytes(iters)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.nlb_pnp_ransac(_lib.ptr(p2d), _lib.ptr(p3d), M, cam, float(ransac_thresh), int(iters), int(seed),
                                  int(lo_rounds), _lib.ptr(pose), _lib.ptr(inl), _lib.ptr(res), _lib.ptr(scratch), nbytes,
                                  _lib.stream()))
    res_h = res.cpu()
    if int(res_h[0]) == 0:
        return {"success": False}
    pose_h = pose.cpu().numpy()
    return {"success": True, "R": pose_h[:9].reshape(3, 3), "tvec": pose_h[9:], "inliers": inl.bool(),
            "num_inliers": int(res_h[1])}


def estimate_pose(matched_kps_2d, matched_kps_3d, K, width, height, ransac_thresh=48, **kw):
    """Same signature and return value as the reference method: (camera-to-world 4x4 numpy, inlier mask numpy) or None."""
    K = K.detach().cpu() if torch.is_tensor(K) else torch.as_tensor(K)
    ret = absolute_pose_estimation(matched_kps_2d, matched_kps_3d, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), ransac_thresh, **kw)
    if not ret["success"]:
        return None
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = ret["R"], ret["tvec"]
    return np.linalg.inv(T), ret["inliers"].cpu().numpy()
