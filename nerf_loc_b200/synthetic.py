"""Deterministic synthetic inputs for parity tests, smoke() and bench.py.

Follows SURVEY.md section 8(d): pinhole intrinsics scaled from 640x480/f=525,
identity query pose, V reference views with a small yaw and baseline, a tilted
plane with a deterministic ripple as the depth maps, uniform images, Gaussian
feature maps.  Weights are generated per parameter *name* (not in module
construction order), so the reference model, the oracle and the CUDA path can
all be loaded with bit-identical values from nothing but a seed.
"""
import math
import zlib

import torch


def synthetic_state_dict(shapes, seed=1234):
    """name -> fp32 tensor.  Matrices/convs ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); LayerNorm-style
    gains ~ 1 + 0.1 U(-1,1); biases ~ 0.1 U(-1,1)."""
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        u = torch.rand(tuple(shape), generator=g) * 2 - 1
        is_gain = name.endswith(".weight") and (
            ".layer_norm." in name or ".norm" in name or (name.startswith("ray_unet.") and ".1.weight" in name)
            or ".bn" in name)
        if is_gain:
            t = 1.0 + 0.1 * u
        elif name.endswith(".weight") and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if name.startswith("ray_unet.trans_conv") and name.endswith(".0.weight"):
                fan_in = shape[0] * shape[2]
            t = u / math.sqrt(fan_in)
        else:
            t = 0.1 * u
        sd[name] = t.float().contiguous()
    return sd


def yaw(deg):
    a = math.radians(deg)
    return torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], dtype=torch.float32)


def make_scene(H=480, W=640, V=8, C=192, seed=1234, near=0.3, far=5.0):
    """Returns the `data`-dict tensors the render path reads (names as nerf_pose_estimator.py:254-279)."""
    g = torch.Generator().manual_seed(seed)
    f = 525.0 * W / 640.0
    K = torch.tensor([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], dtype=torch.float32)
    pose = torch.eye(4)
    poses, Ks = [], []
    for i in range(V):
        T = torch.eye(4)
        T[:3, :3] = yaw(3.0 * i * (1 if i % 2 == 0 else -1))
        T[:3, 3] = torch.tensor([0.1 * (i + 1) * (-1) ** i, 0.02 * i, 0.0])
        poses.append(T)
        Ks.append(K.clone())
    poses = torch.stack(poses)
    Ks = torch.stack(Ks)
    # tilted plane n.X = 2.2 in world coordinates
    n = torch.tensor([0.15, -0.10, 1.0])
    n = n / n.norm()
    vv, uu = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depths = []
    for i in range(V):
        dirs = torch.stack([(uu - K[0, 2]) / K[0, 0], (vv - K[1, 2]) / K[1, 1], torch.ones_like(uu)], -1)
        dw = dirs @ poses[i][:3, :3].t()
        o = poses[i][:3, 3]
        t = (2.2 - (o * n).sum()) / (dw * n).sum(-1)  # camera-z depth since dirs.z == 1
        ripple = 0.02 * torch.sin(0.37 * uu + i) * torch.cos(0.23 * vv)
        depths.append(t + ripple)
    depths = torch.stack(depths).float()
    h4, w4, h8, w8 = H // 4, W // 4, H // 8, W // 8
    scene = {
        "K": K, "pose": pose, "H": H, "W": W,
        "depth_range": torch.tensor([[near, far]], dtype=torch.float32),
        "topk_images": torch.rand(V, 3, H, W, generator=g),
        "topk_depths": depths,
        "topk_poses": poses, "topk_Ks": Ks,
        "feat_fine_src": 0.1 * torch.randn(V, h4, w4, C, generator=g),
        "feat_coarse_src": 0.1 * torch.randn(V, h8, w8, C, generator=g),
        "vis_featmaps": 0.5 * torch.randn(V, 32, h4, w4, generator=g),
        "stride_fine": 4, "stride_coarse": 8, "embedding_a": None,
    }
    return scene


def pixel_rays(K, c2w, px):
    """Unit ray directions through pixel centres px[R,2] = (col,row) (conditional_nerf/utils.py:56-70)."""
    d = torch.stack([(px[:, 0] - K[0, 2]) / K[0, 0], (px[:, 1] - K[1, 2]) / K[1, 1], torch.ones(px.shape[0])], -1)
    d = torch.sum(d[:, None, :] * c2w[:3, :3], -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    o = c2w[:3, 3].expand(d.shape).contiguous()
    return o, d


def random_pixels(H, W, R, seed=7):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(H * W, generator=g)[:R]
    return torch.stack([(idx % W).float(), (idx // W).float()], 1)


def all_pixels(H, W):
    vv, uu = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return torch.stack([uu.reshape(-1), vv.reshape(-1)], 1)
