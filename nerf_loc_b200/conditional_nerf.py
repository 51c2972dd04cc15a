"""Drop-in `ConditionalNeRF` backed by the sm_100a kernels (inference path).

Mirrors nerf_loc/models/conditional_nerf/model.py:29-713: same constructor, same parameter names (a reference
checkpoint loads with `load_state_dict`), same method names, argument meaning and return dicts, and the same
per-frame cache protocol (`support_neural_points = None`, `multiview_aggregator.vis_featmaps = None` force a rebuild,
nerf_pose_estimator.py:289-290).  The arithmetic of `query` / `render_rays` runs in libnerfloc_b200.so; torch is used
for device memory, streams and the tiny per-frame camera algebra.  There is no CPU path.

Not covered here: autograd through the kernels (training, pose optimizer) and
render.N_importance > 0 (0 in every shipped config, nerf_loc/configs/__init__.py:56).
"""
import copy
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, params
from .depth_fusion import DepthFusionNet
from .knn import KnnIndex


def _seq_linear(dims, act):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        layers.append(act)
    return layers


class MixtureLogisticsDistDecoder(nn.Module):
    """Parameter container for visibility_decoder.py:53-97 (evaluated inside aggregate_kernel)."""

    def __init__(self, cfg=None):
        super().__init__()
        def head(k, last):
            return nn.Sequential(nn.Linear(32, 32), nn.ELU(), nn.Linear(32, 32), nn.ELU(), nn.Linear(32, k), last)
        self.mean_decoder = head(2, nn.Softplus())
        self.var_decoder = head(2, nn.Softplus())
        self.aw_decoder = head(1, nn.Sigmoid())
        self.vis_decoder = head(1, nn.Sigmoid())


class MultiviewFeatureAggregator(nn.Module):
    """multiview_aggregator.py:21-222.  `forward` keeps the reference signature."""

    def __init__(self, args, in_channels, out_channels, hidden_dim=64):
        super().__init__()
        self.args = args
        self.depth_fusion = DepthFusionNet(in_channels=in_channels)
        self.vis_featmaps = None
        self.dist_decoder = MixtureLogisticsDistDecoder({})
        self.out_fc = nn.Sequential(nn.Linear((in_channels + 3) * 2 + 2 + 1, hidden_dim), nn.ELU(),
                                    nn.Linear(hidden_dim, out_channels), nn.ELU())
        self._owner = None  # set by ConditionalNeRF (weights are packed there)

    def forward(self, sampled_points, intrinsics, extrinsics, images, featmaps, depths, depth_range):
        if self.vis_featmaps is None:
            with torch.no_grad():
                self.vis_featmaps = self.depth_fusion(images, featmaps, depths, intrinsics, extrinsics, depth_range)
        owner = self._owner()
        maps = owner._maps(intrinsics, extrinsics, images, featmaps.permute(0, 2, 3, 1), self.vis_featmaps, depth_range)
        return owner._aggregate(maps, sampled_points)


class RayUnet(nn.Module):
    """Parameter container for ray_unet.py:5-52 (evaluated inside ray_kernel)."""

    def __init__(self, in_channels, n_samples):
        super().__init__()
        def block(conv, c, s):
            return nn.Sequential(conv, nn.LayerNorm([c, s]), nn.ELU())
        S = n_samples
        self.conv1 = block(nn.Conv1d(in_channels, 64, 3, 1, padding=1), 64, S)
        self.conv2 = block(nn.Conv1d(64, 128, 3, 1, padding=1), 128, S // 2)
        self.conv3 = block(nn.Conv1d(128, 128, 3, 1, padding=1), 128, S // 4)
        self.trans_conv3 = block(nn.ConvTranspose1d(128, 128, 3, 2, padding=1, output_padding=1), 128, S // 4)
        self.trans_conv2 = block(nn.ConvTranspose1d(256, 64, 3, 2, padding=1, output_padding=1), 64, S // 2)
        self.trans_conv1 = block(nn.ConvTranspose1d(128, 32, 3, 2, padding=1, output_padding=1), 32, S)
        self.conv_out = block(nn.Conv1d(in_channels + 32, in_channels, 3, 1, padding=1), in_channels, S)


def get_rays(H, W, K, c2w):
    """conditional_nerf/utils.py:56-70 (unit directions)."""
    dev = K.device
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev), indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], (jj - K[1][2]) / K[1][1], torch.ones_like(ii)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


class _Maps:
    """Device-side view of the reference images / feature maps of one pyramid level (everything nlb_scene needs except
    the support points)."""

    def __init__(self, Ks, c2ws, images, feat_cl, vis_maps, depth_range):
        dev = images.device
        V, _, H, W = images.shape
        self.V, self.H, self.W = V, H, W
        self.h, self.w = feat_cl.shape[1], feat_cl.shape[2]
        if feat_cl.shape[3] != 192 or vis_maps.shape[1] != 32:
            raise RuntimeError("nerfloc_b200 kernels are built for 192-channel features and 32-channel visibility maps")
        self.vh, self.vw = vis_maps.shape[-2], vis_maps.shape[-1]
        self.images = torch.cat([images.permute(0, 2, 3, 1), torch.zeros(V, H, W, 1, device=dev)], -1).float().contiguous()
        self.feat = _lib.f32(feat_cl)
        self.vis = _lib.f32(vis_maps.permute(0, 2, 3, 1))
        Ks, c2ws = Ks.float(), c2ws.float()
        w2c = torch.inverse(c2ws)
        Kh = torch.eye(4, device=dev).expand(V, 4, 4).clone()
        Kh[:, :3, :3] = Ks
        cams = torch.zeros(V, 32, device=dev)
        cams[:, 0:12] = Kh.bmm(w2c)[:, :3].reshape(V, 12)       # ibrnet.py:183-184
        cams[:, 12:24] = (Ks @ w2c[:, :3]).reshape(V, 12)        # depth_fusion.py:90
        cams[:, 24:27] = c2ws[:, :3, 3]
        self.cams = cams.contiguous()
        self.near, self.far = float(depth_range[0]), float(depth_range[1])
        self.featb = None  # [V*h*w, 32], filled by ConditionalNeRF._level_scene when a frame is rendered

    def fill(self, sc):
        sc.V, sc.H, sc.W, sc.h, sc.w = self.V, self.H, self.W, self.h, self.w
        sc.vh, sc.vw = self.vh, self.vw
        sc.near_plane, sc.far_plane = self.near, self.far
        sc.images, sc.featmaps = self.images.data_ptr(), self.feat.data_ptr()
        sc.vis_maps, sc.cams = self.vis.data_ptr(), self.cams.data_ptr()


class _Support:
    """Per-frame device buffers over one level's support neural points."""

    def __init__(self, packed, S, sup):
        L = _lib.load()
        xyz, feat = _lib.f32(sup["xyz"]), _lib.f32(sup["feature"])
        conf, dirs = _lib.f32(sup["confidence"]), _lib.f32(sup["direction"])
        self.M = xyz.shape[0]
        if self.M < 1:
            raise RuntimeError("zero support neural points (no valid reference depth)")
        self.index = KnnIndex(xyz)
        self.pre = torch.empty(self.M, 128, device=xyz.device)
        self.geo = torch.empty(self.M, 8, device=xyz.device)
        _lib.check(L.nlb_support_prepare(_lib.ptr(packed), S, _lib.ptr(xyz), _lib.ptr(feat), _lib.ptr(conf),
                                         _lib.ptr(dirs), self.M, _lib.ptr(self.pre), _lib.ptr(self.geo), _lib.stream()))

    def fill(self, sc):
        sc.M = self.M
        sc.sup_pre, sc.sup_geo, sc.knn_index = self.pre.data_ptr(), self.geo.data_ptr(), self.index.buf.data_ptr()


class ConditionalNeRF(nn.Module):
    def __init__(self, args, activation_func=None):
        super().__init__()
        import weakref
        act = activation_func if activation_func is not None else nn.LeakyReLU(inplace=True)
        self.args = copy.deepcopy(args)
        C, W = self.args.backbone2d_fpn_dim, self.args.model_3d_hidden_dim
        if C != 192 or W != 128 or self.args.multires != 10 or self.args.i_embed != 0:
            raise RuntimeError("nerfloc_b200 kernels are built for backbone2d_fpn_dim=192, model_3d_hidden_dim=128, multires=10")
        self.n_samples = self.args.render.N_samples + self.args.render.N_importance
        view_dim = 3 + 3 * 2 * self.args.multires_views
        self.ray_diff_fc = nn.Sequential(nn.Linear(4, 16), act, nn.Linear(16, view_dim), act)
        self.multiview_aggregator = MultiviewFeatureAggregator(args, in_channels=C, out_channels=W)
        self.multiview_aggregator._owner = weakref.ref(self)
        self.confidence_mlp = nn.Sequential(nn.Linear(W, 64), act, nn.Linear(64, 1), nn.Sigmoid())
        self.keypoint_head = nn.Sequential(nn.Linear(C, 1), nn.Sigmoid())
        self.base_mlp = nn.Sequential(*_seq_linear([3 + C + 63 + view_dim, W, W, W], act))
        self.base_mlp_attn = _AttnParams(W)
        self.base_mlp_agg_weight = nn.Sequential(nn.Linear(W, W), act, nn.Linear(W, 1))
        self.ray_unet = RayUnet(W, self.n_samples)
        self.sigma_mlp = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        if self.args.render.render_feature:
            self.feat_mlp = nn.Sequential(nn.Linear(W, W), act, nn.Linear(W, C))
        self.rgb_blending_mlp = nn.Sequential(nn.Linear(W + (3 + C) + 1 + 4, 32), act, nn.Linear(32, 16), act, nn.Linear(16, 1))
        if self.args.render.use_render_uncertainty:
            self.beta_mlp = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
            self.beta_min = 0.1
        if self.args.use_scene_coord_memorization:
            raise NotImplementedError("use_scene_coord_memorization (per-scene finetuning) is outside the hot path")
        self.proj_layer_3d_coarse = nn.Linear(W + 3 + C, self.args.matcher_hidden_dim)
        self.proj_layer_3d_fine = nn.Linear(W + 3 + C, self.args.matcher_hidden_dim)
        self._support_points = None
        self._packed = None
        self._packed_key = None
        self._packed_src = None
        self._frame = {}
        self.chunk_rays = 76960  # cap on the rays per kernel wave inside nlb_render_rays (148 SMs x 520: a 640x480 frame is four equal waves; 5.0 KB of scratch per sample at V = 8)

    # ---- per-frame cache protocol -------------------------------------------------------------------------------------
    @property
    def support_neural_points(self):
        return self._support_points

    @support_neural_points.setter
    def support_neural_points(self, v):
        self._support_points = v
        self._frame = {}

    # ---- weights ----------------------------------------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        # .cuda() / .to() replace the parameter tensors: drop the cached references to the old ones
        self._packed_src = None
        return super()._apply(fn, *args, **kwargs)

    def packed_weights(self):
        """Flat device buffer in the kernels' layout; re-packed when any parameter changed."""
        L = _lib.load()
        if self._packed_src is None:
            # parameter tensors in the packer's order, looked up once per model (state_dict() + ~100 keys per call otherwise)
            sd = self.state_dict()
            self._packed_src = [sd[n] for n in params.conditional_nerf_shapes(self.n_samples).keys()]
        tensors = self._packed_src
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is None or key != self._packed_key:
            dev = tensors[0].device
            keep = [_lib.f32(t) for t in tensors]
            arr = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            n = L.nlb_render_weights_floats(self.n_samples)
            packed = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.check(L.nlb_render_pack_weights(arr, len(keep), self.n_samples, _lib.ptr(packed), n, _lib.stream()))
            torch.cuda.current_stream().synchronize()
            self._packed, self._packed_key = packed, key
        return self._packed

    # ---- scene assembly -----------------------------------------------------------------------------------------------------
    def _maps(self, Ks, c2ws, images, feat_cl, vis_maps, depth_range):
        return _Maps(Ks, c2ws, images, feat_cl, vis_maps, depth_range)

    def _vis_maps(self, data):
        agg = self.multiview_aggregator
        if agg.vis_featmaps is None:
            with torch.no_grad():
                agg.vis_featmaps = agg.depth_fusion(
                    data['topk_images'], data['feat_fine_src'].permute(0, 3, 1, 2), data['topk_depths'],
                    data['topk_Ks'], data['topk_poses'], data['depth_range'][0])
        return agg.vis_featmaps

    def _level_maps(self, data, level):
        key = "maps_" + level
        if key not in self._frame:
            self._frame[key] = _Maps(data['topk_Ks'], data['topk_poses'], data['topk_images'],
                                     data['feat_%s_src' % level], self._vis_maps(data), data['depth_range'][0])
        return self._frame[key]

    def _level_scene(self, data, level, query_pose=None):
        if self.support_neural_points is None:
            self.build_support_neural_points(data)
        maps = self._level_maps(data, level)
        key = "sup_" + level
        if key not in self._frame:
            self._frame[key] = _Support(self.packed_weights(), self.n_samples, self.support_neural_points[level])
        sc = _lib.NlbScene()
        maps.fill(sc)
        self._frame[key].fill(sc)
        if query_pose is not None:
            # rendering: the colour-blend layer-1 projection of the feature maps, once per frame and level
            if maps.featb is None:
                maps.featb = torch.empty(maps.V * maps.h * maps.w, 32, device=maps.feat.device)
                _lib.check(_lib.load().nlb_blend_prepare(_lib.ptr(self.packed_weights()), self.n_samples, _lib.ptr(maps.feat),
                                                         maps.V * maps.h * maps.w, _lib.ptr(maps.featb), _lib.stream()))
            sc.featmaps_blend = maps.featb.data_ptr()
            c = query_pose[:3, 3].detach().float().cpu()
            sc.query_center[0], sc.query_center[1], sc.query_center[2] = float(c[0]), float(c[1]), float(c[2])
        return sc, maps, self._frame[key]

    def _aggregate(self, maps, points):
        """MultiviewFeatureAggregator.forward on the device: (out [N,128], rgb_feat [N,V,195], vis [N,V,1])."""
        L = _lib.load()
        pts = _lib.f32(points)
        N = pts.shape[0]
        sc = _lib.NlbScene()
        maps.fill(sc)
        out = torch.empty(N, 128, device=pts.device)
        mvf = torch.empty(N, maps.V, 195, device=pts.device)
        mvv = torch.empty(N, maps.V, 1, device=pts.device)
        _lib.check(L.nlb_aggregate_points(ctypes.byref(sc), _lib.ptr(self.packed_weights()), self.n_samples, _lib.ptr(pts),
                                          N, _lib.ptr(out), _lib.ptr(mvf), _lib.ptr(mvv), _lib.stream()))
        return out, mvf, mvv

    # ---- support neural points (model.py:137-275) -----------------------------------------------------------------------------
    def estimate_neural_points_confidence(self, points, intrinsics, extrinsics, images, featmaps, depths, depth_range):
        L = _lib.load()
        mv, _, _ = self.multiview_aggregator(points, intrinsics, extrinsics, images, featmaps, depths, depth_range)
        N = mv.shape[0]
        conf = torch.empty(N, 1, device=mv.device)
        tmp = torch.empty(N, 64, device=mv.device)
        _lib.check(L.nlb_confidence_head(_lib.ptr(self.packed_weights()), self.n_samples, _lib.ptr(mv), N, _lib.ptr(conf),
                                         _lib.ptr(tmp), _lib.stream()))
        return conf

    @torch.no_grad()
    def build_support_neural_points(self, data):
        d = data['topk_depths']
        fc, xc, nc, dc = self.backproject_support_frame(data['topk_images'], data['feat_coarse_src'], d,
                                                        data['topk_Ks'], data['topk_poses'], stride=data['stride_coarse'])
        ff, xf, nf, df = self.backproject_support_frame(data['topk_images'], data['feat_fine_src'], d,
                                                        data['topk_Ks'], data['topk_poses'], stride=data['stride_fine'])
        conf_f = self.estimate_neural_points_confidence(
            xf, data['topk_Ks'], data['topk_poses'], data['topk_images'], data['feat_fine_src'].permute(0, 3, 1, 2), d,
            data['depth_range'][0])
        kp = self.keypoint_head(fc[:, 3:])
        self.support_neural_points = {
            'coarse': {'xyz': xc, 'xyz_ndc': nc, 'feature': fc, 'confidence': torch.ones_like(xc[:, :1]),
                       'direction': dc, 'keypoint_score': kp},
            'fine': {'xyz': xf, 'xyz_ndc': nf, 'feature': ff, 'confidence': conf_f, 'direction': df},
        }
        if len(xc) == 0:
            print(f"Error: zero support_neural_points {data.get('scene')} : {data.get('filename')}")

    def backproject_support_frame(self, imgs, feats, depths, Ks, c2ws, stride=1):
        """model.py:203-265: depth pixels of every reference view -> (feature [M,3+C], xyz_world, xyz_ref, direction [M,4]).
        Point order (view-major, then row-major nonzero order) defines the KNN indices.

        On the device the per-point arithmetic runs in `nlb_backproject_points`, which reproduces the rounding of the
        reference's host operators: xyz / xyz_ndc / direction are bit-identical with the reference run on the CPU (a
        cuBLAS product differs by an ulp in places, which flips near-tied nearest neighbours).  The 3x3 / 4x4 matrices
        come from the reference's own host ops."""
        if not imgs.is_cuda:
            return self._backproject_support_frame_host(imgs, feats, depths, Ks, c2ws, stride)
        L = _lib.load()
        Ks_h, c2ws_h = Ks.detach().float().cpu(), c2ws.detach().float().cpu()
        w2c_ref = torch.inverse(c2ws_h[0])
        outs = ([], [], [], [])
        for v, (img, feat, depth) in enumerate(zip(imgs, feats, depths)):
            H, W = int(img.shape[-2] / stride), int(img.shape[-1] / stride)
            K = Ks_h[v].clone()
            K[:2] /= stride
            c2w = c2ws_h[v]
            mats = torch.cat([torch.inverse(K).reshape(-1), c2w[:3, :3].reshape(-1), c2w[:3, 3],
                              torch.matmul(w2c_ref, c2w)[:3].reshape(-1),
                              torch.stack([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])]).contiguous()
            dm = F.interpolate(depth[None, None], size=(H, W)).squeeze()
            im = F.interpolate(img[None], size=(H, W)).squeeze().permute(1, 2, 0)
            vv, uu = torch.nonzero(dm > 0, as_tuple=True)
            zz = dm[vv, uu].float().contiguous()
            M = zz.shape[0]
            world = torch.empty(M, 3, device=zz.device)
            ref = torch.empty(M, 3, device=zz.device)
            direction = torch.empty(M, 4, device=zz.device)
            _lib.check(L.nlb_backproject_points(mats.data_ptr(), _lib.ptr(uu.contiguous()), _lib.ptr(vv.contiguous()),
                                                _lib.ptr(zz), M, _lib.ptr(world), _lib.ptr(ref), _lib.ptr(direction),
                                                _lib.stream()))
            outs[0].append(torch.cat([im[vv, uu], feat[vv, uu]], 1))
            outs[1].append(world)
            outs[2].append(ref)
            outs[3].append(direction)
        return tuple(torch.cat(o) for o in outs)

    def _backproject_support_frame_host(self, imgs, feats, depths, Ks, c2ws, stride=1):
        """The same function on host tensors with the reference's own operators (what the device kernel is pinned to)."""
        outs = ([], [], [], [])
        w2c_ref = torch.inverse(c2ws[0])
        for img, feat, depth, K, c2w in zip(imgs, feats, depths, Ks, c2ws):
            H, W = int(img.shape[-2] / stride), int(img.shape[-1] / stride)
            K = K.clone()
            K[:2] /= stride
            dm = F.interpolate(depth[None, None], size=(H, W)).squeeze()
            im = F.interpolate(img[None], size=(H, W)).squeeze().permute(1, 2, 0)
            vv, uu = torch.nonzero(dm > 0, as_tuple=True)
            zz = dm[vv, uu]
            uv1 = torch.stack([uu, vv, torch.ones_like(uu)], 0).float()
            cam = torch.matmul(torch.inverse(K), uv1) * zz
            cam_h = torch.cat([cam, torch.ones_like(cam[:1])])
            world = torch.matmul(c2w[:3, :3], cam) + c2w[:3, 3:]
            ref = torch.matmul(torch.matmul(w2c_ref, c2w), cam_h)[:3]
            _, rd = get_rays(H, W, K, c2w)
            outs[0].append(torch.cat([im[vv, uu], feat[vv, uu]], 1))
            outs[1].append(world.t())
            outs[2].append(ref.t())
            outs[3].append(torch.cat([rd[vv, uu], zz.view(-1, 1)], 1))
        return tuple(torch.cat(o) for o in outs)

    def sample_points_3d(self):
        sup = self.support_neural_points['coarse']
        n_points = len(sup['xyz'])
        n = self.args.matching.fine_num_3d_keypoints
        idx = torch.multinomial(sup['keypoint_score'].squeeze(1), n, replacement=n_points < n)
        return sup['xyz'][idx], sup['xyz_ndc'][idx], idx

    # ---- query (model.py:277-436) ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def query(self, data, xyz, support_featmaps=None, support_neural_points=None, direction=None, K=8, embed_a=None,
              target_proj_mat=None, _level=None):
        """Same contract as the reference.  `support_featmaps` / `support_neural_points` must be one of the two levels
        of the current frame (they are, at every reference call site); `_level` skips the identity lookup."""
        L = _lib.load()
        if self.support_neural_points is None:
            self.build_support_neural_points(data)
        level = _level
        if level is None:
            if support_neural_points is self.support_neural_points['fine']:
                level = 'fine'
            elif support_neural_points is None or support_neural_points is self.support_neural_points['coarse']:
                level = 'coarse'
            else:
                raise ValueError("query: support_neural_points must be self.support_neural_points['coarse'] or ['fine'] of the "
                                 "current frame (the kernels read the per-frame precomputes of that level); pass _level to "
                                 "name the level explicitly")
        sc, maps, sup = self._level_scene(data, level)
        pts = _lib.f32(xyz)
        N, dev = pts.shape[0], pts.device
        dirs = _lib.f32(direction[:, :3]) if direction is not None else None
        fagg = torch.empty(N, 128, device=dev)
        feat = torch.empty(N, 128, device=dev)
        wts = torch.empty(N, K, device=dev)
        mvf = torch.empty(N, maps.V, 195, device=dev)
        mvv = torch.empty(N, maps.V, 1, device=dev)
        idx = torch.empty(N, K, dtype=torch.int32, device=dev)
        nb = L.nlb_query_scratch_bytes(N, K)
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(L.nlb_query_points(ctypes.byref(sc), _lib.ptr(self.packed_weights()), self.n_samples, _lib.ptr(pts),
                                      _lib.ptr(dirs), N, K, _lib.ptr(fagg), _lib.ptr(feat), _lib.ptr(wts), _lib.ptr(mvf),
                                      _lib.ptr(mvv), None, _lib.ptr(idx), None, _lib.ptr(scratch), nb, _lib.stream()))
        return {'feature_agg': fagg, 'feature': feat.unsqueeze(1).expand(-1, K, -1), 'weights': wts,
                'multiview_feature': mvf, 'multiview_visibility': mvv, 'knn_idx': idx}

    def _descriptor(self, data, points, level, K):
        L = _lib.load()
        q = self.query(data, points, K=K, _level=level)
        sup = self.support_neural_points[level]
        if K == 1:
            nn_idx = q['knn_idx'][:, 0].long()
        else:
            _, i1 = self._frame['sup_' + level].index.query(points, 1)
            nn_idx = i1[:, 0]
        x = torch.cat([q['feature_agg'], sup['feature'][nn_idx]], 1).contiguous()
        out = torch.empty(x.shape[0], 192, device=x.device)
        _lib.check(L.nlb_descriptor_head(_lib.ptr(self.packed_weights()), self.n_samples, 0 if level == 'coarse' else 1,
                                         _lib.ptr(x), x.shape[0], _lib.ptr(out), _lib.stream()))
        return out

    @torch.no_grad()
    def query_coarse(self, data, points=None, embed_a=None):
        if self.support_neural_points is None:
            self.build_support_neural_points(data)
        if points is None:
            pts3d, pts3d_ndc, sample_idx = self.sample_points_3d()
            L = _lib.load()
            q = self.query(data, pts3d, K=8, _level='coarse')
            x = torch.cat([q['feature_agg'], self.support_neural_points['coarse']['feature'][sample_idx]], 1).contiguous()
            desc = torch.empty(x.shape[0], 192, device=x.device)
            _lib.check(L.nlb_descriptor_head(_lib.ptr(self.packed_weights()), self.n_samples, 0, _lib.ptr(x), x.shape[0],
                                             _lib.ptr(desc), _lib.stream()))
            return desc, pts3d, pts3d_ndc
        w2c_ref = data['topk_poses'][0].inverse()
        pts3d_ndc = (torch.matmul(w2c_ref[:3, :3], points.T) + w2c_ref[:3, 3:]).T
        return self._descriptor(data, points, 'coarse', 8), points, pts3d_ndc

    @torch.no_grad()
    def query_fine(self, data, points, embed_a=None):
        if self.support_neural_points is None:
            self.build_support_neural_points(data)
        return self._descriptor(data, points, 'fine', 1), None, None

    # ---- rendering (model.py:451-639) -------------------------------------------------------------------------------------------
    def sample_depths(self, N_samples, near, far):
        z_steps = torch.linspace(0, 1, N_samples, device=near.device)
        if not self.args.render.lindisp:
            return near * (1 - z_steps) + far * z_steps
        return 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)

    @torch.no_grad()
    def hierarchical_depths(self, data, rays, u=None):
        """model.py:486-496 on the device: (z_vals [R, N_samples + N_importance] sorted, depth_coarse [R], searchsorted
        indices [R, N_importance]).  `u` [R, N_importance] overrides the uniform draws (torch.rand in the reference)."""
        L = _lib.load()
        near, far = rays['depth_range']
        S0, NI = self.args.render.N_samples, self.args.render.N_importance
        sc, maps, sup = self._level_scene(data, 'fine', query_pose=data['pose'])
        px = _lib.f32(rays['pixel_coordinates'])
        dev, R = px.device, px.shape[0]
        # coords2rays of the query camera (depth_fusion.py:9-30): un-normalised directions K^-1 [u, v, 1] in the world frame
        w2c = rays['pose'].float().to(dev).inverse()[:3]
        rot = w2c[:, :3].t()
        trans = -rot @ w2c[:, 3:]
        cam = torch.inverse(rays['K'].float().to(dev)) @ torch.cat([px, torch.ones(R, 1, device=dev)], 1).t()
        dirs = ((rot @ cam + trans).t() - trans.t()).contiguous()
        center = (ctypes.c_float * 3)(*[float(v) for v in trans[:, 0].cpu()])
        zc, zr = _lib.f32(self.sample_depths(64, near, far)), _lib.f32(self.sample_depths(S0, near, far))
        u = torch.rand(R, NI, device=dev) if u is None else _lib.f32(u, dev)
        z = torch.empty(R, S0 + NI, device=dev)
        depth_coarse = torch.empty(R, device=dev)
        inds = torch.empty(R, NI, dtype=torch.int64, device=dev)
        _lib.check(L.nlb_hierarchical_depths(ctypes.byref(sc), _lib.ptr(self.packed_weights()), S0 + NI, center, _lib.ptr(dirs), R,
                                             _lib.ptr(zc), S0, _lib.ptr(zr), _lib.ptr(u), NI, _lib.ptr(z), _lib.ptr(depth_coarse),
                                             _lib.ptr(inds), _lib.stream()))
        return z, depth_coarse, inds

    @torch.no_grad()
    def render_rays(self, data, rays, _debug=False, _u=None, _feat_peers=None):
        """model.py:472-600.  `_feat_peers = (device pointers, row0)` (ray sharding, see distributed.FeatExchange): the rendered
        feature rows are also stored into every listed [R_total,192] buffer at rows row0.. by the ray kernel itself."""
        L = _lib.load()
        near, far = rays['depth_range']
        S = self.n_samples
        sc, maps, sup = self._level_scene(data, 'fine', query_pose=data['pose'])
        ro, rd = _lib.f32(rays['rays_o']), _lib.f32(rays['rays_d'])
        dev, R = ro.device, ro.shape[0]
        depth_coarse = None
        if self.args.render.N_importance > 0:
            z, depth_coarse, _ = self.hierarchical_depths(data, rays, _u)
        else:
            z = _lib.f32(self.sample_depths(S, near, far))
        out = {'rgb': torch.empty(R, 3, device=dev), 'depth': torch.empty(R, device=dev),
               'weights': torch.empty(R, S, device=dev), 'mask': torch.empty(R, dtype=torch.uint8, device=dev),
               'depth_uncertainty': torch.empty(R, device=dev)}
        feat = torch.empty(R, 192, device=dev) if self.args.render.render_feature else None
        dbg_fa = torch.empty(R * S, 128, device=dev) if _debug else None
        dbg_sig = torch.empty(R * S, device=dev) if _debug else None
        # equal chunks (a multiple of the SM count) instead of full chunks plus a small remainder
        cap = max(1, self.chunk_rays * 128 // max(S, 128))       # the scratch of a chunk scales with rays x samples
        n_chunks = max(1, -(-R // cap))
        chunk = min(cap, max(1, -(-(-(-R // n_chunks)) // 148) * 148))
        nb = L.nlb_render_scratch_bytes(chunk, S, maps.V)
        key = "scratch"
        if key not in self._frame or self._frame[key].numel() < nb:
            self._frame.pop(key, None)                            # one scratch buffer per model: release before growing
            self._frame[key] = torch.empty(nb, dtype=torch.uint8, device=dev)
        white = 1 if data.get('white_bkgd', self.args.render.white_bkgd) else 0
        if _feat_peers is not None:
            if _debug:
                raise ValueError("render_rays: _debug and _feat_peers are mutually exclusive")
            ptrs, row0 = _feat_peers
            arr = (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
            _lib.check(L.nlb_render_rays_gather(ctypes.byref(sc), _lib.ptr(self.packed_weights()), S, _lib.ptr(ro), _lib.ptr(rd),
                                                _lib.ptr(z), S if z.dim() == 2 else 0, R, white, chunk, _lib.ptr(out['rgb']),
                                                _lib.ptr(out['depth']), _lib.ptr(out['weights']), _lib.ptr(out['mask']),
                                                _lib.ptr(out['depth_uncertainty']), _lib.ptr(feat), _lib.ptr(self._frame[key]), nb,
                                                arr, len(ptrs), int(row0), _lib.stream()))
        else:
            _lib.check(L.nlb_render_rays(ctypes.byref(sc), _lib.ptr(self.packed_weights()), S, _lib.ptr(ro), _lib.ptr(rd),
                                         _lib.ptr(z), S if z.dim() == 2 else 0, R, white, chunk, _lib.ptr(out['rgb']), _lib.ptr(out['depth']),
                                         _lib.ptr(out['weights']), _lib.ptr(out['mask']), _lib.ptr(out['depth_uncertainty']),
                                         _lib.ptr(feat), _lib.ptr(dbg_fa), _lib.ptr(dbg_sig), _lib.ptr(self._frame[key]), nb,
                                         _lib.stream()))
        out['mask'] = out['mask'].bool()
        if feat is not None:
            out['feat'] = feat
        if depth_coarse is not None:
            out['depth_coarse'] = depth_coarse
        if _debug:
            out['feature_agg'], out['sigma'] = dbg_fa, dbg_sig.view(R, S)
        return out

    @torch.no_grad()
    def render_image(self, data):
        H, W, K, pose = data['H'], data['W'], data['K'], data['pose']
        rays_o, rays_d = get_rays(H, W, K, pose)
        vv, uu = torch.meshgrid(torch.linspace(0, H - 1, H, device=K.device), torch.linspace(0, W - 1, W, device=K.device),
                                indexing="ij")
        ray_batch = {'pixel_coordinates': torch.stack([uu.reshape(-1), vv.reshape(-1)], 1), 'K': K, 'pose': pose,
                     'H': H, 'W': W, 'rays_o': rays_o.reshape(-1, 3), 'rays_d': rays_d.reshape(-1, 3),
                     'depth_range': data['depth_range'][0]}
        ret = self.render_rays(data, ray_batch)  # chunking happens inside nlb_render_rays
        all_ret = {k: v.view(H, W, -1) for k, v in ret.items()}
        if 'target_mask' in data:
            all_ret['rgb'] = all_ret['rgb'] * data['target_mask'][:, :, None].float()
        return all_ret

    def points_2d_to_rays(self, pts2d, H, W, K, pose):
        x, y = pts2d[:, 0].long(), pts2d[:, 1].long()
        rays_o, rays_d = get_rays(H, W, K, pose)
        return {'pose': pose, 'K': K, 'H': H, 'W': W, 'pixel_coordinates': pts2d,
                'rays_o': rays_o[y, x], 'rays_d': rays_d[y, x]}

    def sample_rays(self, n_rays, H, W, K, pose, mask=None):
        u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="ij")
        pts2d = torch.stack([u.reshape(-1).float(), v.reshape(-1).float()], 1)
        if mask is not None:
            pts2d = pts2d[mask[pts2d[:, 1].long(), pts2d[:, 0].long()].bool().cpu()]
        idx = np.random.choice(len(pts2d), n_rays, replace=False)
        return self.points_2d_to_rays(pts2d[idx].to(K.device), H, W, K, pose)

    def compute_render_loss(self, data):
        raise NotImplementedError("training losses need autograd through the kernels (SURVEY.md section 8f rank 2)")


class _AttnParams(nn.Module):
    """Parameter container for ibrnet.py:69-90 MultiHeadAttention(4, 128, 32, 32)."""

    def __init__(self, d_model):
        super().__init__()
        self.w_qs = nn.Linear(d_model, d_model, bias=False)
        self.w_ks = nn.Linear(d_model, d_model, bias=False)
        self.w_vs = nn.Linear(d_model, d_model, bias=False)
        self.fc = nn.Linear(d_model, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
