"""Host mirror of the top-level drop-in surface: `NerfPoseEstimator.forward(batch) -> dict`
(nerf_loc/models/nerf_pose_estimator.py:33-583), inference mode.

The hot path behind it - per-frame scene setup, `query_coarse` / `query_fine`, `Matcher.forward`, `render_image` and the
absolute pose - runs on the CUDA library (conditional_nerf.py, matcher.py, pnp.py).  What sits in front of the hot path is
plain PyTorch here, as SURVEY.md section 8(f) rank 4 leaves it: the 2D backbone is torchvision's ResNet-50 + FPN assembled
exactly like COTR/backbone2d.py:67-124 (same sub-module names, so a reference checkpoint's `backbone2d.*` keys load), and the
appearance statistics / adaptation layers follow appearance_embedding.py:17-67.

Not built (raise): training mode (autograd through the kernels, section 8(f) rank 2) and `optimize_pose` (pose_optimizer.py).
"""
import traceback
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import pnp
from .conditional_nerf import ConditionalNeRF
from .matcher import Matcher, PositionEmbeddingSine


class Backbone2D(nn.Module):
    """COTR/backbone2d.py:67-124: ResNet-50 with frozen BatchNorm, taps after conv1 / layer1 / layer2, FPN (InstanceNorm) on
    the `layer*` taps.  Randomly initialised: the COTR checkpoint is not part of this repository."""

    def __init__(self, return_layers=('conv1', 'layer1', 'layer2'), use_fpn=True, fpn_dim=192):
        super().__init__()
        import torchvision
        from torchvision.models._utils import IntermediateLayerGetter
        from torchvision.ops import FeaturePyramidNetwork
        from torchvision.ops.misc import FrozenBatchNorm2d
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1), persistent=False)
        self.layer_to_channels = {'conv1': 64, 'layer1': 256, 'layer2': 512, 'layer3': 1024, 'layer4': 2048}
        self.layer_to_stride = {'conv1': 2, 'layer1': 4, 'layer2': 8, 'layer3': 16, 'layer4': 32}
        self.return_layers = list(return_layers)
        net = torchvision.models.resnet50(weights=None, replace_stride_with_dilation=[False, False, False],
                                          norm_layer=FrozenBatchNorm2d)
        self.body = IntermediateLayerGetter(net, return_layers={l: l for l in self.return_layers})
        self.use_fpn = use_fpn
        if use_fpn:
            self.fpn = FeaturePyramidNetwork([self.layer_to_channels[l] for l in self.return_layers if 'layer' in l], fpn_dim,
                                             norm_layer=nn.InstanceNorm2d)
            self.layer_to_channels.update({l: fpn_dim for l in self.return_layers if 'layer' in l})

    def forward(self, x):
        y = self.body((x - self.mean) / self.std)
        if self.use_fpn:
            y.update(self.fpn(OrderedDict((l, y[l]) for l in self.return_layers if 'layer' in l)))
        return y


class AppearanceEmbedding(nn.Module):  # appearance_embedding.py:17-36: per-image mean | std of the conv1 features
    def __init__(self, args):
        super().__init__()
        self.dim = args.appearance_emb_dim

    def forward(self, imgs, x):
        f = x['conv1'].flatten(2)
        std, mean = torch.std_mean(f, dim=2)
        return torch.cat([mean, std], 1)


class AppearanceAdaptLayer(nn.Module):  # appearance_embedding.py:38-67: channel-wise affine predicted from the embedding gap
    def __init__(self, args, input_dim, is_rgb=False):
        super().__init__()
        self.input_dim, self.is_rgb = input_dim, is_rgb
        self.mlp = nn.Sequential(nn.Linear(args.appearance_emb_dim, 64), nn.LeakyReLU(inplace=True), nn.Linear(64, 64),
                                 nn.LeakyReLU(inplace=True), nn.Linear(64, input_dim * 2))

    def forward(self, x, embedding, target_embedding):
        a, b = self.mlp(target_embedding - embedding).split([self.input_dim, self.input_dim], dim=-1)
        y = a[:, None, None, :] * x + b[:, None, None, :]
        return y.clip(0., 1.) if self.is_rgb else y


def camera_project(p3d, K):  # models/utils.py:12-21
    uvz = K @ p3d.t()
    return uvz[0] / uvz[2], uvz[1] / uvz[2], uvz[2]


def _cfg(args, name, default):
    return getattr(args, name, default)


class NerfPoseEstimator(nn.Module):
    def __init__(self, args, dataset=None, backbone2d=None):
        super().__init__()
        self.args = args
        self.dataset = dataset
        self.hidden_dim = hidden_dim = args.matcher_hidden_dim
        self.backbone2d = backbone2d if backbone2d is not None else Backbone2D(
            use_fpn=_cfg(args, 'backbone2d_use_fpn', True), fpn_dim=args.backbone2d_fpn_dim)
        self.backbone2d_coarse_layer_name = _cfg(args, 'backbone2d_coarse_layer_name', 'layer2')
        self.backbone2d_fine_layer_name = _cfg(args, 'backbone2d_fine_layer_name', 'layer1')
        if args.encode_appearance:
            self.embedding_a = AppearanceEmbedding(args)
            self.adapt_appearance_coarse = AppearanceAdaptLayer(args, args.backbone2d_fpn_dim)
            self.adapt_appearance_fine = AppearanceAdaptLayer(args, args.backbone2d_fpn_dim)
            if _cfg(args, 'train_nerf', True):
                self.adapt_appearance_rgb = AppearanceAdaptLayer(args, 3, is_rgb=True)
        scale = getattr(dataset, 'scale_factor', 1.0)
        self.coarse_matching_depth_thresh = _cfg(args.matching, 'coarse_matching_depth_thresh', 2.) * scale
        cc = self.backbone2d.layer_to_channels[self.backbone2d_coarse_layer_name]
        cf = self.backbone2d.layer_to_channels[self.backbone2d_fine_layer_name]
        self.proj_layer_2d = nn.Linear(cc, hidden_dim)
        self.pos_emd_2d_fn = PositionEmbeddingSine(hidden_dim // 2, normalize=True, sine_type='lin_sine')
        self._pe3_freqs = 2.0 ** torch.linspace(0.0, hidden_dim // 6 - 1, steps=hidden_dim // 6)
        self.matcher = Matcher(args, hidden_dim, cc, cf, fine_matching=True)
        if _cfg(args, 'cascade_matching', False):
            self.matcher_fine = Matcher(args, hidden_dim, cc, cf, fine_matching=True)
        if _cfg(args, 'simple_3d_model', False):
            raise NotImplementedError("ConditionalNeRFSimple is not part of the hot path (SURVEY.md section 8)")
        self.model_3d = ConditionalNeRF(args)
        if _cfg(args, 'optimize_pose', False):
            raise NotImplementedError("optimize_pose needs gradients through render_rays (SURVEY.md section 8f rank 2)")

    def pos_emd_3d_fn(self, x):
        """get_embedder(hidden_dim // 6, 0, include_input=False) (conditional_nerf/utils.py:5-53): [sin(2^i x), cos(2^i x)]_i"""
        f = self._pe3_freqs.to(x.device)
        return torch.cat([fn(x * fr) for fr in f for fn in (torch.sin, torch.cos)], -1)

    # ---- nerf_pose_estimator.py:94-124 -------------------------------------------------------------------------------------
    def extract_2d(self, imgs):
        pyr = self.backbone2d(imgs)
        feat_coarse = pyr[self.backbone2d_coarse_layer_name].permute(0, 2, 3, 1)
        feat_fine = pyr[self.backbone2d_fine_layer_name].permute(0, 2, 3, 1)
        return {'feat_rgb': imgs.permute(0, 2, 3, 1), 'feat_pyramid': pyr, 'feat_fine': feat_fine, 'feat_coarse': feat_coarse,
                'pos_emb_coarse': self.pos_emd_2d_fn(feat_coarse[..., 0]),
                'stride_coarse': self.backbone2d.layer_to_stride[self.backbone2d_coarse_layer_name],
                'stride_fine': self.backbone2d.layer_to_stride[self.backbone2d_fine_layer_name]}

    # ---- nerf_pose_estimator.py:126-180 (inference branch: ground-truth depth lookup only) -----------------------------------
    def build_3d_2d_pairs(self, img, pts3d, H, W, K, pose, feat_pyramid_2d=None, thr=0.01, stride=1, depth_map=None, data=None):
        with torch.no_grad():
            hom = torch.cat([pts3d, torch.ones_like(pts3d[:, :1])], 1)
            cam = (pose.inverse() @ hom.t()).t()
            u, v, z = camera_project(cam[:, :3], K)
            valid = (u >= 0) & (v >= 0) & (u < W) & (v < H) & (z > 0)
            uv = torch.stack([u, v], 1)
            depth = depth_map[v[valid].long(), u[valid].long()]
            depth_ok = torch.zeros_like(valid)
            depth_ok[valid] = (depth - z[valid]).abs() < thr
            pos = valid & depth_ok
            if pos.sum() < 4:
                pos = valid
            gy, gx = torch.meshgrid(torch.arange(H // stride), torch.arange(W // stride), indexing='ij')
            grid = torch.stack([gx, gy], -1).view(-1, 2).to(pts3d.device)
            idx2d = (uv[pos] / stride).long()
            idx2d = idx2d[:, 0] + idx2d[:, 1] * (W // stride)
            pairs = torch.stack([pos.nonzero(as_tuple=True)[0], idx2d])
        return pts3d, grid, uv / stride, pairs

    def select_3d_keypoints(self, pts3d, topk_poses, K, H, W):  # :183-195
        vis = torch.zeros(len(pts3d), dtype=torch.bool, device=pts3d.device)
        hom = torch.cat([pts3d, torch.ones_like(pts3d[:, :1])], 1)
        for pose in topk_poses:
            u, v, z = camera_project((pose.inverse() @ hom.t()).t()[:, :3], K)
            vis |= (u >= 0) & (v >= 0) & (u < W) & (v < H) & (z > 0)
        return torch.where(vis)[0]

    def build_support_set(self, batch):  # :197-217, inference branch
        n = _cfg(self.args, 'n_views_test', 10)
        d = batch['topk_depths'][0]
        dg = batch['topk_depths_gt'][0] if 'topk_depths_gt' in batch else d
        return batch['topk_images'][0][:n], d[:n], dg[:n], batch['topk_poses'][0][:n], batch['topk_Ks'][0][:n]

    def appearance_adaptation(self, batch, data, outputs, feat_pyramid, feat_pyramid_src):  # :219-237
        embedding_a = None
        if self.args.encode_appearance:
            embedding_a = self.embedding_a(batch['image'], feat_pyramid)
            src = self.embedding_a(data['topk_images'], feat_pyramid_src)
            if _cfg(self.args, 'train_nerf', True):
                data['topk_images'] = self.adapt_appearance_rgb(data['topk_images'].permute(0, 2, 3, 1), src, embedding_a).permute(0, 3, 1, 2)
                outputs['topk_images_adapted'] = data['topk_images']
            data['feat_coarse_src'] = self.adapt_appearance_coarse(data['feat_coarse_src'], src, embedding_a)
            data['feat_fine_src'] = self.adapt_appearance_fine(data['feat_fine_src'], src, embedding_a)
        data['embedding_a'] = embedding_a
        return data

    # ---- nerf_pose_estimator.py:239-405 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, batch):
        if self.training:
            raise NotImplementedError("NerfPoseEstimator: training mode needs autograd through the CUDA kernels "
                                      "(SURVEY.md section 8f rank 2); call .eval()")
        assert batch['image'].shape[0] == 1
        pts3d_all = batch['points3d'][..., :3][0]
        topk_images, topk_depths, topk_depths_gt, topk_poses, topk_Ks = self.build_support_set(batch)
        out2d = self.extract_2d(batch['image'])
        out2d_src = self.extract_2d(topk_images)
        H, W = batch['image'][0].shape[-2:]
        data = {'scene': batch['scene'][0], 'filename': batch['filename'][0], 'img': batch['image'][0], 'depth': batch['depth'][0],
                'K': batch['K'][0], 'pose': batch['pose'][0], 'H': H, 'W': W, 'pts3d_all': pts3d_all,
                'depth_range': torch.stack([batch['near'], batch['far']], dim=1), 'topk_images': topk_images,
                'topk_depths': topk_depths, 'topk_depths_gt': topk_depths_gt, 'topk_poses': topk_poses, 'topk_Ks': topk_Ks,
                'feat_coarse_src': out2d_src['feat_coarse'], 'feat_fine_src': out2d_src['feat_fine']}
        for k in ('white_bkgd', 'target_mask', 'sample_coords'):
            if k in batch:
                data[k] = batch[k][0]
        data.update(out2d)
        outputs = {'loss': 0}
        data = self.appearance_adaptation(batch, data, outputs, out2d['feat_pyramid'], out2d_src['feat_pyramid'])
        data['feat_coarse_src'] = data['feat_coarse_src'].contiguous()
        data['feat_fine_src'] = data['feat_fine_src'].contiguous()
        # per-frame caches: force a rebuild of the support neural points (reference :289-290)
        self.model_3d.support_neural_points = None
        self.model_3d.multiview_aggregator.vis_featmaps = None

        if _cfg(self.args, 'train_pose', True):
            target_points = None
            if _cfg(self.args, 'keypoints_3d_source', 'depth') == 'sfm':
                target_points = data['pts3d_all']
                n = self.args.matching.fine_num_3d_keypoints
                if len(target_points) > n:
                    target_points = target_points[torch.as_tensor(np.random.choice(len(target_points), n, replace=False)).long()]
            desc_3d, pts3d, pts3d_ndc = self.model_3d.query_coarse(data=data, points=target_points, embed_a=data['embedding_a'])
            data.update({'pts3d': pts3d, 'pts3d_ndc': pts3d_ndc, 'desc_3d': desc_3d})
            outputs.update(self.estimate(data, self.matcher))
            if _cfg(self.args, 'cascade_matching', False):
                T_init = torch.as_tensor(outputs['T']).float().to(desc_3d.device)
                sel = self.select_3d_keypoints(pts3d, [T_init], data['K'], H, W)
                if len(sel) > 0:
                    data.update({'pts3d': pts3d[sel], 'pts3d_ndc': pts3d_ndc[sel], 'desc_3d': desc_3d[sel]})
                    outputs['T'] = self.estimate(data, self.matcher_fine)['T']

        if batch.get('render_image', False):
            ret = self.model_3d.render_image(data)
            outputs['rendered_image'], outputs['rendered_depth'] = ret['rgb'], ret['depth']
            if 'depth_coarse' in ret:
                outputs['rendered_depth_coarse'] = ret['depth_coarse']
            if 'feat' in ret:
                outputs['rendered_feat'] = ret['feat']
                outputs['rendered_feat_gt'] = nn.functional.interpolate(
                    data['feat_pyramid']['layer1'], size=(H, W), mode='bilinear', align_corners=False).permute(0, 2, 3, 1)[0]
        return outputs

    # ---- nerf_pose_estimator.py:407-555 ------------------------------------------------------------------------------------
    def estimate(self, data, matcher, need_pose=False):
        fine_matching = matcher.fine_matching
        K, H, W = data['K'], data['H'], data['W']
        desc_map, pos_map = data['feat_coarse'][0], data['pos_emb_coarse'][0]
        pts3d, pts2d, proj_gt, pos_pairs = self.build_3d_2d_pairs(
            data['img'], data['pts3d'], H, W, K, data['pose'], feat_pyramid_2d=data['feat_pyramid'],
            thr=self.coarse_matching_depth_thresh, stride=data['stride_coarse'], depth_map=data['depth'], data=data)
        y, x = pts2d[:, 1].long(), pts2d[:, 0].long()
        desc_2d = self.proj_layer_2d(desc_map[y, x])
        pos_emd_2d = pos_map[y, x]
        if fine_matching:
            data['desc_3d_fine'] = self.model_3d.query_fine(data=data, points=pts3d, embed_a=data['embedding_a'])[0]
        pos_emd_3d = self.pos_emd_3d_fn(data['pts3d_ndc'])
        pts2d = (pts2d * data['stride_coarse'] / data['stride_fine']).float()
        data.update({'kps3d': pts3d, 'kps2d': pts2d, 'desc_2d_coarse': desc_2d, 'pos_emd_3d': pos_emd_3d, 'pos_emd_2d': pos_emd_2d})
        match_res = matcher(data)
        if match_res is None or len(match_res['i_ids']) == 0:
            return {'score_matrix': match_res['score_matrix'] if match_res is not None else torch.zeros(len(pts3d), len(pts2d), device=pts3d.device),
                    'pairs_gt': pos_pairs, 'pairs': [torch.tensor([]), torch.tensor([])], 'mkps2d': [], 'mkps3d': [], 'T': np.eye(4)}
        mkps3d = match_res['mkps3d']
        mkps2d = (match_res['mkps2d_f'] if fine_matching else match_res['mkps2d_c']) * data['stride_fine']
        data['pairs'] = match_res['pairs']
        outputs = {'score_matrix': match_res['score_matrix'], 'pairs_gt': pos_pairs, 'pairs': match_res['pairs'],
                   'mkps2d': mkps2d, 'mkps3d': mkps3d}
        try:
            thr = _cfg(self.args, 'ransac_thresh', 8) * (1 if fine_matching else 2)
            ret = self.estimate_pose(mkps2d.detach(), mkps3d.detach(), K, W, H, ransac_thresh=thr)
            T = ret[0] if ret is not None else np.eye(4)   # the reference's tuple-unpack of None lands in the same except branch
        except Exception:
            traceback.print_exc()
            T = np.eye(4)
        outputs['T'] = T
        return outputs

    def estimate_pose(self, matched_kps_2d, matched_kps_3d, K, width, height, ransac_thresh=48):  # :557-583
        return pnp.estimate_pose(matched_kps_2d, matched_kps_3d, K, width, height, ransac_thresh=ransac_thresh)
