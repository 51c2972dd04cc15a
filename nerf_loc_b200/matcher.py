"""Drop-in `Matcher` (nerf_loc/models/matcher.py:10-131) for inference.

The pairwise scoring MLPs (S2DMatching, FineMatching), the mutual-nearest rule and the fine-window gather run in
libnerfloc_b200.so.  The two SelfCrossTransformers are standard attention blocks and stay on torch
(nn.MultiheadAttention -> SDPA) for now, as SURVEY.md section 8(a15) allows.  Parameter names equal the reference's, so
a reference checkpoint loads unchanged.  Training-mode branches (GT pairs, losses) need autograd through the kernels and
raise.
"""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class _SelfLayer(nn.Module):  # COTR/transformer.py:171-206
    def __init__(self, d, nhead, ffn, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1, self.dropout, self.linear2 = nn.Linear(d, ffn), nn.Dropout(dropout), nn.Linear(ffn, d)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.dropout1, self.dropout2 = nn.Dropout(dropout), nn.Dropout(dropout)

    def forward(self, src, pos):
        qk = src + pos
        src = self.norm1(src + self.dropout1(self.self_attn(qk, qk, src, need_weights=False)[0]))
        return self.norm2(src + self.dropout2(self.linear2(self.dropout(F.relu(self.linear1(src))))))


class _CrossLayer(nn.Module):  # COTR/transformer.py:209-250
    def __init__(self, d, nhead, ffn, dropout):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1, self.dropout, self.linear2 = nn.Linear(d, ffn), nn.Dropout(dropout), nn.Linear(ffn, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.dropout1, self.dropout2, self.dropout3 = nn.Dropout(dropout), nn.Dropout(dropout), nn.Dropout(dropout)

    def forward(self, tgt, memory, query_pos, pos):
        a = self.multihead_attn(tgt + query_pos, memory + pos, memory, need_weights=False)[0]
        tgt = self.norm2(tgt + self.dropout2(a))
        return self.norm3(tgt + self.dropout3(self.linear2(self.dropout(F.relu(self.linear1(tgt))))))


class SelfCrossTransformer(nn.Module):  # COTR/transformer.py:17-63
    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", return_intermediate_dec=False):
        super().__init__()
        self.self_attn_layer0 = _SelfLayer(d_model, nhead, dim_feedforward, dropout)
        self.self_attn_layer1 = _SelfLayer(d_model, nhead, dim_feedforward, dropout)
        self.cross_attn_layer0 = _CrossLayer(d_model, nhead, dim_feedforward, dropout)
        self.cross_attn_layer1 = _CrossLayer(d_model, nhead, dim_feedforward, dropout)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.d_model, self.nhead = d_model, nhead

    def forward(self, v0, pos_embed0, v1, pos_embed1):
        v0, v1, p0, p1 = (t.transpose(0, 1) for t in (v0, v1, pos_embed0, pos_embed1))
        v0 = self.self_attn_layer0(v0, p0)
        v1 = self.self_attn_layer1(v1, p1)
        v0 = self.cross_attn_layer0(v0, v1, p0, p1)
        v1 = self.cross_attn_layer1(v1, v0, p1, p0)
        return v0.transpose(0, 1).contiguous(), v1.transpose(0, 1).contiguous()


class PositionEmbeddingSine(nn.Module):
    """COTR/position_encoding.py:32-80 ('lin_sine'): [B,H,W] -> [B,H,W,2*num_pos_feats]."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None, sine_type='lin_sine'):
        super().__init__()
        if sine_type != 'lin_sine':
            raise NotImplementedError(sine_type)
        self.bases = [i + 1 for i in range(num_pos_feats // 2)]

    @torch.no_grad()
    def forward(self, x):
        ones = torch.ones_like(x)
        y = ones.cumsum(1, dtype=torch.float32)
        xx = ones.cumsum(2, dtype=torch.float32)
        y = (y - 0.5) / (y[:, -1:, :] + 1e-6)
        xx = (xx - 0.5) / (xx[:, :, -1:] + 1e-6)
        p = torch.stack([xx, y], dim=-1)
        return torch.cat([torch.sin(i * math.pi * p) for i in self.bases] + [torch.cos(i * math.pi * p) for i in self.bases], -1)


def _pair_mlp(feat_dim):
    return nn.Sequential(nn.Linear(feat_dim, 128), nn.ReLU(inplace=True), nn.Linear(128, 128), nn.ReLU(inplace=True),
                         nn.Linear(128, 1))


class S2DMatching(nn.Module):
    """matching/sparse_to_dense.py:80-151; scores and the mutual-nearest rule run on the device kernels."""

    def __init__(self, feat_dim, thr=0.1):
        super().__init__()
        self.mlps = _pair_mlp(feat_dim)
        self.thr = thr
        self._owner = None

    def forward(self, desc0, desc1, data):
        assert (desc0.shape[0] > 0) and (desc1.shape[0] > 0)
        if self.training:
            raise NotImplementedError("S2DMatching training loss needs autograd through the kernels")
        score, i_ids, j_ids = self._owner().s2d(desc0, desc1, self.thr)
        data.update({'i_ids': i_ids, 'j_ids': j_ids, 'score_matrix': score})
        return data


class FinePreprocess(nn.Module):
    """matching/fine_matching.py:9-76 (fine_concat_coarse_feat=False)."""

    def __init__(self, config):
        super().__init__()
        if config['fine_concat_coarse_feat']:
            raise NotImplementedError("fine_concat_coarse_feat=True is not used by the reference Matcher")
        self.W = config['fine_window_size']
        self.out_channels = config['out_channels']
        self.proj = nn.Linear(config['in_channels_fine'], config['out_channels'], bias=True)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.kaiming_normal_(p, mode="fan_out", nonlinearity="relu")
        self._owner = None

    def forward(self, feat_f1, feat_c1, data):
        """feat_f1 [1,C,h,w] (as the reference passes it) -> [M,49,192] for the matched cells data['j_ids']."""
        if len(data['j_ids']) == 0:
            return torch.empty(0, self.W ** 2, self.out_channels, device=feat_f1.device)
        stride = data['stride_coarse'] // data['stride_fine']
        return self._owner().fine_windows(feat_f1.permute(0, 2, 3, 1)[0], data['j_ids'], stride, feat_c1.shape[-1])


class FineMatching(nn.Module):
    """matching/fine_matching.py:79-153."""

    def __init__(self, config):
        super().__init__()
        self.correct_thr = config['correct_thr']
        self.loss_type = config['loss_type']
        self.mlps = _pair_mlp(config['feat_dim'])
        self._owner = None

    def forward(self, feat_f0, feat_f1, data):
        M, WW, C = feat_f1.shape
        if M == 0:
            assert self.training is False
            data.update({'expec_f': torch.empty(0, 3, device=feat_f0.device), 'mkps2d_f': data['mkps2d_c']})
            return data
        if self.training:
            raise NotImplementedError("FineMatching training loss needs autograd through the kernels")
        if WW != 49 or C != 192:
            raise RuntimeError("nerfloc_b200 fine matcher is built for 7x7 windows of 192-d descriptors")
        expec, mk = self._owner().fine_match(feat_f0, feat_f1, data['mkps2d_c'])
        data.update({'expec_f': expec, 'mkps2d_f': mk})
        return data


class Matcher(nn.Module):
    def __init__(self, args, hidden_dim, in_channels_coarse, in_channels_fine, fine_matching=True):
        super().__init__()
        import weakref
        if hidden_dim != 192:
            raise RuntimeError("nerfloc_b200 matcher kernels are built for matcher_hidden_dim=192")
        self.coarse_transformer = SelfCrossTransformer(d_model=hidden_dim, dropout=0.1, nhead=8, dim_feedforward=512)
        self.coarse_matcher = S2DMatching(hidden_dim, thr=0.2)
        self.fine_matching = fine_matching
        self.in_channels_fine = in_channels_fine
        if fine_matching:
            self.pos_emd_2d_fn = PositionEmbeddingSine(hidden_dim // 2, normalize=True, sine_type='lin_sine')
            self.fine_window_size = 7
            self.fine_preprocess = FinePreprocess({'fine_concat_coarse_feat': False, 'fine_window_size': 7,
                                                   'in_channels_coarse': in_channels_coarse,
                                                   'in_channels_fine': in_channels_fine, 'out_channels': hidden_dim})
            self.fine_transformer = SelfCrossTransformer(d_model=hidden_dim, dropout=0.1, nhead=8, dim_feedforward=128)
            self.fine_matcher = FineMatching({'feat_dim': hidden_dim, 'correct_thr': 1.0,
                                              'loss_type': args.fine_matching_loss_type})
        else:
            raise NotImplementedError("Matcher(fine_matching=False) is not used by the reference estimator")
        ref = weakref.ref(self)
        self.coarse_matcher._owner = self.fine_preprocess._owner = self.fine_matcher._owner = ref
        self._packed, self._packed_key = None, None

    # ---- device kernels ------------------------------------------------------------------------------------------------
    def packed_weights(self):
        L = _lib.load()
        ts = [p for m in (self.coarse_matcher.mlps, self.fine_matcher.mlps) for i in (0, 2, 4)
              for p in (m[i].weight, m[i].bias)] + [self.fine_preprocess.proj.weight, self.fine_preprocess.proj.bias]
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if self._packed is None or key != self._packed_key:
            keep = [_lib.f32(t) for t in ts]
            arr = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            n = L.nlb_match_weights_floats(self.in_channels_fine)
            packed = torch.empty(n, dtype=torch.float32, device=keep[0].device)
            _lib.check(L.nlb_match_pack_weights(arr, len(keep), self.in_channels_fine, _lib.ptr(packed), n, _lib.stream()))
            torch.cuda.current_stream().synchronize()
            self._packed, self._packed_key = packed, key
        return self._packed

    @torch.no_grad()
    def s2d(self, desc0, desc1, thr):
        L = _lib.load()
        d0, d1 = _lib.f32(desc0), _lib.f32(desc1)
        N, M = d0.shape[0], d1.shape[0]
        if d0.shape[1] != 192 or d1.shape[1] != 192:
            raise RuntimeError("S2D kernel is built for 192-d descriptors")
        dev = d0.device
        pk = self.packed_weights()
        score = torch.empty(N, M, device=dev)
        _lib.check(L.nlb_s2d_scores(_lib.ptr(pk), self.in_channels_fine, _lib.ptr(d0), _lib.ptr(d1), N, M,
                                    _lib.ptr(score), _lib.stream()))
        i_ids = torch.empty(N, dtype=torch.int64, device=dev)
        j_ids = torch.empty(N, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        nb = L.nlb_mutual_scratch_bytes(N, M)
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(L.nlb_mutual_matches(_lib.ptr(score), N, M, float(thr), _lib.ptr(i_ids), _lib.ptr(j_ids),
                                        _lib.ptr(cnt), _lib.ptr(scratch), nb, _lib.stream()))
        n = int(cnt.item())  # the caller indexes with the ids right away: one D2H of 4 bytes
        return score, i_ids[:n], j_ids[:n]

    @torch.no_grad()
    def fine_windows(self, feat_fine_hwc, j_ids, stride, coarse_w):
        L = _lib.load()
        f = _lib.f32(feat_fine_hwc)
        h, w, C = f.shape
        j = j_ids.to(torch.int64).contiguous()
        out = torch.empty(j.shape[0], 49, 192, device=f.device)
        _lib.check(L.nlb_fine_windows(_lib.ptr(self.packed_weights()), C, _lib.ptr(f), h, w, int(stride), int(coarse_w),
                                      _lib.ptr(j), j.shape[0], _lib.ptr(out), _lib.stream()))
        return out

    @torch.no_grad()
    def fine_match(self, f0, f1, mkps2d_c):
        L = _lib.load()
        f0, f1, mk = _lib.f32(f0), _lib.f32(f1), _lib.f32(mkps2d_c)
        M = f0.shape[0]
        expec = torch.empty(M, 3, device=f0.device)
        out = torch.empty(M, 2, device=f0.device)
        _lib.check(L.nlb_fine_match(_lib.ptr(self.packed_weights()), self.in_channels_fine, _lib.ptr(f0), _lib.ptr(f1), M,
                                    _lib.ptr(mk), _lib.ptr(expec), _lib.ptr(out), _lib.stream()))
        return expec, out

    # ---- matcher.py:63-131 -------------------------------------------------------------------------------------------------
    def forward(self, data):
        if self.training:
            raise NotImplementedError("Matcher training mode needs autograd through the kernels (SURVEY 8f rank 2)")
        d3, d2 = self.coarse_transformer(data['desc_3d'][None], data['pos_emd_3d'][None],
                                         data['desc_2d_coarse'][None], data['pos_emd_2d'][None])
        data = self.coarse_matcher(d3[0], d2[0], data)
        i_ids, j_ids = data['i_ids'], data['j_ids']
        data['b_ids'] = torch.zeros_like(i_ids)
        data.update({'mkps3d': data['kps3d'][i_ids], 'mkps2d_c': data['kps2d'][j_ids], 'pairs': [i_ids, j_ids]})
        M = len(i_ids)
        if M == 0:
            data.update({'expec_f': torch.empty(0, 3, device=data['mkps2d_c'].device), 'mkps2d_f': data['mkps2d_c']})
            return data
        feat_fine = data['feat_fine'].permute(0, 3, 1, 2)
        feat_coarse = data['feat_coarse'].permute(0, 3, 1, 2)
        m3 = data['desc_3d_fine'][i_ids][:, None, :]
        p3 = data['pos_emd_3d'][i_ids][:, None, :]
        wins = self.fine_preprocess(feat_fine, feat_coarse, data)
        W = self.fine_window_size
        pe = self.pos_emd_2d_fn(wins[..., 0].view(M, W, W)).view(M, W * W, -1)
        m3, wins = self.fine_transformer(m3, p3, wins, pe)
        return self.fine_matcher(m3[:, 0, :], wins, data)
