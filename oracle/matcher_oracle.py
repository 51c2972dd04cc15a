"""CPU oracle for the NeRF-Loc 3D<->2D coarse-to-fine matcher (TEST INFRASTRUCTURE).

From-scratch fp32 torch restatement of `Matcher.forward` in eval mode
(nerf_loc/models/matcher.py:63-131) and the modules below it.  Same rules as
`oracle/nerfloc_oracle.py`: only tests / smoke / the bench's CPU legs import it.

Parity status: PINNED against the reference itself
(`tests/test_oracle_vs_reference.py`) and `tests/golden/matcher_small.npz`.
Citations are `path:line` below /root/reference/nerf_loc/models/.
"""
import math

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def torch_mha(sd, pfx, q, k, v, nhead=8):
    """torch.nn.MultiheadAttention forward (no masks, no dropout); inputs [L,B,E] (COTR/transformer.py:176,216)."""
    Lq, B, E = q.shape
    Lk = k.shape[0]
    hd = E // nhead
    Wi, bi = sd[pfx + ".in_proj_weight"], sd[pfx + ".in_proj_bias"]
    qp = F.linear(q, Wi[:E], bi[:E])
    kp = F.linear(k, Wi[E:2 * E], bi[E:2 * E])
    vp = F.linear(v, Wi[2 * E:], bi[2 * E:])
    qp = qp.reshape(Lq, B * nhead, hd).transpose(0, 1)
    kp = kp.reshape(Lk, B * nhead, hd).transpose(0, 1)
    vp = vp.reshape(Lk, B * nhead, hd).transpose(0, 1)
    att = torch.softmax(torch.bmm(qp * (1.0 / math.sqrt(hd)), kp.transpose(1, 2)), dim=-1)
    o = torch.bmm(att, vp).transpose(0, 1).reshape(Lq, B, E)
    return _lin(sd, pfx + ".out_proj", o)


def self_layer(sd, pfx, src, pos):
    """COTR/transformer.py:171-206 (post-norm encoder layer, q=k=src+pos, v=src)."""
    qk = src + pos
    src = _ln(sd, pfx + ".norm1", src + torch_mha(sd, pfx + ".self_attn", qk, qk, src))
    ff = _lin(sd, pfx + ".linear2", F.relu(_lin(sd, pfx + ".linear1", src)))
    return _ln(sd, pfx + ".norm2", src + ff)


def cross_layer(sd, pfx, tgt, mem, query_pos, pos):
    """COTR/transformer.py:209-250 (cross attention only; norm1 is unused)."""
    a = torch_mha(sd, pfx + ".multihead_attn", tgt + query_pos, mem + pos, mem)
    tgt = _ln(sd, pfx + ".norm2", tgt + a)
    ff = _lin(sd, pfx + ".linear2", F.relu(_lin(sd, pfx + ".linear1", tgt)))
    return _ln(sd, pfx + ".norm3", tgt + ff)


def self_cross_transformer(sd, pfx, v0, pe0, v1, pe1):
    """COTR/transformer.py:43-63; inputs [B,N,C]."""
    v0, v1, pe0, pe1 = (t.transpose(0, 1) for t in (v0, v1, pe0, pe1))
    v0 = self_layer(sd, pfx + ".self_attn_layer0", v0, pe0)
    v1 = self_layer(sd, pfx + ".self_attn_layer1", v1, pe1)
    v0 = cross_layer(sd, pfx + ".cross_attn_layer0", v0, v1, pe0, pe1)
    v1 = cross_layer(sd, pfx + ".cross_attn_layer1", v1, v0, pe1, pe0)
    return v0.transpose(0, 1).contiguous(), v1.transpose(0, 1).contiguous()


def pair_mlp(sd, pfx, x):
    """192->128->128->1 with ReLU (matching/sparse_to_dense.py:83-89, fine_matching.py:101-107)."""
    x = F.relu(_lin(sd, pfx + ".mlps.0", x))
    x = F.relu(_lin(sd, pfx + ".mlps.2", x))
    return _lin(sd, pfx + ".mlps.4", x).squeeze(-1)


def s2d_scores(sd, pfx, desc0, desc1, chunk=256):
    """matching/sparse_to_dense.py:125-127: sigmoid(MLP(a_n * b_m)) for every pair -> [N,M] (chunked over N)."""
    out = []
    for s in range(0, desc0.shape[0], chunk):
        x = desc0[s:s + chunk, None, :] * desc1[None, :, :]
        out.append(torch.sigmoid(pair_mlp(sd, pfx, x)))
    return torch.cat(out)


def mutual_matches(score, thr=0.2):
    """matching/sparse_to_dense.py:136-142: > thr AND row max AND column max (exact float equality)."""
    mask = score > thr
    mask = mask * (score == score.max(dim=1, keepdim=True)[0]) * (score == score.max(dim=0, keepdim=True)[0])
    mv, all_j = mask.max(dim=1)
    i_ids = torch.where(mv)[0]
    return i_ids, all_j[i_ids]


def pos_embed_2d(x_like):
    """COTR/position_encoding.py:32-80 with num_pos_feats=96, lin_sine: input [B,H,W] -> [B,H,W,192]."""
    ones = torch.ones_like(x_like)
    y = ones.cumsum(1, dtype=torch.float32)
    x = ones.cumsum(2, dtype=torch.float32)
    y = (y - 0.5) / (y[:, -1:, :] + 1e-6)
    x = (x - 0.5) / (x[:, :, -1:] + 1e-6)
    p = torch.stack([x, y], -1)
    bases = [i + 1 for i in range(48)]
    return torch.cat([torch.sin(i * math.pi * p) for i in bases] + [torch.cos(i * math.pi * p) for i in bases], -1)


def fine_windows(feat_fine_chw, j_ids, stride, win=7):
    """matching/fine_matching.py:53-57: F.unfold 7x7 windows (stride = stride_c/stride_f, pad 3), pick cells."""
    C = feat_fine_chw.shape[1]
    u = F.unfold(feat_fine_chw, kernel_size=(win, win), stride=stride, padding=win // 2)  # [1, C*49, L]
    u = u.view(1, C, win * win, -1).permute(0, 3, 2, 1)  # n l ww c
    return u[0, j_ids]


def fine_match(sd, pfx, f0, f1, mkps2d_c, win=7):
    """matching/fine_matching.py:109-153 -> expec_f [M,3], mkps2d_f [M,2]."""
    M, WW, C = f1.shape
    sim = pair_mlp(sd, pfx, f0[:, None, :] * f1)
    heat = torch.softmax(sim * (1.0 / C ** 0.5), dim=1)
    lin = (torch.linspace(0, win - 1, win) / (win - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    grid = torch.stack([gx, gy], -1).reshape(1, -1, 2)
    coords = torch.stack([(heat * grid[..., 0]).sum(1), (heat * grid[..., 1]).sum(1)], -1)
    var = torch.sum(grid ** 2 * heat.view(-1, WW, 1), dim=1) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    return torch.cat([coords, std.unsqueeze(1)], -1), mkps2d_c + coords * (win // 2)


def matcher_forward(sd, data, win=7):
    """matcher.py:63-131, eval mode.  data keys: desc_3d [N3,192], pos_emd_3d, desc_2d_coarse [Mc,192], pos_emd_2d,
    kps3d [N3,3], kps2d [Mc,2], feat_fine [1,h,w,C], desc_3d_fine [N3,192], stride_coarse, stride_fine."""
    d3, d2 = self_cross_transformer(sd, "coarse_transformer", data["desc_3d"][None], data["pos_emd_3d"][None],
                                    data["desc_2d_coarse"][None], data["pos_emd_2d"][None])
    score = s2d_scores(sd, "coarse_matcher", d3[0], d2[0])
    i_ids, j_ids = mutual_matches(score)
    out = {"score_matrix": score, "i_ids": i_ids, "j_ids": j_ids, "desc_3d_t": d3[0], "desc_2d_t": d2[0],
           "mkps3d": data["kps3d"][i_ids], "mkps2d_c": data["kps2d"][j_ids]}
    M = len(i_ids)
    if M == 0:
        out.update({"expec_f": torch.empty(0, 3), "mkps2d_f": out["mkps2d_c"]})
        return out
    feat_fine = data["feat_fine"].permute(0, 3, 1, 2)
    m3 = data["desc_3d_fine"][i_ids][:, None, :]
    p3 = data["pos_emd_3d"][i_ids][:, None, :]
    wins = _lin(sd, "fine_preprocess.proj", fine_windows(feat_fine, j_ids, data["stride_coarse"] // data["stride_fine"], win))
    pe = pos_embed_2d(wins[..., 0].view(M, win, win)).view(M, win * win, -1)
    m3, wins = self_cross_transformer(sd, "fine_transformer", m3, p3, wins, pe)
    expec, mk = fine_match(sd, "fine_matcher", m3[:, 0, :], wins, out["mkps2d_c"], win)
    out.update({"expec_f": expec, "mkps2d_f": mk, "fine_desc_3d": m3[:, 0, :], "fine_windows": wins})
    return out
