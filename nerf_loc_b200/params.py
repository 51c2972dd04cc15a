"""Parameter inventory of the hot path, using the reference's own `state_dict` names.

`conditional_nerf_shapes` / `matcher_shapes` list every tensor a reference
checkpoint holds for `ConditionalNeRF` (nerf_loc/models/conditional_nerf/model.py:29-135)
and `Matcher` (nerf_loc/models/matcher.py:10-61) that the render / match path
reads.  The per-frame-setup CNN (`multiview_aggregator.depth_fusion.*`,
conditional_nerf/depth_fusion.py:239-282) is listed separately because it runs
before the hot path.
"""
from collections import OrderedDict


def conditional_nerf_shapes(n_samples=64, C=192, W=128):
    S = n_samples
    sh = OrderedDict()

    def lin(name, o, i, bias=True):
        sh[name + ".weight"] = (o, i)
        if bias:
            sh[name + ".bias"] = (o,)

    lin("ray_diff_fc.0", 16, 4)
    lin("ray_diff_fc.2", 27, 16)
    for head, k in (("mean_decoder", 2), ("var_decoder", 2), ("aw_decoder", 1), ("vis_decoder", 1)):
        p = "multiview_aggregator.dist_decoder." + head
        lin(p + ".0", 32, 32)
        lin(p + ".2", 32, 32)
        lin(p + ".4", k, 32)
    lin("multiview_aggregator.out_fc.0", 64, (C + 3) * 2 + 3)
    lin("multiview_aggregator.out_fc.2", W, 64)
    lin("confidence_mlp.0", 64, W)
    lin("confidence_mlp.2", 1, 64)
    lin("keypoint_head.0", 1, C)
    lin("base_mlp.0", W, (C + 3) + 63 + 27)
    lin("base_mlp.2", W, W)
    lin("base_mlp.4", W, W)
    for n in ("w_qs", "w_ks", "w_vs", "fc"):
        lin("base_mlp_attn." + n, W, W, bias=False)
    sh["base_mlp_attn.layer_norm.weight"] = (W,)
    sh["base_mlp_attn.layer_norm.bias"] = (W,)
    lin("base_mlp_agg_weight.0", W, W)
    lin("base_mlp_agg_weight.2", 1, W)

    def conv(name, co, ci, s_level, transposed=False):
        sh[f"ray_unet.{name}.0.weight"] = (ci, co, 3) if transposed else (co, ci, 3)
        sh[f"ray_unet.{name}.0.bias"] = (co,)
        sh[f"ray_unet.{name}.1.weight"] = (co, s_level)
        sh[f"ray_unet.{name}.1.bias"] = (co, s_level)

    conv("conv1", 64, W, S)
    conv("conv2", 128, 64, S // 2)
    conv("conv3", 128, 128, S // 4)
    conv("trans_conv3", 128, 128, S // 4, True)
    conv("trans_conv2", 64, 256, S // 2, True)
    conv("trans_conv1", 32, 128, S, True)
    conv("conv_out", W, W + 32, S)
    lin("sigma_mlp.0", 1, W)
    lin("feat_mlp.0", W, W)
    lin("feat_mlp.2", C, W)
    lin("rgb_blending_mlp.0", 32, W + (3 + C) + 1 + 4)
    lin("rgb_blending_mlp.2", 16, 32)
    lin("rgb_blending_mlp.4", 1, 16)
    lin("beta_mlp.0", 1, W)
    lin("proj_layer_3d_coarse", 192, W + 3 + C)
    lin("proj_layer_3d_fine", 192, W + 3 + C)
    return sh


def _transformer_shapes(sh, pfx, d, ffn):
    for layer, attn in (("self_attn_layer0", "self_attn"), ("self_attn_layer1", "self_attn"),
                        ("cross_attn_layer0", "multihead_attn"), ("cross_attn_layer1", "multihead_attn")):
        p = f"{pfx}.{layer}"
        sh[f"{p}.{attn}.in_proj_weight"] = (3 * d, d)
        sh[f"{p}.{attn}.in_proj_bias"] = (3 * d,)
        sh[f"{p}.{attn}.out_proj.weight"] = (d, d)
        sh[f"{p}.{attn}.out_proj.bias"] = (d,)
        sh[f"{p}.linear1.weight"] = (ffn, d)
        sh[f"{p}.linear1.bias"] = (ffn,)
        sh[f"{p}.linear2.weight"] = (d, ffn)
        sh[f"{p}.linear2.bias"] = (d,)
        norms = ("norm1", "norm2") if attn == "self_attn" else ("norm1", "norm2", "norm3")
        for n in norms:
            sh[f"{p}.{n}.weight"] = (d,)
            sh[f"{p}.{n}.bias"] = (d,)


def matcher_shapes(d=192, c_fine=192):
    sh = OrderedDict()
    _transformer_shapes(sh, "coarse_transformer", d, 512)
    for p in ("coarse_matcher", ):
        sh[p + ".mlps.0.weight"] = (128, d); sh[p + ".mlps.0.bias"] = (128,)
        sh[p + ".mlps.2.weight"] = (128, 128); sh[p + ".mlps.2.bias"] = (128,)
        sh[p + ".mlps.4.weight"] = (1, 128); sh[p + ".mlps.4.bias"] = (1,)
    sh["fine_preprocess.proj.weight"] = (d, c_fine)
    sh["fine_preprocess.proj.bias"] = (d,)
    _transformer_shapes(sh, "fine_transformer", d, 128)
    p = "fine_matcher"
    sh[p + ".mlps.0.weight"] = (128, d); sh[p + ".mlps.0.bias"] = (128,)
    sh[p + ".mlps.2.weight"] = (128, 128); sh[p + ".mlps.2.bias"] = (128,)
    sh[p + ".mlps.4.weight"] = (1, 128); sh[p + ".mlps.4.bias"] = (1,)
    return sh
