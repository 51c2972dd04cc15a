"""Default hyper-parameters of the hot path (values of nerf_loc/configs/__init__.py:4-92) as plain namespaces, so the
package does not need yacs.  A yacs CfgNode from the reference works just as well wherever `args` is accepted."""
import types


def default_args(n_samples=64, n_importance=0):
    ns = types.SimpleNamespace
    return ns(
        backbone2d_fpn_dim=192, model_3d_hidden_dim=128, matcher_hidden_dim=192,
        use_scene_coord_memorization=False, encode_appearance=True, appearance_emb_dim=128,
        multires=10, multires_views=4, i_embed=0, use_depth_supervision=False,
        fine_matching_loss_type='l2_with_std',
        render=ns(N_samples=n_samples, N_importance=n_importance, N_rand=1024, chunk=2048, lindisp=False,
                  white_bkgd=False, use_render_uncertainty=True, render_feature=True),
        matching=ns(fine_num_3d_keypoints=1024, coarse_num_3d_keypoints=1024),
    )
