"""nerf_loc_b200: B200-native (sm_100a) render-and-match hot path of NeRF-Loc behind the reference's call signatures.

    from nerf_loc_b200 import ConditionalNeRF, Matcher, knn_points, knn_gather

The arithmetic lives in libnerfloc_b200.so (C ABI: include/nerfloc_b200.h); importing this package does not need a GPU,
calling into it does.
"""
from .config import default_args  # noqa: F401
from .conditional_nerf import ConditionalNeRF  # noqa: F401
from .knn import KnnIndex, knn_gather, knn_points  # noqa: F401
from .matcher import Matcher, S2DMatching  # noqa: F401
