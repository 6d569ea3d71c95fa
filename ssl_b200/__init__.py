"""ssl_b200 -- the Self-Similarity Graph (SSG) loss of ChrisDud0257/SSL, rebuilt for B200 (sm_100a).

Public surface:
    SelfSimilarityLoss, ssl          batched fused loss (names from BASELINE.json north_star)
    similarity_map                   drop-in for basicsr.losses.loss_util.similarity_map
    compute_similarity               drop-in for basicsr.losses.similarity.similaritywrapper.compute_similarity
    build_edge_list, ssg_rows, laplacian_mask   the pieces
    ssl_step_host                    the same step on host buffers (H2D + step + D2H in one C-ABI call)
    paired_random_crop_img_mask, TrainingPairPool   the crop and the pair pool that carry the mask upstream of the loss
Importing the package does not need a GPU; calling any operator without CUDA tensors or without
the built library raises (there is no fallback path).
"""
from .compat import similarity_map
from .functional import (EdgeList, build_edge_list, compute_similarity, laplacian_mask, ssg_rows,
                         ssl_step_host)
from .loss import SelfSimilarityLoss, ssl
from .pool import TrainingPairPool, paired_random_crop_img_mask

__all__ = ["SelfSimilarityLoss", "ssl", "similarity_map", "compute_similarity", "build_edge_list", "ssg_rows",
           "laplacian_mask", "EdgeList", "ssl_step_host", "TrainingPairPool", "paired_random_crop_img_mask"]
__version__ = "0.1.0"
