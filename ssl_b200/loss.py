"""``SelfSimilarityLoss`` / ``ssl()``: the batched, fused form of the reference training-step block.

One call replaces the per-image Python loop, the two ``similarity_map`` constructions per image,
the ``torch.cat`` of rows and the ``L1Loss`` / ``KLDistanceLoss`` modules of
GAN-Based-SR/basicsr/models/realesrganssl_model.py:378-430 (identical block in the seven other
``*ssl_model.py`` files, in train_BSGRAN/models/model_ssl.py:285-334 and in
Diffusion-Based-SR/ldm/models/diffusion/ddpmssl.py:438-513).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import functional as F_
from .dist import grad_scale, make_reducer


_PATHS = {"auto": 0, "point": 1, "plane": 2}


def ssl(sr: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None, kernel_size_search: int = 25,
        kernel_size_window: int = 9, sigma: float = 0.004, generalization: bool = True, eps: float = 1e-10,
        loss_weight: float = 1.0, kl_weight: float = 0.0, mask_stride: int = 0, mask_threshold: float = 20.0,
        max_edges: Optional[int] = None, parity: str = "ddp", group=None, return_parts: bool = False,
        path: str = "auto"):
    """Self-similarity loss of a batch.

    sr, gt : [B,C,H,W] CUDA tensors (fp32 / bf16 / fp16); gradients flow into ``sr`` only
             (the reference detaches nothing but GT never requires grad, realesrganssl_model.py:400).
    mask   : [B,1|3,H,W] float edge mask (channel 0, ``== 1``), or None for the on-GPU Laplacian mask
             of the GT crop (generate_mask.py:22-31).
    Returns ``loss_weight * mean|S_sr - S_gt| + kl_weight * KL`` as a 0-dim fp32 tensor -- the sum of
    the reference's ``l_selfsim`` and ``l_selfsim_kl``; with ``return_parts`` also the two terms
    (detached, for the loss dict).  An all-empty mask gives 0 (the reference omits the term).
    ``max_edges`` (rows capacity) makes the call free of host syncs and CUDA-graph capturable; if the mask
    holds more edge pixels than that, the loss and the gradient come back NaN (never a silently truncated
    batch).
    ``parity`` : "ddp" (default: mean over this rank's rows, what the reference does under DDP), "global"
    or "global_ddp" (normalise by the all-reduced global count; see ssl_b200/dist.py).
    ``path``: "auto" | "point" | "plane" -- which kernels run (include/ssl_b200.h SSL_B200_PATH_*).
    """
    F_._require_cuda(sr, "sr")
    F_._require_cuda(gt, "gt")
    if sr.dim() != 4 or sr.shape != gt.shape:
        raise ValueError(f"sr and gt must be [B,C,H,W] of equal shape, got {tuple(sr.shape)} / {tuple(gt.shape)}")
    b, c, h, w = sr.shape
    F_._check_kernel_sizes(kernel_size_search, kernel_size_window, h, w)
    if mask is None:
        mask = F_.laplacian_mask(gt.detach(), mask_threshold)
    elif mask.shape[0] != b or mask.shape[-2:] != (h, w):
        raise ValueError(f"mask {tuple(mask.shape)} does not match images {tuple(sr.shape)}")
    if mask.dtype != torch.float32 or not mask.is_contiguous():
        mask = mask.contiguous().float()
    if max_edges is None:
        # one 4-byte device->host read to size the rows workspace; give max_edges to stay asynchronous
        n = F_.build_edge_list(mask, mask_stride).count()
    else:
        n = int(max_edges)
    mode = F_.rows_mode(generalization)
    total, l1, kl = F_._SSLLoss.apply(sr, gt.detach(), mask, int(mask_stride), n, int(kernel_size_search),
                                      int(kernel_size_window), float(sigma), float(eps), mode, float(loss_weight),
                                      float(kl_weight), make_reducer(parity, group), _PATHS[path],
                                      grad_scale(parity, group))
    return (total, l1, kl) if return_parts else total


class SelfSimilarityLoss(nn.Module):
    """Module form of :func:`ssl`; parameter names follow the reference's ``ssl_setting`` /
    ``selfsim_opt`` YAML blocks (options/train/RealESRGANSSL/train_RealESRGANSSL_x4.yml:113-119,149-157)."""

    def __init__(self, kernel_size_search: int = 25, kernel_size_window: int = 9, sigma: float = 0.004,
                 generalization: bool = True, eps: float = 1e-10, loss_weight: float = 1.0, kl_weight: float = 0.0,
                 mask_stride: int = 0, mask_threshold: float = 20.0, max_edges: Optional[int] = None,
                 parity: str = "ddp", path: str = "auto"):
        super().__init__()
        self.path = path
        self.kernel_size_search = kernel_size_search
        self.kernel_size_window = kernel_size_window
        self.sigma = sigma
        self.generalization = generalization
        self.eps = eps
        self.loss_weight = loss_weight
        self.kl_weight = kl_weight
        self.mask_stride = mask_stride
        self.mask_threshold = mask_threshold
        self.max_edges = max_edges
        self.parity = parity
        self.last_l1 = None
        self.last_kl = None

    def forward(self, sr: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        total, self.last_l1, self.last_kl = ssl(
            sr, gt, mask, self.kernel_size_search, self.kernel_size_window, self.sigma, self.generalization, self.eps,
            self.loss_weight, self.kl_weight, self.mask_stride, self.mask_threshold, self.max_edges, self.parity,
            return_parts=True, path=self.path)
        return total

    def extra_repr(self) -> str:
        return (f"k_s={self.kernel_size_search}, k_w={self.kernel_size_window}, sigma={self.sigma}, "
                f"generalization={self.generalization}, loss_weight={self.loss_weight}, kl_weight={self.kl_weight}, "
                f"mask_stride={self.mask_stride}")
