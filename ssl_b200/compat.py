"""Drop-in ``similarity_map`` with the reference's constructor signatures.

GAN side  (GAN-Based-SR/basicsr/losses/loss_util.py:165-248):
    similarity_map(img, mask, ssl_mode, kernel_size_search, generalization, kernel_size_window, sigma).getitem()
Diffusion side (Diffusion-Based-SR/basicsr/losses/loss_util.py:242-339,1239-1252), shipped strategy only:
    similarity_map(img, mask, simself_strategy='areaarea_mask_nonlocalavg_cuda_v1', kernel_size, scaling_factor,
                   kernel_size_center, softmax, **ignored).getitem()

Both return ``Tensor[1, num, k_s^2]`` with rows in row-major order of the edge pixels, differentiable
with respect to ``img``.  ``ssl_mode`` accepts the reference's 'cuda' and 'pytorch' (both are served
by the sm_100a kernels -- they are documented to be equivalent, README.md:109-125) and 'b200'.
"""
from __future__ import annotations

import torch

from . import functional as F_

_GAN_MODES = ("cuda", "pytorch", "b200")
_DM_STRATEGY = "areaarea_mask_nonlocalavg_cuda_v1"


class similarity_map:  # noqa: N801  (reference spelling)
    def __init__(self, img, mask=None, ssl_mode="cuda", kernel_size_search=5, generalization=True,
                 kernel_size_window=9, sigma=0.004, *, simself_strategy=None, kernel_size=None, scaling_factor=None,
                 kernel_size_center=None, softmax=None, **_ignored):
        eps = 1e-10  # loss_util.py:227,242
        if ssl_mode is not None and not isinstance(ssl_mode, str):
            # the diffusion-side constructor's third positional parameter is `img_sr` (loss_util.py:243), the GAN
            # side's is `ssl_mode`: a tensor here means diffusion-style positional arguments, which cannot be told
            # apart safely -- ddpmssl.py:452-467 passes everything by keyword, and so must other callers
            raise TypeError("similarity_map: pass the diffusion-side arguments (simself_strategy, kernel_size, ...) "
                            "by keyword; the third positional argument is ssl_mode")
        if simself_strategy is not None:  # diffusion-side keyword set (ddpmssl.py:452-467)
            if simself_strategy != _DM_STRATEGY:
                raise ValueError(f"only simself_strategy={_DM_STRATEGY!r} (the shipped configuration) is supported, "
                                 f"got {simself_strategy!r}")
            # defaults of the diffusion-side constructor (Diffusion-Based-SR/basicsr/losses/loss_util.py:243-247):
            # kernel_size=5 (the shared positional default), scaling_factor=4, kernel_size_center=9, softmax=True
            kernel_size_search = kernel_size if kernel_size is not None else kernel_size_search
            kernel_size_window = kernel_size_center if kernel_size_center is not None else 9
            sigma = scaling_factor if scaling_factor is not None else 4
            generalization = True if softmax is None else bool(softmax)
            eps = 1e-20  # Diffusion-Based-SR/basicsr/losses/loss_util.py:1250
        elif ssl_mode not in _GAN_MODES:
            raise ValueError("The ssl_mode should either be cuda or pytorch.")  # loss_util.py:178-179
        if mask is None:
            raise ValueError("similarity_map needs the edge mask")
        if img.dim() != 4 or img.shape[0] != 1:
            raise ValueError(f"img must be [1,C,H,W] (the reference reads img[0] only), got {tuple(img.shape)}")
        # ssl_pytorch unfolds the whole mask, so a [1,3,H,W] mask selects every edge pixel once per
        # channel, channel-major (loss_util.py:195-199); ssl_cuda reads channel 0 only
        channels = range(mask.shape[1]) if (simself_strategy is None and ssl_mode == "pytorch") else (0,)
        parts = []
        for ch in channels:
            el = F_.build_edge_list(mask[:, ch:ch + 1])
            parts.append(F_.ssg_rows(img, el, kernel_size_search, kernel_size_window, sigma, generalization, eps))
        rows = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        self.s = rows.unsqueeze(0)

    def getitem(self):
        return self.s
