"""Seeded synthetic crops for parity tests and the benchmark.

Everything is generated on the CPU in fp32 from a ``torch.Generator`` so the CPU oracle and the
GPU path see identical bits.  The recipe gives a smooth, natural-like GT (blurred noise plus a few
step edges) and ``SR = GT + small noise`` so that the similarity graph at sigma = 0.004 is not
one-hot; i.i.d. images collapse the SSG to a delta and make every loss ~1e-16.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _gaussian_taps(sigma_px: float, radius: int) -> torch.Tensor:
    x = torch.arange(-radius, radius + 1, dtype=torch.float32)
    k = torch.exp(-0.5 * (x / sigma_px) ** 2)
    return k / k.sum()


def _blur(x: torch.Tensor, sigma_px: float = 3.0, radius: int = 9) -> torch.Tensor:
    k = _gaussian_taps(sigma_px, radius)
    c = x.shape[1]
    x = F.pad(x, (radius, radius, radius, radius), mode="reflect")
    x = F.conv2d(x, k.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)
    x = F.conv2d(x, k.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)
    return x


def make_images(batch: int, height: int, width: int, seed: int, noise: float = 0.02, channels: int = 3):
    """Return (sr, gt) fp32 CPU tensors [B,C,H,W] in [0,1]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    field = _blur(torch.randn(batch, channels, height, width, generator=g))
    field = field / field.std(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)
    yy = torch.arange(height, dtype=torch.float32).view(1, 1, height, 1)
    xx = torch.arange(width, dtype=torch.float32).view(1, 1, 1, width)
    steps = torch.zeros(batch, channels, height, width)
    for _ in range(8):
        theta = torch.rand(batch, 1, 1, 1, generator=g) * (2 * math.pi)
        off = torch.rand(batch, 1, 1, 1, generator=g)
        amp = torch.rand(batch, channels, 1, 1, generator=g) - 0.5
        side = (torch.cos(theta) * (xx - off * width) + torch.sin(theta) * (yy - off * height)) > 0
        steps = steps + amp * side
    gt = (0.5 + 0.10 * field + 0.15 * steps).clamp(0.0, 1.0)
    sr = (gt + noise * torch.randn(batch, channels, height, width, generator=g)).clamp(0.0, 1.0)
    return sr.contiguous(), gt.contiguous()


def make_mask(batch: int, height: int, width: int, seed: int, density: float = 0.114,
              force_borders: bool = True) -> torch.Tensor:
    """Bernoulli(density) float 0/1 mask [B,1,H,W]; corners and one pixel per border are forced on
    so the reflect path is always exercised."""
    g = torch.Generator(device="cpu").manual_seed(seed + 7919)
    m = (torch.rand(batch, 1, height, width, generator=g) < density).float()
    if force_borders:
        for y in (0, height - 1):
            for x in (0, width - 1):
                m[:, :, y, x] = 1.0
        m[:, :, 0, width // 2] = 1.0
        m[:, :, height - 1, width // 3] = 1.0
        m[:, :, height // 2, 0] = 1.0
        m[:, :, height // 3, width - 1] = 1.0
    return m.contiguous()


def make_case(batch: int, height: int, width: int, seed: int, density: float = 0.114, noise: float = 0.02):
    sr, gt = make_images(batch, height, width, seed, noise)
    return sr, gt, make_mask(batch, height, width, seed, density)
