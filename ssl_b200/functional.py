"""Torch-facing operators over the C ABI: edge lists, SSG rows (with autograd) and the fused loss.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic of the
SSG path runs in libssl_b200.so.  Nothing in this module has a CPU implementation: CPU tensors
raise, exactly one code path exists.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ROWS_EXP, ROWS_NORM, ROWS_RAW


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"ssl_b200: `{name}` must be a CUDA tensor (got {getattr(t, 'device', type(t))}); "
                           "there is no CPU path")


def _check_kernel_sizes(ks: int, kw: int, h: int, w: int) -> None:
    if ks < 1 or ks % 2 == 0 or kw < 1 or kw % 2 == 0:
        raise ValueError(f"kernel_size_search ({ks}) and kernel_size_window ({kw}) must be odd")
    if kw > ks:
        # for K > P the two reference paths disagree with each other (similarity.cu:43 vs the
        # zero-padded unfold of loss_util.py:208), so there is nothing well defined to match
        raise ValueError(f"kernel_size_window ({kw}) must not exceed kernel_size_search ({ks})")
    if ks // 2 >= min(h, w):
        raise ValueError(f"reflect padding by {ks // 2} needs crops larger than {h}x{w}")


def rows_mode(generalization: bool, raw: bool = False) -> int:
    return ROWS_RAW if raw else (ROWS_NORM if generalization else ROWS_EXP)


# ------------------------------------------------------------------------------------------
# edge list
# ------------------------------------------------------------------------------------------

@dataclass
class EdgeList:
    """Device-resident list of edge pixels of a batch (flat indices b*H*W + y*W + x, ascending)."""
    edges: torch.Tensor    # int32 [capacity]
    counts: torch.Tensor   # int32 [2 + B]: written, found, per-image
    batch: int
    height: int
    width: int

    @property
    def capacity(self) -> int:
        return self.edges.numel()

    def count(self) -> int:
        """Number of listed edge pixels.  Device->host sync (4 bytes)."""
        found = int(self.counts[1].item())
        if found > self.capacity:
            raise RuntimeError(f"edge list overflow: {found} edge pixels > capacity {self.capacity}")
        return found

    def positions(self) -> torch.Tensor:
        """int64 [n, 3] (b, y, x) -- for inspection and tests."""
        n = self.count()
        flat = self.edges[:n].long()
        hw = self.height * self.width
        return torch.stack([flat // hw, (flat % hw) // self.width, flat % self.width], dim=1)


def build_edge_list(mask: torch.Tensor, mask_stride: int = 0, capacity: Optional[int] = None) -> EdgeList:
    """Edge pixels of ``mask`` [B,1|3,H,W] (channel 0, value == 1 exactly), no host sync.

    Batched replacement of ``nonzero(mask_pad == 1)`` (similaritywrapper.py:64-68) and of the
    ``mask_stride`` product / empty test of realesrganssl_model.py:385-388.
    """
    _require_cuda(mask, "mask")
    if mask.dim() == 3:
        mask = mask.unsqueeze(0)
    if mask.dim() != 4:
        raise ValueError(f"mask must be [B,1|3,H,W], got {tuple(mask.shape)}")
    if mask.dtype != torch.float32 or not mask.is_contiguous():
        mask = mask.contiguous().float()
    b, mc, h, w = mask.shape
    n_px = b * h * w
    cap = n_px if capacity is None else int(capacity)
    edges = torch.empty(max(cap, 1), dtype=torch.int32, device=mask.device)
    counts = torch.empty(2 + b, dtype=torch.int32, device=mask.device)
    ws_bytes = int(_lib.load().ssl_b200_edge_list_workspace_bytes(n_px))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.call("ssl_b200_build_edge_list", _ptr(mask), b, mc, h, w, int(mask_stride), _ptr(edges), cap,
                  _ptr(counts), _ptr(ws), ws_bytes, _stream())
    return EdgeList(edges[:cap] if cap else edges[:0], counts, b, h, w)


def laplacian_mask(gt: torch.Tensor, threshold: float = 20.0) -> torch.Tensor:
    """Edge mask of generate_mask.py:22-31 computed on the GT crop; float 0/1 [B,1,H,W]."""
    _require_cuda(gt, "gt")
    if gt.dim() != 4 or gt.shape[1] != 3:
        raise ValueError(f"gt must be [B,3,H,W], got {tuple(gt.shape)}")
    gt = gt.contiguous()
    b, _, h, w = gt.shape
    mask = torch.empty(b, 1, h, w, dtype=torch.float32, device=gt.device)
    with torch.cuda.device(gt.device):
        _lib.call("ssl_b200_laplacian_mask", _ptr(gt), _lib.dtype_code(gt.dtype), b, h, w, float(threshold),
                  _ptr(mask), _stream())
    return mask


# ------------------------------------------------------------------------------------------
# SSG rows with autograd
# ------------------------------------------------------------------------------------------

def _rows_forward(img, img2, el: EdgeList, n: int, ks, kw, sigma, eps, mode, n_dev=True):
    b, c, h, w = img.shape
    rows = torch.empty(n, ks * ks, dtype=torch.float32, device=img.device)
    rows2 = torch.empty_like(rows) if img2 is not None else None
    if n:
        with torch.cuda.device(img.device):
            _lib.call("ssl_b200_ssg_rows_forward", _ptr(img), _ptr(img2), _lib.dtype_code(img.dtype), b, c, h, w,
                      _ptr(el.edges), _ptr(el.counts) if n_dev else _ptr(None), n, ks, kw, float(sigma), float(eps),
                      mode, _ptr(rows), _ptr(rows2), _stream())
    return rows, rows2


def _rows_backward(img, el: EdgeList, n: int, ks, kw, gq, n_dev=True):
    b, c, h, w = img.shape
    grad = torch.zeros(b, c, h, w, dtype=torch.float32, device=img.device)
    if n:
        with torch.cuda.device(img.device):
            _lib.call("ssl_b200_ssg_rows_backward", _ptr(img), _lib.dtype_code(img.dtype), b, c, h, w, _ptr(el.edges),
                      _ptr(el.counts) if n_dev else _ptr(None), n, ks, kw, _ptr(gq), _ptr(grad), _stream())
    return grad


def _use_plane(path: str, b: int, c: int, h: int, w: int, ks: int, kw: int, n: int) -> bool:
    """Same rule as the library's SSL_B200_PATH_AUTO: plane kernels when they exist for (k_s, k_w, C) and the
    mask holds at least 2 % of the pixels (they pay per image pixel, the point kernels per edge pixel)."""
    if path == "point" or not _lib.load().ssl_b200_plane_supported(ks, kw, c):
        if path == "plane":
            raise ValueError(f"no plane kernels for k_s={ks} k_w={kw} C={c}")
        return False
    if n + 3 * b * (h // 48 + 1) * (w // 64 + 2) * 8 >= (1 << 23):   # slot packing of the backward lists (ssl_b200.cu: plane_slots_fit)
        if path == "plane":
            raise ValueError(f"batch too large for the plane kernels ({n} edge pixels); use path='auto' or 'point'")
        return False
    return path == "plane" or n >= 0.02 * b * h * w


def _plane_rows_forward(img, el: EdgeList, n: int, ks, kw, sigma, eps, mode):
    b, c, h, w = img.shape
    rows = torch.empty(n, ks * ks, dtype=torch.float32, device=img.device)
    if n:
        lib = _lib.load()
        ws_bytes = int(lib.ssl_b200_plane_rows_workspace_bytes(b, h, w, ks, kw, n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
        with torch.cuda.device(img.device):
            _lib.call("ssl_b200_plane_rows_forward", _ptr(img), _ptr(None), _lib.dtype_code(img.dtype), b, c, h, w,
                      _ptr(el.edges), _ptr(el.counts), n, ks, kw, _ptr(rows), _ptr(None), _ptr(ws), ws_bytes, _stream())
            if mode != ROWS_RAW:
                _lib.call("ssl_b200_rows_from_distance", _ptr(rows), _ptr(el.counts), n, ks, kw, c, float(sigma),
                          float(eps), mode, _stream())
    return rows


def _plane_rows_backward(img, el: EdgeList, n: int, ks, kw, gq):
    b, c, h, w = img.shape
    grad = torch.zeros(b, c, h, w, dtype=torch.float32, device=img.device)
    if n:
        ws_bytes = int(_lib.load().ssl_b200_plane_rows_backward_workspace_bytes(b, h, w, ks, kw, n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
        with torch.cuda.device(img.device):
            _lib.call("ssl_b200_plane_rows_backward", _ptr(img), _lib.dtype_code(img.dtype), b, c, h, w, _ptr(el.edges),
                      _ptr(el.counts), n, ks, kw, _ptr(gq), _ptr(grad), _ptr(ws), ws_bytes, _stream())
    return grad


class _SSGRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, el, n, ks, kw, sigma, eps, mode, plane=False):
        img_c = img.contiguous()
        if plane:
            rows = _plane_rows_forward(img_c, el, n, ks, kw, sigma, eps, mode)
        else:
            rows, _ = _rows_forward(img_c, None, el, n, ks, kw, sigma, eps, mode)
        ctx.save_for_backward(img_c, rows)
        ctx.el, ctx.n, ctx.cfg, ctx.plane = el, n, (ks, kw, sigma, mode), plane
        return rows

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_rows):
        img, rows = ctx.saved_tensors
        ks, kw, sigma, mode = ctx.cfg
        n = ctx.n
        gq = grad_rows.contiguous().float().clone()
        if n and mode != ROWS_RAW:
            with torch.cuda.device(img.device):
                _lib.call("ssl_b200_rows_grad_to_distance_grad", _ptr(rows), _ptr(gq), _ptr(ctx.el.counts), n, ks, kw,
                          img.shape[1], float(sigma), mode, _stream())
        grad = (_plane_rows_backward if ctx.plane else _rows_backward)(img, ctx.el, n, ks, kw, gq)
        return grad.to(img.dtype), None, None, None, None, None, None, None, None


def ssg_rows(img: torch.Tensor, edge_list: EdgeList, kernel_size_search: int = 25, kernel_size_window: int = 9,
             sigma: float = 0.004, generalization: bool = True, eps: float = 1e-10, raw: bool = False,
             n: Optional[int] = None, path: str = "auto") -> torch.Tensor:
    """Similarity rows [n, k_s^2] of every listed edge pixel of ``img`` [B,C,H,W]; differentiable in ``img``.

    Equals ``torch.cat([similarity_map(img_b, mask_b, ...).getitem() for b ...], dim=1)[0]`` of the
    reference (loss_util.py:165-248).  ``n`` defaults to ``edge_list.count()`` (one 4-byte sync).
    ``path``: "auto" | "point" | "plane" (which kernel family computes the rows and their backward).
    """
    _require_cuda(img, "img")
    if img.dim() != 4:
        raise ValueError(f"img must be [B,C,H,W], got {tuple(img.shape)}")
    b, c, h, w = img.shape
    if (b, h, w) != (edge_list.batch, edge_list.height, edge_list.width):
        raise ValueError("edge list was built for a different batch shape")
    _check_kernel_sizes(kernel_size_search, kernel_size_window, h, w)
    n = edge_list.count() if n is None else int(n)
    plane = _use_plane(path, b, c, h, w, int(kernel_size_search), int(kernel_size_window), n)
    return _SSGRows.apply(img, edge_list, n, int(kernel_size_search), int(kernel_size_window), float(sigma),
                          float(eps), rows_mode(generalization, raw), plane)


def compute_similarity(image: torch.Tensor, mask: torch.Tensor, psize: int = 25, ksize: int = 9) -> torch.Tensor:
    """Same contract as the reference ``compute_similarity`` (similaritywrapper.py:59-69):
    image [C,H,W], mask [H,W] -> raw patch distances [mc, psize, psize], differentiable in image."""
    _require_cuda(image, "image")
    _require_cuda(mask, "mask")
    el = build_edge_list(mask.reshape(1, 1, *mask.shape[-2:]))
    rows = ssg_rows(image.unsqueeze(0), el, psize, ksize, raw=True)
    return rows.view(-1, psize, psize)


# ------------------------------------------------------------------------------------------
# fused loss: rows(SR), rows(GT), L1 (+KL), backward -- rows never leave the op
# ------------------------------------------------------------------------------------------

class _SSLLoss(torch.autograd.Function):
    """total = w_l1 * mean|S_sr - S_gt| + w_kl * KL_mean(S_sr, S_gt); d total / d sr computed in forward.

    The 1/N of the 'mean' is applied with the (optionally all-reduced) global element count, so the
    kernels never wait for it; the backward pass is a single scale of the stored gradient.
    ``grad_scale`` multiplies the gradient only (parity="global_ddp": world size, see dist.py).
    """

    @staticmethod
    def forward(ctx, sr, gt, mask, mask_stride, n, ks, kw, sigma, eps, mode, w_l1, w_kl, reducer, path=0,
                grad_scale=1.0):
        sr_c, gt_c = sr.contiguous(), gt.contiguous()
        dev = sr_c.device
        need_grad = ctx.needs_input_grad[0]
        b, c, h, w = sr_c.shape
        lib = _lib.load()
        if gt_c.dtype != sr_c.dtype and not _use_plane({0: "auto", 1: "point", 2: "plane"}[int(path)], b, c, h, w, ks,
                                                        kw, n):
            # mixed precision (e.g. bf16 SR under autocast, fp32 GT) on the point kernels, which take one element
            # type: never round the target graph's input -- both images go in as fp32 (the plane path reads each
            # image in its own type; the arithmetic is fp32 either way)
            sr_c, gt_c = sr_c.float(), gt_c.float()
        terms = torch.empty(3, dtype=torch.float64, device=dev)   # sum|d|, sum KL, n_rows
        grad = torch.empty(sr_c.shape, dtype=torch.float32, device=dev) if need_grad else None
        ws_bytes = int(lib.ssl_b200_loss_step_workspace_bytes(b, c, h, w, ks, kw, n, int(path)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("ssl_b200_loss_step", _ptr(sr_c), _lib.dtype_code(sr_c.dtype), _ptr(gt_c),
                      _lib.dtype_code(gt_c.dtype), _ptr(mask), mask.shape[1], int(mask_stride), b, c, h, w, n, ks, kw,
                      float(sigma), float(eps), mode, float(w_l1), float(w_kl), _ptr(grad), _ptr(terms), _ptr(ws),
                      ws_bytes, int(path), _stream())
        if reducer is not None:
            terms = reducer(terms)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)   # total, l1 part, kl part, gradient factor
        with torch.cuda.device(dev):
            _lib.call("ssl_b200_loss_from_terms", _ptr(terms), ks * ks, float(w_l1), float(w_kl), float(grad_scale),
                      _ptr(out4), _stream())
        total, part_l1, part_kl = out4[0], out4[1], out4[2]
        ctx.grad_sr = grad
        ctx.inv_n = out4[3]
        ctx.sr_dtype = sr.dtype
        ctx.mark_non_differentiable(part_l1, part_kl)  # logging values; `total` carries the gradient
        return total, part_l1, part_kl

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_total, _g_l1, _g_kl):
        if ctx.grad_sr is None:
            return (None,) * 15
        g = (ctx.grad_sr * (g_total.to(torch.float32) * ctx.inv_n)).to(ctx.sr_dtype)
        return (g,) + (None,) * 14


def ssl_step_host(sr, gt, mask, kernel_size_search: int = 25, kernel_size_window: int = 9, sigma: float = 0.004,
                  generalization: bool = True, eps: float = 1e-10, loss_weight: float = 1.0, kl_weight: float = 0.0,
                  mask_stride: int = 0, want_grad: bool = True, out_grad: Optional[torch.Tensor] = None,
                  device: Optional[torch.device] = None):
    """The whole SSL block on HOST tensors through the C ABI's host entry (ssl_b200_loss_step_host):
    H2D of sr/gt/mask, edge list, fused step, 'mean', D2H of (total, l1, kl) and d total/d sr.

    sr, gt: fp32 CPU [B,C,H,W]; mask: fp32 CPU [B,1|3,H,W] (pin them for full copy speed).
    Returns (loss[3] CPU tensor, grad CPU tensor or None, n_rows).  Synchronous.
    """
    for name, t in (("sr", sr), ("gt", gt), ("mask", mask)):
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"ssl_step_host: `{name}` must be a contiguous fp32 CPU tensor")
    if not torch.cuda.is_available():
        raise RuntimeError("ssl_b200: no CUDA device; there is no CPU path")
    b, c, h, w = sr.shape
    if gt.shape != sr.shape or mask.shape[0] != b or tuple(mask.shape[-2:]) != (h, w):
        raise ValueError("sr / gt / mask shapes do not agree")
    _check_kernel_sizes(kernel_size_search, kernel_size_window, h, w)
    loss = torch.empty(3, dtype=torch.float32).pin_memory()
    grad = None
    if want_grad:
        grad = out_grad if out_grad is not None else torch.empty(sr.shape, dtype=torch.float32).pin_memory()
    n_rows = ctypes.c_int64(0)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        _lib.call("ssl_b200_loss_step_host", _ptr(sr), _ptr(gt), _ptr(mask), mask.shape[1], b, c, h, w,
                  int(mask_stride), int(kernel_size_search), int(kernel_size_window), float(sigma), float(eps),
                  rows_mode(generalization), float(loss_weight), float(kl_weight), _ptr(loss), _ptr(grad),
                  ctypes.byref(n_rows), _stream())
    return loss, grad, int(n_rows.value)
