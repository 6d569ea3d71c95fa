"""The step just upstream of the loss: the paired random crop and the training-pair pool
that carry the edge mask along with LQ / GT.

* ``paired_random_crop_img_mask`` -- same signature, same random draws (Python's ``random.randint``, top first,
  then left) and same results as GAN-Based-SR/basicsr/data/transforms.py:93-149.  CUDA tensors are cropped by
  ``ssl_b200_crop`` into contiguous tensors (the reference returns views and makes them contiguous later);
  NumPy arrays (the dataset-side use of the reference) are sliced exactly as the reference does.
* ``TrainingPairPool`` -- ``_dequeue_and_enqueue`` of GAN-Based-SR/basicsr/models/realesrganssl_model.py:326-367
  (same in Diffusion-Based-SR/ldm/models/diffusion/ddpmssl.py:297-339), including its quirk: the mask queue is
  allocated with the GT's channel count, so a 1-channel mask comes back from a full pool with 3 identical channels
  (:339-341).  The reference physically permutes the whole pool every step (``queue = queue[randperm]``, 3 x
  ``queue_size`` images gathered); here the permutation is kept as an index (``order``) and only the ``b`` samples
  that leave / enter are moved (``ssl_b200_pool_exchange``).  The draws (``torch.randperm(queue_size)`` on the CPU
  generator) and every returned batch are identical to the reference's -- see tests/test_pool.py.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .functional import _ptr, _require_cuda, _stream


def _crop_cuda(t: torch.Tensor, top: int, left: int, h: int, w: int) -> torch.Tensor:
    _require_cuda(t, "tensor")
    if t.dim() != 4:
        raise ValueError(f"expected [B,C,H,W], got {tuple(t.shape)}")
    if t.element_size() != 4:
        raise TypeError(f"ssl_b200_crop moves 4-byte elements, got {t.dtype}")
    t = t.contiguous()
    b, c, hh, ww = t.shape
    out = torch.empty(b, c, h, w, dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        _lib.call("ssl_b200_crop", _ptr(t), _ptr(out), b * c, hh, ww, int(top), int(left), int(h), int(w), _stream())
    return out


def paired_random_crop_img_mask(img_gts, img_lqs, masks, gt_patch_size, scale, gt_path=None):
    """Drop-in for basicsr.data.transforms.paired_random_crop_img_mask (transforms.py:93-149)."""
    if not isinstance(img_gts, list):
        img_gts = [img_gts]
    if not isinstance(img_lqs, list):
        img_lqs = [img_lqs]
    if not isinstance(masks, list):
        masks = [masks]
    is_tensor = torch.is_tensor(img_gts[0])
    if is_tensor:
        h_lq, w_lq = img_lqs[0].size()[-2:]
        h_gt, w_gt = img_gts[0].size()[-2:]
    else:
        h_lq, w_lq = img_lqs[0].shape[0:2]
        h_gt, w_gt = img_gts[0].shape[0:2]
    lq_patch_size = gt_patch_size // scale
    if h_gt != h_lq * scale or w_gt != w_lq * scale:
        raise ValueError(f'Scale mismatches. GT ({h_gt}, {w_gt}) is not {scale}x ',
                         f'multiplication of LQ ({h_lq}, {w_lq}).')
    if h_lq < lq_patch_size or w_lq < lq_patch_size:
        raise ValueError(f'LQ ({h_lq}, {w_lq}) is smaller than patch size '
                         f'({lq_patch_size}, {lq_patch_size}). '
                         f'Please remove {gt_path}.')
    # the reference's two draws, in its order (transforms.py:126-127)
    top = random.randint(0, h_lq - lq_patch_size)
    left = random.randint(0, w_lq - lq_patch_size)
    top_gt, left_gt = int(top * scale), int(left * scale)
    if is_tensor:
        img_lqs = [_crop_cuda(v, top, left, lq_patch_size, lq_patch_size) for v in img_lqs]
        img_gts = [_crop_cuda(v, top_gt, left_gt, gt_patch_size, gt_patch_size) for v in img_gts]
        masks = [_crop_cuda(v, top_gt, left_gt, gt_patch_size, gt_patch_size) for v in masks]
    else:
        img_lqs = [v[top:top + lq_patch_size, left:left + lq_patch_size, ...] for v in img_lqs]
        img_gts = [v[top_gt:top_gt + gt_patch_size, left_gt:left_gt + gt_patch_size, ...] for v in img_gts]
        masks = [v[top_gt:top_gt + gt_patch_size, left_gt:left_gt + gt_patch_size, ...] for v in masks]
    if len(img_gts) == 1:
        img_gts = img_gts[0]
    if len(masks) == 1:
        masks = masks[0]
    if len(img_lqs) == 1:
        img_lqs = img_lqs[0]
    return img_gts, img_lqs, masks


class PoolOrder:
    """Index bookkeeping of the pool, free of tensors (so it is testable without a GPU).

    ``order[k]`` = physical slot of the sample at logical position k.  The reference's
    ``queue = queue[idx]`` is ``order = order[idx]``; the samples it then reads and overwrites at positions
    0..b-1 live in slots ``order[:b]``."""

    def __init__(self, queue_size: int):
        self.queue_size = int(queue_size)
        self.order: List[int] = list(range(self.queue_size))
        self.ptr = 0

    def full(self) -> bool:
        return self.ptr == self.queue_size

    def enqueue_slots(self, b: int) -> List[int]:
        """Slots the next b samples are stored in while the pool is filling (realesrganssl_model.py:362-366)."""
        slots = self.order[self.ptr:self.ptr + b]
        self.ptr += b
        return slots

    def exchange_slots(self, b: int, idx: Sequence[int]) -> List[int]:
        """Apply the step's permutation ``idx`` and return the slots of the b samples that are dequeued and
        replaced by the incoming batch (:344-356)."""
        self.order = [self.order[int(i)] for i in idx]
        return self.order[:b]


class TrainingPairPool:
    """``_dequeue_and_enqueue`` as an object: ``lq, gt, gt_mask = pool(lq, gt, gt_mask)``."""

    def __init__(self, queue_size: int):
        self.queue_size = int(queue_size)
        self.idx: Optional[PoolOrder] = None
        self.queue_lr = self.queue_gt = self.queue_gt_mask = None

    def _exchange(self, queue, t, out, slots_dev, b, bcast):
        sample = queue[0].numel()
        with torch.cuda.device(queue.device):
            _lib.call("ssl_b200_pool_exchange", _ptr(queue), _ptr(t), _ptr(out), _ptr(slots_dev), b, sample, bcast, _stream())

    @torch.no_grad()
    def __call__(self, lq: torch.Tensor, gt: torch.Tensor, gt_mask: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        for name, t in (("lq", lq), ("gt", gt), ("gt_mask", gt_mask)):
            _require_cuda(t, name)
            if t.dtype != torch.float32:
                raise TypeError(f"the pool holds float32 tensors (like the reference's torch.zeros queues), {name} is {t.dtype}")
        b, c, h, w = lq.size()
        if self.idx is None:
            assert self.queue_size % b == 0, f'queue size {self.queue_size} should be divisible by batch size {b}'
            self.queue_lr = torch.zeros(self.queue_size, c, h, w, device=lq.device)
            _, c, h, w = gt.size()
            self.queue_gt = torch.zeros(self.queue_size, c, h, w, device=gt.device)
            # the reference's quirk (:339-341): the mask queue takes the GT's channel count
            self.queue_gt_mask = torch.zeros(self.queue_size, c, h, w, device=gt.device)
            self.idx = PoolOrder(self.queue_size)
        cq = self.queue_gt_mask.shape[1]
        cm = gt_mask.shape[1]
        if cm != cq and cm != 1:
            raise RuntimeError(f"mask with {cm} channels cannot be assigned into a pool of {cq}-channel masks")
        bcast = cq if cm == 1 and cq > 1 else 0
        lq_c, gt_c, m_c = lq.contiguous(), gt.contiguous(), gt_mask.contiguous()
        if self.idx.full():
            perm = torch.randperm(self.queue_size)                       # the reference's draw (:346), CPU generator
            slots = self.idx.exchange_slots(b, perm.tolist())
            slots_dev = torch.tensor(slots, dtype=torch.int32).to(lq.device, non_blocking=True)
            lq_out = torch.empty_like(lq_c)
            gt_out = torch.empty_like(gt_c)
            m_out = torch.empty(b, cq, *gt_c.shape[-2:], device=gt.device)
            self._exchange(self.queue_lr, lq_c, lq_out, slots_dev, b, 0)
            self._exchange(self.queue_gt, gt_c, gt_out, slots_dev, b, 0)
            self._exchange(self.queue_gt_mask, m_c, m_out, slots_dev, b, bcast)
            return lq_out, gt_out, m_out
        slots = self.idx.enqueue_slots(b)
        slots_dev = torch.tensor(slots, dtype=torch.int32).to(lq.device, non_blocking=True)
        self._exchange(self.queue_lr, lq_c, None, slots_dev, b, 0)
        self._exchange(self.queue_gt, gt_c, None, slots_dev, b, 0)
        self._exchange(self.queue_gt_mask, m_c, None, slots_dev, b, bcast)
        return lq, gt, gt_mask
