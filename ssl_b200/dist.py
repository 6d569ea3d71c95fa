"""Multi-GPU plumbing of the SSG loss: one process per GPU, batch sharded by image.

The path is embarrassingly parallel over images (every row depends only on its own crop), so the
only exchange is one all-reduce of three scalars [sum|d|, sum KL, n_rows] at the end of the forward
(SURVEY.md section 8e).  The reference has no collective on this path at all: under DDP every
rank takes the mean over its local rows and DDP averages parameter gradients
(GAN-Based-SR/basicsr/models/base_model.py:94-98) -- that is ``parity="ddp"`` here (no
communication, the default); ``parity="global"`` normalises by the global element count so N ranks
reproduce the single-device loss and gradient on the concatenated batch, and ``"global_ddp"`` does the
same for callers whose parameter gradients DDP will average afterwards.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


PARITIES = ("ddp", "global", "global_ddp")


def make_reducer(parity: str = "ddp", group=None) -> Optional[Callable[[torch.Tensor], torch.Tensor]]:
    """Return the function applied to the local [sum_l1, sum_kl, n_rows] float64 vector (None = no exchange).

    ``"ddp"``         the reference's behaviour under DistributedDataParallel: every rank takes the mean over
                      its LOCAL rows, no communication; DDP then averages parameter gradients across ranks
                      (base_model.py:94-98).  Default, so the module drops into the reference's trainers.
    ``"global"``      loss AND gradient of the single-device run on the concatenated batch: both are
                      normalised by the all-reduced global element count.  For callers that SUM (or do not
                      reduce) gradients across ranks -- the benchmark, evaluation.
    ``"global_ddp"``  the global loss value, with the gradient multiplied by the world size so that DDP's
                      averaging of parameter gradients reproduces the single-device gradient exactly.
    """
    if parity not in PARITIES:
        raise ValueError(f"parity must be one of {PARITIES}, got {parity!r}")
    if parity == "ddp" or not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_world_size(group) == 1:
        return None

    def reduce_terms(terms: torch.Tensor) -> torch.Tensor:
        out = terms.clone()
        if dist.get_backend(group) == "gloo" and out.is_cuda:
            host = out.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
            return host.to(out.device)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)  # 24 bytes; latency-bound on NVLink
        return out

    return reduce_terms


def grad_scale(parity: str, group=None) -> float:
    """Factor applied to the gradient only: the world size for "global_ddp" (DDP divides it out again)."""
    if parity == "global_ddp" and dist.is_available() and dist.is_initialized():
        return float(dist.get_world_size(group))
    return 1.0


def shard_range(global_batch: int, rank: int, world_size: int) -> range:
    """Images [r*B/W, (r+1)*B/W) go to rank r (what EnlargedSampler + DDP do, data_sampler.py:6-48)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return range(rank * per, (rank + 1) * per)


def mean_from_terms(terms: torch.Tensor, row_len: int, w_l1: float = 1.0, w_kl: float = 0.0) -> torch.Tensor:
    """Loss value from (possibly all-reduced) terms; shared by the op and by the CPU tests."""
    n_tot = (terms[2] * row_len).clamp_min(1.0)
    return w_l1 * terms[0] / n_tot + w_kl * terms[1] / n_tot
