"""paired_random_crop_img_mask and the training-pair pool, index-exact against a restatement of
the reference (GAN-Based-SR/basicsr/data/transforms.py:93-149, basicsr/models/realesrganssl_model.py:326-367).
Pure data movement: every comparison is bit-exact."""
import random

import numpy as np
import pytest
import torch


class ReferencePool:
    """`_dequeue_and_enqueue` restated line by line on whatever device the tensors live on
    (realesrganssl_model.py:326-367); `.cuda()` of the reference dropped."""

    def __init__(self, queue_size):
        self.queue_size = queue_size

    def step(self, lq, gt, gt_mask):
        b, c, h, w = lq.size()
        if not hasattr(self, 'queue_lr'):
            assert self.queue_size % b == 0
            self.queue_lr = torch.zeros(self.queue_size, c, h, w)
            _, c, h, w = gt.size()
            self.queue_gt = torch.zeros(self.queue_size, c, h, w)
            self.queue_gt_mask = torch.zeros(self.queue_size, c, h, w)
            self.queue_ptr = 0
        if self.queue_ptr == self.queue_size:
            idx = torch.randperm(self.queue_size)
            self.queue_lr = self.queue_lr[idx]
            self.queue_gt = self.queue_gt[idx]
            self.queue_gt_mask = self.queue_gt_mask[idx]
            lq_dequeue = self.queue_lr[0:b, :, :, :].clone()
            gt_dequeue = self.queue_gt[0:b, :, :, :].clone()
            gt_mask_dequeue = self.queue_gt_mask[0:b, :, :, :].clone()
            self.queue_lr[0:b, :, :, :] = lq.clone()
            self.queue_gt[0:b, :, :, :] = gt.clone()
            self.queue_gt_mask[0:b, :, :, :] = gt_mask.clone()
            return lq_dequeue, gt_dequeue, gt_mask_dequeue
        self.queue_lr[self.queue_ptr:self.queue_ptr + b, :, :, :] = lq.clone()
        self.queue_gt[self.queue_ptr:self.queue_ptr + b, :, :, :] = gt.clone()
        self.queue_gt_mask[self.queue_ptr:self.queue_ptr + b, :, :, :] = gt_mask.clone()
        self.queue_ptr = self.queue_ptr + b
        return lq, gt, gt_mask


def reference_crop(img_gts, img_lqs, masks, gt_patch_size, scale):
    """Tensor branch of transforms.py:93-149 for single tensors."""
    h_lq, w_lq = img_lqs.size()[-2:]
    lq_patch_size = gt_patch_size // scale
    top = random.randint(0, h_lq - lq_patch_size)
    left = random.randint(0, w_lq - lq_patch_size)
    lq = img_lqs[:, :, top:top + lq_patch_size, left:left + lq_patch_size]
    top_gt, left_gt = int(top * scale), int(left * scale)
    gt = img_gts[:, :, top_gt:top_gt + gt_patch_size, left_gt:left_gt + gt_patch_size]
    m = masks[:, :, top_gt:top_gt + gt_patch_size, left_gt:left_gt + gt_patch_size]
    return gt, lq, m


def _batches(n, b=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        lq = torch.rand(b, 3, 6, 5, generator=g)
        gt = torch.rand(b, 3, 24, 20, generator=g)
        m = (torch.rand(b, 1, 24, 20, generator=g) < 0.3).float()
        out.append((lq, gt, m))
    return out


def test_pool_order_bookkeeping_equals_physical_shuffle():
    """CPU: the index form (PoolOrder) names exactly the samples the reference's physical shuffle touches."""
    from ssl_b200.pool import PoolOrder
    q, b = 12, 3
    ref_ids = list(range(q))                 # physical content of the reference's queue: sample ids
    order = PoolOrder(q)
    phys = list(range(q))                    # content of OUR slots
    nxt = 100
    for _ in range(q // b):
        assert order.enqueue_slots(b) == list(range(order.ptr - b, order.ptr))
    assert order.full()
    torch.manual_seed(3)
    for step in range(20):
        idx = torch.randperm(q).tolist()
        ref_ids = [ref_ids[i] for i in idx]
        want_out = ref_ids[:b]
        new = list(range(nxt, nxt + b)); nxt += b
        ref_ids[:b] = new
        slots = order.exchange_slots(b, idx)
        got_out = [phys[s] for s in slots]
        for s, v in zip(slots, new):
            phys[s] = v
        assert got_out == want_out
        assert sorted(phys) == sorted(ref_ids)


def test_crop_numpy_branch_and_errors():
    """The NumPy branch (dataset side) slices like the reference; shape errors are the reference's."""
    from ssl_b200.pool import paired_random_crop_img_mask
    rng = np.random.default_rng(0)
    gt, lq, m = rng.random((32, 40, 3)), rng.random((8, 10, 3)), rng.random((32, 40))
    random.seed(5)
    g, l, mm = paired_random_crop_img_mask(gt, lq, m, 16, 4)
    random.seed(5)
    top, left = random.randint(0, 8 - 4), random.randint(0, 10 - 4)
    assert np.array_equal(l, lq[top:top + 4, left:left + 4]) and np.array_equal(g, gt[4 * top:4 * top + 16, 4 * left:4 * left + 16])
    assert np.array_equal(mm, m[4 * top:4 * top + 16, 4 * left:4 * left + 16])
    with pytest.raises(ValueError, match="Scale mismatches"):
        paired_random_crop_img_mask(gt, lq[:7], m, 16, 4)
    with pytest.raises(ValueError, match="smaller than patch size"):
        paired_random_crop_img_mask(gt, lq, m, 64, 4)


@pytest.mark.gpu
def test_pool_is_index_exact_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ssl_b200.pool import TrainingPairPool
    dev = torch.device("cuda:0")
    data = _batches(14)
    ref, ours = ReferencePool(6), TrainingPairPool(6)
    torch.manual_seed(11)
    want = [ref.step(*x) for x in data]
    torch.manual_seed(11)
    got = [ours(*(t.to(dev) for t in x)) for x in data]
    for i, (w, g) in enumerate(zip(want, got)):
        for a, b in zip(w, g):
            assert a.shape == b.shape, f"step {i}"
            assert torch.equal(a, b.cpu()), f"step {i}"
    # the quirk: once the pool is full the 1-channel mask comes back with the GT's three channels, all equal
    assert want[2][2].shape[1] == 1 and got[-1][2].shape[1] == 3
    assert torch.equal(got[-1][2][:, 0], got[-1][2][:, 2])
    # ... and such a mask feeds the loss unchanged (channel 0 is read)
    from ssl_b200 import ssl
    m3 = got[-1][2]
    sr = torch.rand(2, 3, 24, 20, device=dev).requires_grad_(True)
    l3 = ssl(sr, got[-1][1], m3, 7, 3)
    l1 = ssl(sr, got[-1][1], m3[:, :1].contiguous(), 7, 3)
    assert float(l3) == float(l1)


@pytest.mark.gpu
def test_crop_is_index_exact_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ssl_b200.pool import paired_random_crop_img_mask
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    gt, lq = torch.rand(3, 3, 96, 128, generator=g), torch.rand(3, 3, 24, 32, generator=g)
    m = (torch.rand(3, 1, 96, 128, generator=g) < 0.2).float()
    for seed in range(4):
        random.seed(seed)
        wg, wl, wm = reference_crop(gt, lq, m, 64, 4)
        random.seed(seed)
        gg, gl, gm = paired_random_crop_img_mask(gt.to(dev), lq.to(dev), m.to(dev), 64, 4)
        assert torch.equal(wg, gg.cpu()) and torch.equal(wl, gl.cpu()) and torch.equal(wm, gm.cpu())
    # list form (Use_sharpen: [gt, gt_usm], realesrganssl_model.py:302-305)
    random.seed(9)
    (g1, g2), l1, m1 = paired_random_crop_img_mask([gt.to(dev), (gt * 0.5).to(dev)], lq.to(dev), m.to(dev), 32, 4)
    assert torch.equal(g2, g1 * 0.5) and g1.shape[-1] == 32 and l1.shape[-1] == 8 and m1.shape[-1] == 32
