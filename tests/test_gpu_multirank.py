"""Two real ranks (one process each) through the CUDA path: images sharded by `shard_range`, the 24-byte
[sum|d|, sumKL, n_rows] all-reduce, and the three parity modes of ssl_b200/dist.py against the
single-process run on the concatenated batch.

With two GPUs the ranks use NCCL on their own devices; on a one-GPU box both processes share cuda:0 and
the exchange runs over gloo (the reducer stages the three doubles through the host) -- same kernels,
same reduction, so the test runs wherever `-m gpu` runs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KS, KW = 25, 9
BATCH = 4


def _case():
    from ssl_b200 import synth
    sr, gt, mask = synth.make_case(BATCH, 96, 80, seed=11, density=0.114)
    mask[1, :, :40] = 0            # ranks get different row counts: per-rank means != global mean
    mask[2, :, :, ::2] = 0
    return sr, gt, mask


def _worker(rank, world, port, out, parity, bf16):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    backend = "nccl" if ndev >= world else "gloo"
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from ssl_b200 import ssl
    from ssl_b200.dist import shard_range
    sr, gt, mask = _case()
    if bf16:
        sr, gt = sr.bfloat16(), gt.bfloat16()
    idx = list(shard_range(BATCH, rank, world))
    x = sr[idx].to(dev).requires_grad_(True)
    loss = ssl(x, gt[idx].to(dev), mask[idx].to(dev), KS, KW, parity=parity, path="plane")
    loss.backward()
    torch.cuda.synchronize()
    out[rank] = (float(loss), x.grad.float().cpu().numpy(), backend, int(mask[idx].sum()))
    dist.barrier()
    dist.destroy_process_group()


def _run(parity, bf16=False):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out, parity, bf16), nprocs=2, join=True)
        return dict(out)


@pytest.fixture(scope="module")
def single():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ssl_b200 import ssl
    dev = torch.device("cuda:0")
    res = {}
    for bf16 in (False, True):
        sr, gt, mask = _case()
        if bf16:
            sr, gt = sr.bfloat16(), gt.bfloat16()
        x = sr.to(dev).requires_grad_(True)
        loss = ssl(x, gt.to(dev), mask.to(dev), KS, KW, path="plane")
        loss.backward()
        res[bf16] = (float(loss), x.grad.float().cpu().numpy())
    return res


def test_global_parity_equals_single_device(single):
    """parity="global": loss and gradient of two ranks == the single-device run on all four crops, 1e-6."""
    res = _run("global")
    loss1, grad1 = single[False]
    gmax = np.abs(grad1).max()
    for r in (0, 1):
        assert res[r][0] == pytest.approx(loss1, rel=1e-6)
        got = res[r][1]
        assert np.abs(got - grad1[2 * r:2 * r + 2]).max() <= 1e-6 * gmax
    assert res[0][3] != res[1][3]          # the shards really hold different numbers of rows


def test_global_parity_bf16_storage(single):
    """Config 3 storage: bf16 crops on both ranks; the returned gradient is bf16-rounded on both sides."""
    res = _run("global", bf16=True)
    loss1, grad1 = single[True]
    gmax = np.abs(grad1).max()
    for r in (0, 1):
        assert res[r][0] == pytest.approx(loss1, rel=1e-6)
        assert np.abs(res[r][1] - grad1[2 * r:2 * r + 2]).max() <= 2.0 ** -8 * gmax


def test_ddp_and_global_ddp_gradient_scale(single):
    """ "ddp" (the default) = mean over the rank's own rows, the reference's behaviour under DDP;
    "global_ddp" = global loss value with the gradient pre-multiplied by the world size, so that DDP's
    averaging of parameter gradients gives back the single-device gradient."""
    from ssl_b200 import ssl
    loss1, grad1 = single[False]
    gmax = np.abs(grad1).max()
    res = _run("global_ddp")
    for r in (0, 1):
        assert res[r][0] == pytest.approx(loss1, rel=1e-6)
        assert np.abs(res[r][1] - 2.0 * grad1[2 * r:2 * r + 2]).max() <= 2e-6 * gmax
    res = _run("ddp")
    dev = torch.device("cuda:0")
    sr, gt, mask = _case()
    for r in (0, 1):
        x = sr[2 * r:2 * r + 2].to(dev).requires_grad_(True)
        loss = ssl(x, gt[2 * r:2 * r + 2].to(dev), mask[2 * r:2 * r + 2].to(dev), KS, KW, path="plane")
        loss.backward()
        assert res[r][0] == pytest.approx(float(loss), rel=1e-6)
        assert np.abs(res[r][1] - x.grad.cpu().numpy()).max() <= 1e-6 * np.abs(x.grad.cpu().numpy()).max()
    assert res[0][0] != pytest.approx(loss1, rel=1e-3)   # per-rank means differ from the global mean here
