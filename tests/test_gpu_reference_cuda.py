"""The reference's OWN CUDA operator (similarity.cu, compiled unmodified into oracle/_ref by
oracle/build_ref.py) as a second oracle: `_compute_similarity` / `_compute_similarity_backward`
(similarity.h:2-23) against the drop-in entry points of libssl_b200.so, the plane kernels and the
fp64 CPU oracle, on the reference's calling convention (reflect-padded image, int32 (row, col)
positions in padded coordinates, zero-initialised outputs; similaritywrapper.py:25-69).

Forward: 2e-6 relative (both sides sum the same 243 / 75 non-negative fp32 terms in a different order).
Backward: 1e-5 * max|grad| -- the reference scatters with fp32 atomicAdd (similarity.cu:124-128), so its
own result is order-noisy; both implementations are therefore also held to the fp64 oracle."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import build_ref
from oracle import ssl_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ssl_b200  # noqa: F401
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ref():
    lib = build_ref.load()
    if lib is None:
        pytest.skip("oracle/_ref/libsimilarity_ref.so is not built (python -m oracle.build_ref)")
    return lib


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


def _padded_inputs(dev, img, mask, ks):
    P = ks // 2
    img_pad = torch.nn.functional.pad(torch.from_numpy(img), (P, P, P, P), mode="reflect").contiguous().to(dev)
    pos = oracle.edge_positions(mask)
    pos_pad = torch.from_numpy(pos + P).int().contiguous().to(dev)
    return img_pad, pos, pos_pad


@pytest.mark.parametrize("name", ["configA_seed0", "k25w9_48x56", "k7w7_20x24"])
def test_forward_equals_reference_cuda(dev, ref, name):
    from ssl_b200 import _lib
    case = load_golden(name)
    ks, kw = int(case["ks"]), int(case["kw"])
    img, mask = case["sr"][0], case["mask"][0, 0]
    img_pad, pos, pos_pad = _padded_inputs(dev, img, mask, ks)
    mc = len(pos)
    c, hp, wp = img_pad.shape
    out_ref = torch.zeros(mc, ks, ks, device=dev)          # the reference accumulates into zeros (similaritywrapper.py:29)
    torch.cuda.synchronize()
    assert ref.ref_compute_similarity(_vp(img_pad), _vp(pos_pad), _vp(out_ref), mc, ks, kw, hp, wp, c) == 0
    torch.cuda.synchronize()
    out = torch.full((mc, ks, ks), float("nan"), device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("ssl_b200_compute_similarity", _vp(img_pad), _vp(pos_pad), _vp(out), mc, ks, kw, hp, wp, c, st)
    torch.cuda.synchronize()
    q64 = oracle.raw_distance(img.astype(np.float64), pos, ks, kw).reshape(mc, ks, ks)
    np.testing.assert_allclose(out_ref.cpu().numpy(), q64, rtol=2e-6, atol=1e-7)   # the reference vs the fp64 oracle
    np.testing.assert_allclose(out.cpu().numpy(), q64, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.cpu().numpy(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["configA_seed0", "k25w9_48x56"])
def test_plane_rows_equal_reference_cuda(dev, ref, name):
    """The kernels the benchmark runs (plane family, unpadded image + flat edge list) against similarity.cu."""
    import ssl_b200
    case = load_golden(name)
    ks, kw = int(case["ks"]), int(case["kw"])
    img, mask = case["sr"][0], case["mask"][0, 0]
    img_pad, pos, pos_pad = _padded_inputs(dev, img, mask, ks)
    mc = len(pos)
    c, hp, wp = img_pad.shape
    out_ref = torch.zeros(mc, ks, ks, device=dev)
    torch.cuda.synchronize()
    assert ref.ref_compute_similarity(_vp(img_pad), _vp(pos_pad), _vp(out_ref), mc, ks, kw, hp, wp, c) == 0
    torch.cuda.synchronize()
    el = ssl_b200.build_edge_list(torch.from_numpy(mask).view(1, 1, *mask.shape).to(dev))
    rows = ssl_b200.ssg_rows(torch.from_numpy(img).unsqueeze(0).to(dev), el, ks, kw, raw=True, path="plane")
    np.testing.assert_allclose(rows.cpu().numpy(), out_ref.reshape(mc, -1).cpu().numpy(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["configA_seed0", "k25w9_48x56"])
def test_backward_equals_reference_cuda(dev, ref, name):
    from ssl_b200 import _lib
    case = load_golden(name)
    ks, kw = int(case["ks"]), int(case["kw"])
    img, mask = case["sr"][0], case["mask"][0, 0]
    img_pad, pos, pos_pad = _padded_inputs(dev, img, mask, ks)
    mc = len(pos)
    c, hp, wp = img_pad.shape
    g = torch.Generator().manual_seed(5)
    grads = torch.randn(mc, ks * ks, generator=g).to(dev)
    gi_ref = torch.zeros_like(img_pad)                      # similaritywrapper.py:47
    torch.cuda.synchronize()
    assert ref.ref_compute_similarity_backward(_vp(img_pad), _vp(grads), _vp(pos_pad), _vp(gi_ref), mc, ks, kw, hp,
                                               wp, c) == 0
    torch.cuda.synchronize()
    gi = torch.zeros_like(img_pad)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("ssl_b200_compute_similarity_backward", _vp(img_pad), _vp(grads), _vp(pos_pad), _vp(gi), mc, ks, kw, hp,
              wp, c, st)
    torch.cuda.synchronize()
    P = ks // 2
    g64 = oracle.raw_distance_backward(img_pad.cpu().numpy().astype(np.float64), pos + P,
                                       grads.cpu().numpy().astype(np.float64), ks, kw)
    gmax = np.abs(g64).max()
    assert np.abs(gi_ref.cpu().numpy() - g64).max() <= 1e-5 * gmax     # the reference (atomics) vs fp64
    assert np.abs(gi.cpu().numpy() - g64).max() <= 1e-5 * gmax
    assert np.abs(gi.cpu().numpy() - gi_ref.cpu().numpy()).max() <= 1e-5 * gmax


def test_full_ssl_cuda_pipeline_equals_reference_cuda(dev, ref):
    """ssl_cuda of the reference (loss_util.py:231-244: op, /(C k_w^2), exp, normalise) rebuilt around the
    reference's CUDA op with torch ops, against similarity_map of this repo, forward and backward."""
    from ssl_b200 import similarity_map
    case = load_golden("k25w9_48x56")
    ks, kw, sigma = 25, 9, 0.004
    img, mask = case["sr"][0], case["mask"][0, 0]
    img_pad, pos, pos_pad = _padded_inputs(dev, img, mask, ks)
    mc = len(pos)
    c, hp, wp = img_pad.shape
    q = torch.zeros(mc, ks, ks, device=dev)
    torch.cuda.synchronize()
    assert ref.ref_compute_similarity(_vp(img_pad), _vp(pos_pad), _vp(q), mc, ks, kw, hp, wp, c) == 0
    torch.cuda.synchronize()
    s_ref = q / (c * kw * kw)
    s_ref = s_ref.reshape(1, mc, ks * ks)
    s_ref = torch.exp(-1 * s_ref / sigma)
    s_ref = 1 / (torch.sum(s_ref, dim=-1) + 1e-10).unsqueeze(-1) * s_ref
    x = torch.from_numpy(img).unsqueeze(0).to(dev).requires_grad_(True)
    s = similarity_map(img=x, mask=torch.from_numpy(mask).view(1, 1, *mask.shape).to(dev), ssl_mode="cuda",
                       kernel_size_search=ks, generalization=True, kernel_size_window=kw, sigma=sigma).getitem()
    assert s.shape == s_ref.shape
    err = (s.detach() - s_ref).abs() / s_ref.max(dim=-1, keepdim=True).values
    assert float(err.max()) <= 1e-5


@pytest.mark.parametrize("density,family", [(0.2, "ssg_plane"), (0.004, "ssg_point")])
def test_dropin_entries_pick_the_kernel_family_by_density(dev, ref, density, family):
    """similarity.h drop-ins: dense masks run on the plane kernels (deterministic, no atomics), sparse ones on the
    point kernels; both agree with the reference's CUDA op and with the fp64 oracle, forward and backward."""
    from ssl_b200 import _lib, synth
    ks, kw = 25, 9
    sr, _, mask = synth.make_case(1, 96, 72, seed=17, density=density)
    img, m = sr[0].numpy(), mask[0, 0].numpy()
    img_pad, pos, pos_pad = _padded_inputs(dev, img, m, ks)
    mc = len(pos)
    c, hp, wp = img_pad.shape
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out_ref = torch.zeros(mc, ks, ks, device=dev)
    torch.cuda.synchronize()
    assert ref.ref_compute_similarity(_vp(img_pad), _vp(pos_pad), _vp(out_ref), mc, ks, kw, hp, wp, c) == 0
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(6)
    grads = torch.randn(mc, ks * ks, generator=g).to(dev)
    gi_ref = torch.zeros_like(img_pad)
    assert ref.ref_compute_similarity_backward(_vp(img_pad), _vp(grads), _vp(pos_pad), _vp(gi_ref), mc, ks, kw, hp, wp,
                                               c) == 0
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    out = torch.full((mc, ks, ks), float("nan"), device=dev)
    _lib.call("ssl_b200_compute_similarity", _vp(img_pad), _vp(pos_pad), _vp(out), mc, ks, kw, hp, wp, c, st)
    gi = torch.full_like(img_pad, 0.25)                       # contributions are ADDED into the caller's buffer
    _lib.call("ssl_b200_compute_similarity_backward", _vp(img_pad), _vp(grads), _vp(pos_pad), _vp(gi), mc, ks, kw, hp,
              wp, c, st)
    torch.cuda.synchronize()
    stages = _lib.profile_read()
    _lib.profile_enable(False)
    assert family + "_fwd" in stages and family + "_bwd" in stages, stages
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.cpu().numpy(), rtol=2e-6, atol=1e-7)
    P = ks // 2
    g64 = oracle.raw_distance_backward(img_pad.cpu().numpy().astype(np.float64), pos + P,
                                       grads.cpu().numpy().astype(np.float64), ks, kw)
    gmax = np.abs(g64).max()
    assert np.abs(gi.cpu().numpy() - 0.25 - g64).max() <= 1e-5 * gmax
    assert np.abs(gi.cpu().numpy() - 0.25 - gi_ref.cpu().numpy()).max() <= 1e-5 * gmax
    if family == "ssg_plane":   # no atomics on this path: a second call gives the same bits
        gi2 = torch.full_like(img_pad, 0.25)
        _lib.call("ssl_b200_compute_similarity_backward", _vp(img_pad), _vp(grads), _vp(pos_pad), _vp(gi2), mc, ks, kw,
                  hp, wp, c, st)
        torch.cuda.synchronize()
        assert torch.equal(gi, gi2)
