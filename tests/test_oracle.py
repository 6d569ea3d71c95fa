"""The CPU oracle against the golden vectors produced by the real reference (oracle/make_golden.py).

This is what pins the oracle: the reference has no tests of its own for the SSG path.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ssl_oracle as oracle


def _params(case):
    return int(case["ks"]), int(case["kw"]), float(case["sigma"]), bool(case["gen"])


def _rows_c(case, key, dtype):
    ks, kw, sigma, gen = _params(case)
    img, mask = case[key].astype(dtype), case["mask"]
    out = []
    for i in range(img.shape[0]):
        if mask[i].sum() == 0:
            continue
        # a multi-channel mask selects each pixel once per channel, channel-major (loss_util.py:195-199)
        for ch in range(mask.shape[1]):
            out.append(oracle.rows(img[i], mask[i, ch], ks, kw, sigma, gen))
    return np.concatenate(out, 0)


def test_c_oracle_rows_f64(golden_case):
    for key, ref in (("sr", "rows_sr_f64"), ("gt", "rows_gt_f64")):
        got = _rows_c(golden_case, key, np.float64)
        assert got.shape == golden_case[ref].shape
        np.testing.assert_allclose(got, golden_case[ref], rtol=1e-11, atol=1e-300)


def test_c_oracle_rows_f32(golden_case):
    # fp32 loops vs the reference's fp32 torch ops: only the summation order differs
    for key, ref in (("sr", "rows_sr_f32"), ("gt", "rows_gt_f32")):
        got = _rows_c(golden_case, key, np.float32)
        want = golden_case[ref]
        scale = want.max(axis=1, keepdims=True)
        assert np.abs(got - want).max() <= 1e-5 * scale.max()
        assert (np.abs(got - want) <= 1e-5 * scale + 1e-30).all()


def test_c_oracle_loss_and_grad_f64(golden_case):
    ks, kw, sigma, gen = _params(golden_case)
    mask = golden_case["mask"]
    reps = mask.shape[1]  # tripled rows leave the mean unchanged and triple nothing in the gradient
    l1, kl, grad, n = oracle.loss_and_grad(golden_case["sr"].astype(np.float64), golden_case["gt"].astype(np.float64),
                                           mask, ks, kw, sigma, gen, loss_weight=1.0)
    assert n * reps == golden_case["rows_sr_f64"].shape[0]
    assert l1 == pytest.approx(float(golden_case["l1_f64"]), rel=1e-10)
    g_ref = golden_case["grad_l1_f64"]
    assert np.abs(grad - g_ref).max() <= 1e-9 * np.abs(g_ref).max()
    # KL term on the same rows (basic_loss.py:269-282)
    _, kl, grad_kl, _ = oracle.loss_and_grad(golden_case["sr"].astype(np.float64),
                                             golden_case["gt"].astype(np.float64), mask, ks, kw, sigma, gen,
                                             loss_weight=0.0, kl_weight=1.0)
    assert kl == pytest.approx(float(golden_case["kl_f64"]), rel=1e-9)
    gk_ref = golden_case["grad_kl_f64"]
    assert np.abs(grad_kl - gk_ref).max() <= 1e-8 * np.abs(gk_ref).max()


def test_c_oracle_loss_and_grad_f32(golden_case):
    ks, kw, sigma, gen = _params(golden_case)
    l1, _, grad, _ = oracle.loss_and_grad(golden_case["sr"], golden_case["gt"], golden_case["mask"], ks, kw, sigma,
                                          gen, loss_weight=1.0)
    assert l1 == pytest.approx(float(golden_case["l1_f64"]), rel=1e-5)
    g_ref = golden_case["grad_l1_f64"]
    assert np.abs(grad - g_ref).max() <= 2e-5 * np.abs(g_ref).max()


def test_torch_port_matches_reference_bitwise_f32(golden_case):
    """ssl_pytorch_port runs the same torch ops in the same order as the reference => same bits."""
    ks, kw, sigma, gen = _params(golden_case)
    sr, mask = torch.from_numpy(golden_case["sr"]), torch.from_numpy(golden_case["mask"])
    out = []
    for i in range(sr.shape[0]):
        if mask[i].sum() == 0:
            continue
        out.append(oracle.ssl_pytorch_port(sr[i:i + 1], mask[i:i + 1], ks, kw, sigma, gen))
    got = torch.cat(out, 1)[0].numpy()
    np.testing.assert_array_equal(got, golden_case["rows_sr_f32"])


def test_torch_port_chunked_equals_unchunked():
    case = load_golden("configA_seed0")
    ks, kw, sigma, gen = _params(case)
    sr, mask = torch.from_numpy(case["sr"]), torch.from_numpy(case["mask"])
    a = oracle.ssl_pytorch_port(sr, mask, ks, kw, sigma, gen)
    b = oracle.ssl_pytorch_port_chunked(sr, mask, max_px=17, kernel_size_search=ks, kernel_size_window=kw,
                                        sigma=sigma, generalization=gen)
    np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-6, atol=0)


def test_torch_step_port_matches_golden():
    case = load_golden("configA_batch2_single_pixel")
    ks, kw, sigma, gen = _params(case)
    loss, grad, n = oracle.ssl_step_pytorch_port(torch.from_numpy(case["sr"]), torch.from_numpy(case["gt"]),
                                                 torch.from_numpy(case["mask"]), ks, kw, sigma, gen, max_px=20)
    assert n == case["rows_sr_f32"].shape[0]
    assert loss == pytest.approx(float(case["l1_f64"]), rel=2e-5)
    g_ref = case["grad_l1_f64"]
    assert np.abs(grad.numpy() - g_ref).max() <= 2e-5 * np.abs(g_ref).max()


def test_laplacian_mask_matches_cv2_fixture():
    fx = load_golden("laplacian_mask")
    got = oracle.laplacian_mask(fx["rgb"], float(fx["threshold"]))
    np.testing.assert_array_equal(got, fx["mask"])


def test_edge_positions_row_major():
    m = np.zeros((5, 6), np.float32)
    m[3, 1] = m[0, 4] = m[3, 0] = 1
    m[2, 2] = 0.5  # compared with == 1 exactly (loss_util.py:196)
    np.testing.assert_array_equal(oracle.edge_positions(m), [[0, 4], [3, 0], [3, 1]])


def test_stride_mask_rule():
    sm = oracle.stride_mask(7, 8, 3)
    ref = torch.eye(3).repeat(3, 3)[:7, :8].numpy()  # realesrganssl_model.py:64-72
    np.testing.assert_array_equal(sm, ref)
