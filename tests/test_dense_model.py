"""The plane formulation (tests/dense_model.py == what ssg_plane_*.cuh computes) against the oracle, on the CPU."""
import numpy as np
import pytest

import dense_model as dm
from oracle import ssl_oracle as oracle


@pytest.mark.parametrize("h,w,ks,kw", [(12, 14, 7, 3), (10, 11, 7, 5), (9, 9, 5, 5), (12, 10, 5, 1)])
def test_plane_algebra_matches_oracle(h, w, ks, kw):
    rng = np.random.default_rng(h * 100 + w)
    img = rng.random((3, h, w))
    mask = (rng.random((h, w)) < 0.15).astype(np.float32)
    mask[0, 0] = mask[h - 1, w - 1] = mask[0, w // 2] = mask[h // 2, 0] = 1   # reflect corners / borders
    pos = oracle.edge_positions(mask)
    q_ref = oracle.raw_distance(img, pos, ks, kw)
    np.testing.assert_allclose(dm.forward_planes(img, pos, ks, kw), q_ref, rtol=0, atol=1e-12)
    gq = rng.standard_normal(q_ref.shape)
    g_ref = oracle.raw_distance_backward(img, pos, gq, ks, kw)
    np.testing.assert_allclose(dm.backward_planes(img, pos, gq, ks, kw), g_ref, rtol=0, atol=1e-10)


def test_clip_ranges():
    # A(t) is the full window for |t| <= P-K and shrinks by one per step beyond it
    P, K = 12, 4
    assert [dm.hi(t, P, K) - dm.lo(t, P, K) + 1 for t in range(-P, P + 1)] == \
        [5, 6, 7, 8] + [9] * 17 + [8, 7, 6, 5]
