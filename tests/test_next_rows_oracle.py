"""CPU tests of the rows next to the hot path: the oracle restatements against the
golden vectors produced by the reference's own Python (tests/golden/make_golden_next.py)."""
import os

import numpy as np
import pytest

from oracle import next_rows as N

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def loss_gold():
    return np.load(os.path.join(GOLD, "loss_l1_ssim.npz"))


def test_window_matches_reference(loss_gold):
    # torch and numpy sum the 11 taps in different orders: 1 ulp
    np.testing.assert_allclose(N.gaussian_window(), loss_gold["window"], rtol=2e-7)
    from dmgs_b200.loss_utils import gaussian_window  # the product's host code uses the reference's torch expression
    assert np.array_equal(gaussian_window().numpy(), loss_gold["window"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_l1_ssim_oracle_matches_reference(loss_gold, tag):
    g = loss_gold
    l1, ss, dl1, dss = N.l1_ssim(g[f"img_{tag}"], g[f"gt_{tag}"])
    assert abs(l1 - g[f"l1_{tag}"]) <= 1e-6
    assert abs(ss - g[f"ssim_{tag}"]) <= 1e-6
    lam = 0.2
    assert abs((1 - lam) * l1 + lam * (1 - ss) - g[f"loss_{tag}"]) <= 1e-6
    assert np.array_equal(dl1.astype(np.float32) != 0, g[f"grad_l1_{tag}"] != 0)  # sign(0) = 0 on exact ties
    np.testing.assert_allclose(dl1, g[f"grad_l1_{tag}"], rtol=1e-5, atol=1e-10)
    ref = g[f"grad_ssim_{tag}"]
    assert np.abs(dss - ref).max() <= 1e-4 * np.abs(ref).max()
    tot = (1 - lam) * dl1 - lam * dss
    assert np.linalg.norm(tot - g[f"grad_{tag}"]) <= 1e-4 * np.linalg.norm(g[f"grad_{tag}"])


def test_ssim_batched_per_sample(loss_gold):
    g = loss_gold
    got = [np.mean([N.l1_ssim(g["img_batch"][b, c:c + 1], g["gt_batch"][b, c:c + 1])[1] for c in range(3)]) for b in range(2)]
    np.testing.assert_allclose(got, g["ssim_batch"], atol=1e-6)


@pytest.fixture(scope="module")
def fr_gold():
    return np.load(os.path.join(GOLD, "frustum.npz"))


def test_frustum_faces_match_reference(fr_gold):
    g = fr_gold
    m = N.in_frustum(g["proj"], g["verts"], faces=g["faces"])
    assert 0 < m.sum() < m.size
    assert np.array_equal(m, g["face_mask"])
    assert np.array_equal(g["faces"][m], g["faces_visible"])
    assert np.array_equal(N.in_frustum(g["proj"], g["verts"]), g["vert_mask"])


@pytest.mark.parametrize("pid,npc", [(-1, 1), (0, 2), (1, 2), (0, 4), (1, 4), (2, 4), (3, 4)])
def test_frustum_colmap_pieces_match_reference(fr_gold, pid, npc):
    g = fr_gold
    m = N.in_frustum(g["proj"], g["grid"], cube_len=float(g["cube_len"]), piece_id=pid, n_piece=npc)
    assert np.array_equal(m, g[f"grid_mask_{pid}_{npc}"])


def test_adam_oracle_matches_torch():
    g = np.load(os.path.join(GOLD, "adam.npz"))
    for k, lr in zip(g["names"], g["lrs"]):
        p = g[f"p0_{k}"]
        m, v = np.zeros_like(p), np.zeros_like(p)
        for t in range(int(g["steps"])):
            p, m, v = N.adam_step(p, g[f"g{t}_{k}"], m, v, t + 1, float(lr))
            # one Adam step moves a parameter by ~lr: the tolerance is relative to |p| + lr
            ref = g[f"p{t + 1}_{k}"]
            assert np.all(np.abs(p - ref) <= 2e-6 * (np.abs(ref) + float(lr)))
        assert np.abs(m - g[f"m_{k}"]).max() <= 2e-6 * np.abs(g[f"m_{k}"]).max()
        assert np.abs(v - g[f"v_{k}"]).max() <= 2e-6 * np.abs(g[f"v_{k}"]).max()
