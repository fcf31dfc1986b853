"""CPU: the ListMLE oracle pinned by known answers, closed-form properties and an independent
fp64 autograd (torch) of the same TF-Ranking 0.3.1 expression.  (Parity vs the TF-Ranking
binary is unpinned: neither TensorFlow nor tensorflow_ranking can be installed here.)"""
import math

import numpy as np
import torch

from oracle import listmle_oracle as lo


def test_plackett_luce_known_answer():
    # Scores / labels of TF-Ranking's own published unit test (tensorflow_ranking/python/losses_test.py,
    # test_list_mle_loss, as recalled -- the package cannot be fetched here): expected per-list NLL
    # -(ln(3/6) + ln(2/3) + ln(1/1)) and -(ln(3/6) + ln(1/3) + ln(2/2)), mean 1.4451859.
    # P(order) = prod_i w_i / sum_{j>=i} w_j with w = exp(score)
    s = np.array([[0, np.log(3), np.log(2)], [0, np.log(2), np.log(3)]], np.float32)
    lab = np.array([[0, 2, 1], [1, 0, 2]], np.float32)
    nll, g = lo.listmle_per_list(lab, s)
    assert np.allclose(nll, [math.log(2) + math.log(1.5), math.log(2) + math.log(3)], rtol=1e-6)
    assert abs(nll.mean() - 1.4451859) < 1e-6
    assert np.allclose(g.sum(axis=1), 0, atol=1e-12)


def test_uniform_scores_give_log_factorial():
    for K in (1, 2, 5, 10, 50):
        lab = np.arange(K, 0, -1, dtype=np.float32)[None]
        nll, g = lo.listmle_per_list(lab, np.full((1, K), 0.37, np.float32))
        assert abs(nll[0] - math.lgamma(K + 1)) < 1e-9 * max(1, K)
    nll, g = lo.listmle_per_list(np.ones((1, 1), np.float32), np.ones((1, 1), np.float32))
    assert nll[0] == 0 and g[0, 0] == 0


def test_invariances():
    rs = np.random.RandomState(0)
    lab = rs.rand(6, 7).astype(np.float32)
    s = rs.randn(6, 7).astype(np.float32)
    nll, g = lo.listmle_per_list(lab, s)
    perm = rs.permutation(7)
    nll2, g2 = lo.listmle_per_list(lab[:, perm], s[:, perm])
    assert np.allclose(nll, nll2, rtol=1e-12) and np.allclose(g[:, perm], g2, rtol=1e-10, atol=1e-12)
    nll3, _ = lo.listmle_per_list(lab, s.astype(np.float64) + 3.0)
    assert np.allclose(nll, nll3, rtol=1e-6)


def torch_listmle(labels, scores):
    """Independent restatement with autograd: same ops as TF-Ranking's ListMLELoss."""
    labels = torch.as_tensor(labels, dtype=torch.float64)
    scores = torch.as_tensor(scores, dtype=torch.float64).clone().requires_grad_(True)
    valid = labels >= 0
    lab0 = torch.where(valid, labels, torch.zeros_like(labels))
    logits = torch.where(valid, scores, torch.full_like(scores, float(lo.LOG_EPS)))
    key = torch.where(valid, lab0, lab0.min(dim=1, keepdim=True).values - 1e-6)
    order = torch.argsort(key, dim=1, descending=True, stable=True)
    s = torch.gather(logits, 1, order)
    s = s - s.max(dim=1, keepdim=True).values
    sums = torch.flip(torch.cumsum(torch.flip(torch.exp(s), [1]), 1), [1])
    nll = (torch.log(sums) - s).sum(dim=1)
    nll.sum().backward()
    return nll.detach().numpy(), scores.grad.numpy()


def test_closed_form_gradient_equals_autograd():
    rs = np.random.RandomState(1)
    for K in (2, 5, 10, 50):
        lab = rs.rand(40, K).astype(np.float32)
        s = (rs.randn(40, K) * 2).astype(np.float32)
        nll, g = lo.listmle_per_list(lab, s)
        nll_t, g_t = torch_listmle(lab, s)
        assert np.allclose(nll, nll_t, rtol=1e-10)
        assert np.allclose(g, g_t, rtol=1e-9, atol=1e-12)


def test_invalid_labels_follow_tf_ranking():
    # negative labels: logit := log(1e-10), sorted last, no gradient; they still add log(#remaining)
    lab = np.array([[0.9, -1.0, 0.4, -1.0, 0.1]], np.float32)
    s = np.array([[0.3, 5.0, -0.2, 7.0, 0.8]], np.float32)
    nll, g = lo.listmle_per_list(lab, s)
    nll_t, g_t = torch_listmle(lab, s)
    assert np.allclose(nll, nll_t, rtol=1e-10) and np.allclose(g, g_t, atol=1e-12)
    assert g[0, 1] == 0 and g[0, 3] == 0
    nll_valid, _ = lo.listmle_per_list(lab[:, [0, 2, 4]], s[:, [0, 2, 4]])
    assert abs((nll[0] - nll_valid[0]) - math.log(2)) < 1e-6


def test_full_loss_mean_scatter_and_duplicates():
    B, H, W, K, R = 2, 4, 5, 3, 6
    rs = np.random.RandomState(2)
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    idx = rs.randint(0, H * W, size=(B, R, K))
    idx[0, 0] = [7, 7, 3]                       # duplicate pixel in one list
    depth = np.sort(rs.rand(B, R, K), axis=2)[:, :, ::-1]
    y_true = np.stack([idx.astype(np.float32), depth.astype(np.float32)], axis=-1)
    loss, grad, per = lo.hourglass_nll(y_true, pred, B, K)
    assert grad.shape == pred.shape and per.shape == (B * R,)
    assert abs(loss - per.mean()) < 1e-12
    assert abs(grad.sum()) < 1e-12              # per-list gradients sum to zero
    # finite differences on three pixels
    for (b, p) in ((0, 7), (1, 0), (0, 3)):
        d = np.zeros_like(pred)
        d.reshape(B, -1)[b, p] = 1e-3
        lp, _, _ = lo.hourglass_nll(y_true, pred + d, B, K)
        lm, _, _ = lo.hourglass_nll(y_true, pred - d, B, K)
        assert abs((lp - lm) / 2e-3 - grad.reshape(B, -1)[b, p]) < 1e-5
    ls, gs, _ = lo.hourglass_nll(y_true, pred, B, K, reduction="sum")
    assert abs(ls - loss * B * R) < 1e-9 and np.allclose(gs, grad * B * R)
    lg, gg, _ = lo.hourglass_nll(y_true, pred, B, K, global_lists=4 * B * R)
    assert abs(lg - loss / 4) < 1e-12 and np.allclose(gg, grad / 4)


# ---- pin against the real TF-Ranking binary (files exist only after tools/pin_tfranking.py has run somewhere with
# TensorFlow + tensorflow_ranking==0.3.1; until then the oracle stays "parity unpinned") -------------------------------
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
TFR_GOLDEN = sorted(glob.glob(os.path.join(_HERE, "golden", "listmle_*.npz")))


def _pin_module():
    spec = importlib.util.spec_from_file_location("pin_tfranking", os.path.join(_HERE, "..", "tools", "pin_tfranking.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pin_generator_cases_are_deterministic_and_tie_free():
    """The golden generator's inputs (NumPy part, no TensorFlow needed): reproducible, labels distinct per list except
    the injected invalid ones, indices inside the map -- and the oracle runs on every case."""
    pin = _pin_module()
    a, b = pin.make_cases(), pin.make_cases()
    assert list(a) == list(b) and len(a) >= 8
    for name, c in a.items():
        assert np.array_equal(c["y_true"], b[name]["y_true"]) and np.array_equal(c["y_pred"], b[name]["y_pred"])
        B, K = c["batch_size"], c["ranking_size"]
        y = c["y_true"].reshape(-1, K, 2)
        assert y[..., 0].min() >= 0 and y[..., 0].max() < c["y_pred"].size // B
        for row in y[..., 1]:
            v = row[row >= 0]
            assert len(np.unique(v)) == len(v), name
        loss, grad, per_list = lo.hourglass_nll(c["y_true"], c["y_pred"], B, K)
        assert np.isfinite(loss) and np.isfinite(grad).all() and per_list.shape[0] == y.shape[0]


import pytest  # noqa: E402


@pytest.mark.skipif(not TFR_GOLDEN, reason="no tests/golden/listmle_*.npz yet (run tools/pin_tfranking.py where TF exists)")
@pytest.mark.parametrize("path", TFR_GOLDEN, ids=[os.path.basename(p)[8:-4] for p in TFR_GOLDEN])
def test_oracle_matches_tfranking_goldens(path):
    g = np.load(path)
    B, K = int(g["batch_size"]), int(g["ranking_size"])
    loss, grad, per_list = lo.hourglass_nll(g["y_true"], g["y_pred"], B, K, dtype=np.float32)
    assert abs(loss - float(g["loss"])) <= 1e-6 * max(abs(float(g["loss"])), 1e-12)
    assert np.abs(per_list - g["per_list"]).max() <= 1e-6 * np.abs(g["per_list"]).max()
    assert np.abs(grad - g["grad"]).max() <= 1e-6 * np.abs(g["grad"]).max()
