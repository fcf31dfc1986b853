"""CPU: metric oracles against the live reference module where it is importable."""
import numpy as np
import pytest

from oracle import metrics_oracle as mo


def test_ordinal_error_properties():
    rs = np.random.RandomState(0)
    gt = rs.rand(64, 64).astype(np.float32)
    assert mo.ordinal_error(gt, gt, (64, 64), 500) == 0
    assert mo.ordinal_error(-gt, gt, (64, 64), 500) == 1
    e = mo.ordinal_error(rs.rand(64, 64).astype(np.float32), gt, (64, 64), 1000)
    assert 0.4 < e < 0.6


def test_calc_d_matches_cv2_normalisation():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(1)
    op = (rs.randn(224, 224) * 3 + 1).astype(np.float32)
    gt = rs.rand(224, 224).astype(np.float32)
    a = mo.calc_d(op, gt)
    b = mo.calc_d(op, gt, normalize=lambda x: cv2.normalize(x, None, 0, 1, cv2.NORM_MINMAX))
    assert abs(a - b) <= 1e-6 * abs(b)
    assert abs(mo.calc_d(gt, gt) - 1.0) < 1e-5        # min-max normalising gt itself barely moves it


def test_oracle_lists_quirk_last_row_zero():
    rs = np.random.RandomState(2)
    gt = rs.rand(16, 16).astype(np.float32)
    pts = np.stack([rs.randint(0, 16, 12), rs.randint(0, 16, 12)], 1)
    out = mo.oracle_lists(gt, pts, 6, (16, 16, 3))
    assert out.shape == (2, 6, 2) and (out[1] == 0).all() and (np.diff(out[0, :, 1]) <= 0).all()
