"""GPU: evaluation metrics, .npy ranking lists and the active-learning list producer vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import listmle_oracle as lo
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


def test_ordinal_error_bit_exact(cuda_device):
    from pldepth_b200 import metrics
    rs = np.random.RandomState(0)
    N, H, W = 5, 112, 96
    gt = rs.rand(N, H, W).astype(np.float32)
    op = (gt + 0.3 * rs.randn(N, H, W)).astype(np.float32)
    op[0] = gt[0]
    op[1] = -gt[1]
    err = metrics.ordinal_error(torch.from_numpy(op).to(cuda_device), torch.from_numpy(gt).to(cuda_device), (H, W), 2000)
    want = np.array([mo.ordinal_error(op[i], gt[i], (H, W), 2000) for i in range(N)], dtype=np.float32)
    assert np.array_equal(err.cpu().numpy(), want)
    assert err[0].item() == 0 and err[1].item() == 1
    assert abs(metrics.calc_err(torch.from_numpy(op).to(cuda_device), torch.from_numpy(gt).to(cuda_device), (H, W))
               - np.mean([mo.ordinal_error(op[i], gt[i], (H, W)) for i in range(N)])) < 1e-6


@pytest.mark.parametrize("list_size", [1, 7, 200, 1000])
def test_ndcg_matches_oracle(cuda_device, list_size):
    from pldepth_b200 import metrics
    rs = np.random.RandomState(list_size)
    N, H, W = 4, 64, 64
    gt = rs.rand(N, H, W).astype(np.float32)
    op = (rs.randn(N, H, W) * 2).astype(np.float32)
    got = metrics.calc_d(torch.from_numpy(op).to(cuda_device), torch.from_numpy(gt).to(cuda_device), (H, W), list_size)
    want = np.array([mo.calc_d(op[i], gt[i], (H, W), list_size) for i in range(N)])
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-6, atol=0)


def test_npy_lists_round_trip_into_the_loss(cuda_device, tmp_path):
    from pldepth_b200 import ops, rankings_io
    from tests.test_gpu_listmle import assert_close, make_problem
    B, H, W, K, R = 3, 20, 24, 5, 40
    y_true, pred = make_problem(B, H, W, K, R + 5, 9)
    paths = []
    for b in range(B):
        p = str(tmp_path / ("%d.npy" % (b + 1)))
        rankings_io.save_rankings(p, y_true[b])
        paths.append(p)
    yt = rankings_io.load_rankings(paths, cuda_device, rankings_per_image=R)
    assert tuple(yt.shape) == (B, R, K, 2)
    idx, depth = rankings_io.to_compact(yt)
    assert torch.equal(rankings_io.from_compact(idx, depth), yt)
    loss, _, grad, _ = ops.listmle_fwd_bwd(yt, torch.from_numpy(pred).to(cuda_device), B, K, 1.0 / (B * R))
    want_loss, want_grad, _ = lo.hourglass_nll(y_true[:, :R], pred, B, K)
    assert_close(loss.item(), want_loss, "loss")
    assert_close(grad.cpu().numpy(), want_grad, "gradient")
    with pytest.raises(ValueError):
        rankings_io.load_rankings(paths, cuda_device, rankings_per_image=R + 100)


def test_active_learning_oracle_lists(cuda_device):
    from pldepth_b200 import rankings_io
    rs = np.random.RandomState(3)
    H = W = 32
    gt = ((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W)
    pts = np.stack([rs.randint(0, H, 100), rs.randint(0, W, 100)], 1)
    shuffled = pts.copy()
    np.random.RandomState(5).shuffle(shuffled)
    want = mo.oracle_lists(gt, shuffled, 6, (H, W, 3))
    got = rankings_io.oracle_lists(torch.from_numpy(gt).to(cuda_device), pts, 6, (H, W, 3), rng=np.random.RandomState(5))
    assert np.array_equal(got.cpu().numpy(), want)
