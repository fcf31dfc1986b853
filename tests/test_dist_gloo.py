"""CPU, world_size 2 over gloo: the sharding / reduction logic of pldepth_b200.dist with the
oracle standing in for the CUDA step (tests may use the oracle; the product never does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import listmle_oracle as lo
from oracle import sampler_oracle as so
from pldepth_b200.dist import LossWindow, ShardedPLStep, shard_bounds

B, H, W, K, R = 5, 12, 10, 4, 30


def make_inputs():
    rs = np.random.RandomState(0)
    gt = np.stack([((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W) for _ in range(B)])
    mask = (rs.rand(B, H, W) > 0.2).astype(np.float32)
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    return gt, mask, pred


def oracle_local_step(gt, mask, pred, image_base, global_batch):
    """CPU stand-in with the FusedPLStep contract: per-image streams keyed by the global image index."""
    gt, mask, pred = gt.numpy(), mask.numpy(), pred.numpy()
    n = gt.shape[0]
    ranks = []
    for b in range(n):
        rng = np.random.RandomState(1000 + image_base + b)
        r, _ = so.sample_masked_rankings((H, W), mask[b], gt[b], R, 1.0, K, rng)
        ranks.append(r)
    y_true = np.stack(ranks)
    _, grad, per_list = lo.hourglass_nll(y_true, pred, n, K, global_lists=global_batch * R)
    return dict(loss_sum=torch.tensor([per_list.sum()], dtype=torch.float64),
                grad=torch.from_numpy(grad.astype(np.float32)), rankings=torch.from_numpy(y_true))


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gt, mask, pred = make_inputs()
        step = ShardedPLStep(K, R, B, local_step=oracle_local_step)
        lo_, hi_ = step.lo, step.hi
        out = step.run(torch.from_numpy(gt[lo_:hi_]), torch.from_numpy(mask[lo_:hi_]), torch.from_numpy(pred[lo_:hi_]))
        full = step.gather_grad(out["grad"])
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=out["loss"].numpy(), grad=full.numpy(),
                 lo=lo_, hi=hi_)
    finally:
        dist.destroy_process_group()


def window_worker(rank, world, port, out_dir):
    """Three steps with loss_every=2: one full window (reduced inside run) + one flushed partial window."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gt, mask, pred = make_inputs()
        step = ShardedPLStep(K, R, B, local_step=oracle_local_step, loss_every=2)
        lo_, hi_ = step.lo, step.hi
        got = []
        for i in range(3):
            p = pred * (1.0 + 0.5 * i)
            out = step.run(torch.from_numpy(gt[lo_:hi_]), torch.from_numpy(mask[lo_:hi_]), torch.from_numpy(p[lo_:hi_]))
            if i == 0:
                assert out["loss"] is None and out["losses"] is None
            if i == 1:
                assert out["losses"].numel() == 2
                got += out["losses"].tolist()
        got += step.flush().tolist()
        assert step.flush() is None
        np.save(os.path.join(out_dir, "win%d.npy" % rank), np.array(got))
    finally:
        dist.destroy_process_group()


def test_loss_window_reduces_once_per_n_steps(tmp_path):
    world = 2
    mp.spawn(window_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    gt, mask, pred = make_inputs()
    want = []
    for i in range(3):
        single = oracle_local_step(torch.from_numpy(gt), torch.from_numpy(mask), torch.from_numpy(pred * (1.0 + 0.5 * i)), 0, B)
        want.append(single["loss_sum"].item() / (B * R))
    w0, w1 = np.load(tmp_path / "win0.npy"), np.load(tmp_path / "win1.npy")
    assert np.array_equal(w0, w1) and w0.shape == (3,)
    assert np.allclose(w0, want, rtol=1e-6)


def test_loss_window_without_process_group():
    w = LossWindow(3, "cpu")
    for i in range(3):
        w.slot(i).fill_(float(i + 1))
        full = w.mark()
    assert full and w.reduce() is None and w.values().tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        LossWindow(0, "cpu")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_the_batch():
    for Bn in (1, 5, 8, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(Bn, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == Bn
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo_ for lo_, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    gt, mask, pred = make_inputs()
    single = oracle_local_step(torch.from_numpy(gt), torch.from_numpy(mask), torch.from_numpy(pred), 0, B)
    want_loss = single["loss_sum"].item() / (B * R)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 3, 3, 5)
    assert abs(float(r0["loss"][0]) - want_loss) < 1e-6 and float(r0["loss"][0]) == float(r1["loss"][0])
    assert np.array_equal(r0["grad"], r1["grad"])
    assert np.allclose(r0["grad"], single["grad"].numpy(), rtol=0, atol=1e-9)
