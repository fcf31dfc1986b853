"""Independent batches in flight on separate streams: every lane has a private context (own scratch and lookup tables),
so concurrent steps must give exactly what the same steps give one after the other."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strategy,K", [("purely", 5), ("purely", 50), ("thresholded", 5), ("information", 7)])
def test_concurrent_lanes_equal_sequential_steps(cuda_device, strategy, K):
    from pldepth_b200._lib import Context
    from pldepth_b200.step import FusedPLStep
    dev = cuda_device
    B, H, W, R, lanes, steps = 4, 96, 80, 20000, 3, 9
    rs = np.random.RandomState(K)
    data = []
    for s in range(lanes):
        gt = torch.from_numpy(rs.rand(B, H, W).astype(np.float32)).to(dev)
        mask = torch.from_numpy((rs.rand(B, H, W) > 0.2).astype(np.float32)).to(dev)
        pred = torch.from_numpy(rs.randn(B, H, W, 1).astype(np.float32)).to(dev)
        data.append((gt, mask, pred))

    def make(lane, private):
        return FusedPLStep(K, R, seed=11, strategy=strategy, context=Context(dev.index or 0) if private else None,
                           first_step=lane << 20)

    # sequential reference on the thread's context, deterministic gradients for an exact comparison
    want = []
    seq = [make(l, False) for l in range(lanes)]
    Context.current(dev.index or 0).set_deterministic(True)
    try:
        for i in range(steps):
            l = i % lanes
            out = seq[l].run(*data[l])
            want.append((out["loss"].clone(), out["grad"].clone(), out["rankings"].clone()))
    finally:
        Context.current(dev.index or 0).set_deterministic(False)
    torch.cuda.synchronize(dev)

    par = [make(l, True) for l in range(lanes)]
    for p in par:
        p._ctx.set_deterministic(True)
    streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
    outs = [FusedPLStep.new_buffers(B, H, W, H, W, R, K, dev) for _ in range(steps)]
    for i in range(steps):
        l = i % lanes
        with torch.cuda.stream(streams[l]):
            par[l].run(*data[l], out=outs[i])
    torch.cuda.synchronize(dev)
    for p in par:
        p.check(dev)
    for i in range(steps):
        assert torch.equal(outs[i]["rankings"], want[i][2])
        assert torch.equal(outs[i]["loss"], want[i][0])
        assert torch.equal(outs[i]["grad"], want[i][1])
