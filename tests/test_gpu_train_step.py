"""GPU: the sampler + PL loss as a differentiable training criterion (BASELINE config 4's integration,
pldepth/PLDepth.py:129-134 + pldepth/models/pl_hourglass.py:43-100 with the convolutions left to the framework)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import listmle_oracle as lo

pytestmark = pytest.mark.gpu


class TinyHourglass(nn.Module):
    """Two-level encoder-decoder with one skip, standing in for the reference's EfficientNet hourglass."""

    def __init__(self):
        super().__init__()
        self.e0 = nn.Conv2d(3, 8, 3, padding=1)
        self.e1 = nn.Conv2d(8, 16, 3, stride=2, padding=1)
        self.d0 = nn.Conv2d(16, 8, 3, padding=1)
        self.head = nn.Conv2d(16, 1, 3, padding=1)

    def forward(self, x):
        a = F.relu(self.e0(x))
        b = F.relu(self.e1(a))
        u = F.interpolate(F.relu(self.d0(b)), scale_factor=2, mode="bilinear", align_corners=False)
        return self.head(torch.cat([u, a], 1))


def batch(dev, B=4, H=32, W=48, seed=0, hole=True):
    from pldepth_b200 import synth
    gt = np.stack([synth.depth_map(H, W, 100 * seed + b) for b in range(B)])
    mask = np.stack([synth.valid_mask(H, W, 7 + b, 0.1 if hole else 0.0) for b in range(B)])
    rs = np.random.RandomState(seed)
    images = np.repeat(gt[:, None], 3, axis=1) + 0.05 * rs.randn(B, 3, H, W).astype(np.float32)
    return (torch.from_numpy(images.astype(np.float32)).to(dev), torch.from_numpy(gt).to(dev),
            torch.from_numpy(mask).to(dev))


@pytest.mark.parametrize("strategy,K", [("purely", 5), ("information", 5), ("thresholded", 20)])
def test_criterion_gradient_is_the_step_gradient_and_reaches_the_decoder(cuda_device, strategy, K):
    from pldepth_b200.losses import SampledHourglassNLL
    from pldepth_b200.step import FusedPLStep
    images, gt, mask = batch(cuda_device)
    torch.manual_seed(0)
    model = TinyHourglass().to(cuda_device)
    crit = SampledHourglassNLL(K, 300, strategy=strategy, seed=3, emit_rankings=True)
    pred = model(images)                                    # [B,1,H,W]
    pred.retain_grad()
    loss = crit(gt, mask, pred)
    loss.backward()
    # the same step, called directly with the same Philox stream, gives the gradient autograd received
    ref = FusedPLStep(K, 300, seed=3, strategy=strategy).run(gt, mask, pred.detach().contiguous())
    assert torch.equal(crit.last["rankings"], ref["rankings"])
    g_step = ref["grad"].reshape(pred.shape)
    assert torch.allclose(pred.grad, g_step, rtol=1e-5, atol=1e-9)
    assert abs(loss.item() - ref["loss"].item()) <= 1e-6 * abs(ref["loss"].item())
    # ... and equals the oracle's loss / gradient on the emitted lists
    B = gt.shape[0]
    want_loss, want_grad, _ = lo.hourglass_nll(crit.last["rankings"].cpu().numpy(),
                                               pred.detach().cpu().numpy().reshape(B, -1), B, K)
    assert abs(loss.item() - want_loss) <= 1e-5 * abs(want_loss)
    err = np.abs(pred.grad.cpu().numpy().reshape(B, -1) - want_grad).max() / np.abs(want_grad).max()
    assert err <= 1e-5, err
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if name == "head.bias":
            # the PL loss is invariant to a shift of all scores (every list's gradient sums to zero), so the bias of
            # the last layer receives exactly the rounding residue of that sum -- which may be 0.0
            assert p.grad.abs().sum().item() <= 1e-4 * pred.grad.abs().sum().item(), name
        else:
            assert p.grad.abs().sum().item() > 0, name
    # pixels outside the mask are never drawn: no gradient there
    assert (pred.grad[:, 0][mask == 0] == 0).all()


def test_training_reduces_the_loss_and_the_ordinal_error(cuda_device):
    """A few Adam steps (amsgrad, PLDepth.py:133) on one batch: fresh lists every step, the PL loss and the
    ordinal error of the prediction both fall."""
    from pldepth_b200 import metrics
    from pldepth_b200.losses import SampledHourglassNLL
    images, gt, mask = batch(cuda_device, B=4, H=32, W=48, seed=1, hole=False)
    torch.manual_seed(1)
    model = TinyHourglass().to(cuda_device)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, amsgrad=True)
    crit = SampledHourglassNLL(5, 500, strategy="thresholded", seed=11)
    with torch.no_grad():
        err0 = metrics.ordinal_error(model(images)[:, 0], gt, (32, 48), 500).mean().item()
    hist = []
    for _ in range(60):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = model(images)
        loss = crit(gt, mask, pred.float())
        loss.backward()
        opt.step()
        hist.append(loss.item())
    with torch.no_grad():
        err1 = metrics.ordinal_error(model(images)[:, 0], gt, (32, 48), 500).mean().item()
    assert np.isfinite(hist).all()
    assert np.mean(hist[-5:]) < 0.8 * np.mean(hist[:5]), (hist[:5], hist[-5:])
    assert err1 < err0 and err1 < 0.25, (err0, err1)
    assert crit.step.step_index == 60                      # one fresh Philox offset per call


def test_sharded_criteria_sum_to_the_global_step(cuda_device):
    """Two shards of one batch (global_batch / image_base, as two data-parallel ranks would run) reproduce the
    unsharded loss and gradient."""
    from pldepth_b200.losses import SampledHourglassNLL
    images, gt, mask = batch(cuda_device, B=4, seed=2)
    pred = torch.randn(4, 32, 48, 1, device=cuda_device, requires_grad=True)
    whole = SampledHourglassNLL(5, 200, seed=5)(gt, mask, pred)
    g_whole, = torch.autograd.grad(whole, pred)
    parts, grads = [], []
    for lo_, hi_ in ((0, 3), (3, 4)):
        p = pred[lo_:hi_].detach().clone().requires_grad_(True)
        l = SampledHourglassNLL(5, 200, seed=5, global_batch=4, image_base=lo_)(gt[lo_:hi_], mask[lo_:hi_], p)
        grads.append(torch.autograd.grad(l, p)[0])
        parts.append(l)
    assert abs((parts[0] + parts[1]).item() - whole.item()) <= 1e-6 * abs(whole.item())
    assert torch.allclose(torch.cat(grads), g_whole, rtol=1e-5, atol=1e-9)
