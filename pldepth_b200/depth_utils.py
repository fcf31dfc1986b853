"""Mirror of pldepth/data/depth_utils.py for the functions on the accelerated path."""
import torch

from . import ops
from ._lib import check


def get_depth_relation(depth1, depth2, threshold=None):
    """Ordinal relation of two depths: 1 (first deeper), -1, or 0 ("equal").  With a threshold the ratio
    (d1 + 1e-10) / (d2 + 1e-10) is compared with 1 + t and 1 / (1 + t)  (depth_utils.py:5-21)."""
    if threshold is None:
        return (depth1 > depth2) - (depth1 < depth2)
    ratio = (depth1 + 1e-10) / (depth2 + 1e-10)
    if ratio >= 1 + threshold:
        return 1
    if ratio <= 1 / (1 + threshold):
        return -1
    return 0


def prepare_fully_fledged_loss_input(labels, logits, batch_size, ranking_size, debug=False):
    """(selected_depths [B*R, K], reshaped_labels [B*R, K]) as depth_utils.py:39-61: the predictions gathered
    at the rankings' flat indices and the rankings' depths.  Device tensors in, device tensors out.  The fused
    loss does this gather inside its kernel; this standalone form serves callers that want the intermediates."""
    labels = ops.as_cuda(labels, torch.float32, "labels")
    logits = ops.as_cuda(logits, torch.float32, "logits")
    B, K = int(batch_size), int(ranking_size)
    if labels.numel() % (B * K * 2) != 0 or logits.numel() % B != 0:
        raise ValueError("labels / logits cannot be viewed as [B, -1, K, 2] / [B, -1]")
    R = labels.numel() // (B * K * 2)
    HW = logits.numel() // B
    ctx = ops._ctx(logits)
    with torch.cuda.device(logits.device):
        selected = torch.empty((B * R, K), dtype=torch.float32, device=logits.device)
        out_labels = torch.empty((B * R, K), dtype=torch.float32, device=logits.device)
        check(ctx.lib.pld_gather_predictions(ctx.handle, ops._p(labels), ops._p(logits), B, R, K, HW, ops._p(selected),
                                             ops._p(out_labels), ops._stream(logits.device)))
    if debug:
        print("point_coords:", labels.reshape(B, -1, K, 2)[..., 0].flatten()[:10])
        print("selected_depths:", selected.flatten()[:10])
        print("reshaped_labels:", out_labels.flatten()[:10])
    return selected, out_labels
