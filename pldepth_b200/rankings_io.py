"""On-disk / precomputed ranking lists feeding the same fused loss.

The reference dumps sampled lists as ``(n, K, 2) float32`` ``.npy`` files, one per image
(active_learning/offline_data.py:104-118: ``np.save(path + "lists/" + str(i) + ".npy", ele[1])``),
and its active-learning ``oracle()`` (active_learning_method.py:59-76) builds GT-sorted lists in the
same format from chosen pixels.  Both produce exactly the ``y_true`` rows the loss consumes.
"""
import numpy as np
import torch

from . import ops


def save_rankings(path, rankings):
    """One image's lists, reference format: float32 [n, K, 2] = (flat index, depth)."""
    arr = rankings.detach().cpu().numpy() if isinstance(rankings, torch.Tensor) else np.asarray(rankings)
    if arr.ndim != 3 or arr.shape[-1] != 2:
        raise ValueError("rankings must be [n, K, 2]")
    np.save(path, arr.astype(np.float32))


def load_rankings(paths, device, rankings_per_image=None):
    """Stack per-image ``.npy`` lists into ``y_true [B, R, K, 2]`` on ``device`` (R = the common /
    requested number of lists; longer files are truncated like ``result[:batch_size]``)."""
    arrs = [np.load(p).astype(np.float32) for p in paths]
    for a in arrs:
        if a.ndim != 3 or a.shape[-1] != 2:
            raise ValueError("not a ranking file: shape %r" % (a.shape,))
    R = min(a.shape[0] for a in arrs) if rankings_per_image is None else int(rankings_per_image)
    if any(a.shape[0] < R for a in arrs):
        raise ValueError("a file holds fewer than %d lists" % R)
    K = arrs[0].shape[1]
    if any(a.shape[1] != K for a in arrs):
        raise ValueError("ranking_size differs between files")
    return torch.from_numpy(np.stack([a[:R] for a in arrs])).to(device)


def to_compact(rankings):
    """(flat index int32 [..., K], depth float32 [..., K]) -- 8 bytes/point like the float pair, but with
    exact indices beyond 2^24 pixels and directly usable as gather indices."""
    return rankings[..., 0].to(torch.int32), rankings[..., 1].contiguous()


def from_compact(index, depth):
    return torch.stack([index.to(torch.float32), depth.to(torch.float32)], dim=-1)


def oracle_lists(gt, pos_xy, ranking_size, img_size=(224, 224, 3), shuffle=True, rng=None):
    """GT-ordered lists from chosen pixels, as active_learning_method.py:59-76 ``oracle()``:
    shuffle the (x, y) points, cut them into consecutive groups of ``ranking_size``, order each group by
    ground-truth depth descending, flat index = x * img_size[0] + y.  Like the reference, the loop
    ``range(0, N - K, K)`` leaves the final row(s) of the ``int(N / K)``-row buffer zero.

    gt: device tensor [H, W]; pos_xy: int array [N, 2].  Returns float32 [int(N/K), K, 2] on the device.
    """
    pos_xy = np.array(pos_xy, dtype=np.int64, copy=True)
    if shuffle:
        (np.random if rng is None else rng).shuffle(pos_xy)
    K = int(ranking_size)
    N = pos_xy.shape[0]
    rows = int(N / K)
    filled = len(range(0, N - K, K))
    gt = ops.as_cuda(gt, torch.float32, "gt")
    H, W = gt.shape[-2], gt.shape[-1]
    out = torch.zeros((rows, K, 2), dtype=torch.float32, device=gt.device)
    if filled == 0:
        return out
    pts = pos_xy[:filled * K]
    sel = torch.from_numpy((pts[:, 0] * W + pts[:, 1]).astype(np.int32)).to(gt.device)   # gt[x, y]
    n_valid = torch.tensor([-H * W], dtype=torch.int32, device=gt.device)                # identity table
    vf = torch.empty((1, 1), dtype=torch.int32, device=gt.device)
    lists = ops.sample_lists_fed(gt.reshape(1, H, W), vf, n_valid, K, sel.reshape(1, filled, K))[0]
    if int(img_size[0]) != W:        # the reference's flat index uses img_size[0] as the row stride
        r = torch.div(lists[..., 0].long(), W, rounding_mode="floor")
        c = lists[..., 0].long() - r * W
        lists = torch.stack([(r * int(img_size[0]) + c).float(), lists[..., 1]], dim=-1)
    out[:filled] = lists
    return out
