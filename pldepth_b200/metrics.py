"""Evaluation metrics of the reference on the GPU (pldepth/active_learning/metrics.py:60-120).

``ordinal_error`` / ``calc_d`` keep the reference's semantics including its fixed pixel samples
(``np.random.seed(10)`` / ``np.random.seed(69)`` + ``np.random.choice(..., replace=False)``): the
index sets are drawn once per image size with a private legacy ``RandomState`` (same stream, the
global NumPy state is left alone) and cached on the device; whole batches of maps are then scored
by one kernel (one CTA per image).
"""
import numpy as np
import torch

from . import ops
from ._lib import check

_PAIR_CACHE = {}
_SAMPLE_CACHE = {}


def _pairs(hw, num, device):
    key = (hw, num, str(device))
    if key not in _PAIR_CACHE:
        rs = np.random.RandomState(10)                                 # metrics.py:61
        idx = rs.choice(list(range(hw)), num * 2, replace=False)       # metrics.py:62
        i0, i1 = np.split(idx, 2)
        _PAIR_CACHE[key] = (torch.from_numpy(i0.astype(np.int32)).to(device),
                            torch.from_numpy(i1.astype(np.int32)).to(device))
    return _PAIR_CACHE[key]


def _samples(hw, n, device):
    key = (hw, n, str(device))
    if key not in _SAMPLE_CACHE:
        rs = np.random.RandomState(69)                                 # metrics.py:96
        ids = rs.choice(np.arange(hw), size=n, replace=False)          # metrics.py:97
        _SAMPLE_CACHE[key] = torch.from_numpy(ids.astype(np.int32)).to(device)
    return _SAMPLE_CACHE[key]


def _maps(x, name):
    x = ops.as_cuda(x, torch.float32, name)
    return x.reshape(x.shape[0], -1) if x.dim() > 2 else x.reshape(1, -1)


def ordinal_error(op, gt, imsize=(448, 448), num=5000):
    """op, gt: device tensors [N,H,W(,1)] (or one map) -> float32 [N] ordinal errors
    (metrics.py:60-70, ``imsize`` only fixes the index range like in the reference)."""
    op, gt = _maps(op, "op"), _maps(gt, "gt")
    hw = int(imsize[0]) * int(imsize[1])
    if op.shape[1] < hw or gt.shape[1] < hw:
        raise IndexError("maps are smaller than imsize")
    i0, i1 = _pairs(hw, num, op.device)
    ctx = ops._ctx(op)
    with torch.cuda.device(op.device):
        err = torch.empty(op.shape[0], dtype=torch.float32, device=op.device)
        check(ctx.lib.pld_ordinal_error(ctx.handle, ops._p(op), ops._p(gt), ops._p(i0), ops._p(i1), op.shape[0],
                                        op.shape[1], num, ops._p(err), ops._stream(op.device)))
    return err


def calc_err(preds, gts, img_size=(448, 448)):
    """metrics.py:73-80 with the model call factored out: mean ordinal error over the maps."""
    return float(ordinal_error(preds, gts, img_size).mean().item())


def calc_d(op, gt, imsize=(224, 224), list_size=200):
    """nDCG-style score per map (metrics.py:92-110): float32 [N]."""
    op, gt = _maps(op, "op"), _maps(gt, "gt")
    hw = int(imsize[0]) * int(imsize[1])
    if op.shape[1] < hw or gt.shape[1] < hw:
        raise IndexError("maps are smaller than imsize")
    ids = _samples(hw, list_size, op.device)
    ctx = ops._ctx(op)
    with torch.cuda.device(op.device):
        out = torch.empty(op.shape[0], dtype=torch.float32, device=op.device)
        check(ctx.lib.pld_ndcg(ctx.handle, ops._p(op), ops._p(gt), ops._p(ids), op.shape[0], op.shape[1], list_size,
                               ops._p(out), ops._stream(op.device)))
    return out


def dcg_metric(preds, gts, list_size=200, imsize=(224, 224)):
    """metrics.py:113-120 with the model call factored out."""
    return float(calc_d(preds, gts, imsize, list_size).mean().item())
