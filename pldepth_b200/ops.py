"""Functional wrappers over the C ABI for tensors that live on a CUDA device.

Every wrapper runs on the calling thread's context (``Context.current``) unless a private one is passed as
``ctx=`` (the staged helpers a step object with its own context uses).

PyTorch is used only as plumbing: it owns device memory and the current stream.  Any object
exporting ``__dlpack__`` (e.g. a TensorFlow tensor via ``tf.experimental.dlpack``) is
accepted and viewed zero-copy.  Every function enqueues on ``torch.cuda.current_stream()``
and returns without synchronising unless stated.
"""
import ctypes

import torch

from . import _lib
from ._lib import Context, check, c_void_p


def as_cuda(x, dtype, name):
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            raise TypeError("%s: expected a CUDA tensor / DLPack exporter, got %r" % (name, type(x)))
    if not x.is_cuda:
        raise _lib.PLDError("%s must live on a CUDA device (no CPU fallback)" % name)
    if x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_contiguous():
        x = x.contiguous()
    return x


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)


def _stream(dev):
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ctx(t):
    return Context.current(t.device.index if t.device.index is not None else torch.cuda.current_device())


def mask_compact(mask, H, W, ctx=None):
    """mask f32[B,Hm,Wm] -> (valid_flat i32[B,Hm*Wm], n_valid i32[B]).  sampling.py:124-135.
    A negative n_valid[b] marks the identity table of a fully valid image-resolution mask (row
    not written; M = -n_valid[b]) -- see include/pldepth_b200.h."""
    mask = as_cuda(mask, torch.float32, "mask")
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    if mask.dim() == 4 and mask.shape[-1] == 1:
        mask = mask[..., 0].contiguous()
    B, Hm, Wm = mask.shape
    ctx = ctx if ctx is not None else _ctx(mask)
    with torch.cuda.device(mask.device):
        valid_flat = torch.empty((B, Hm * Wm), dtype=torch.int32, device=mask.device)
        n_valid = torch.empty((B,), dtype=torch.int32, device=mask.device)
        check(ctx.lib.pld_mask_compact(ctx.handle, _p(mask), B, Hm, Wm, int(H), int(W), _p(valid_flat),
                                       _p(n_valid), _stream(mask.device)))
    return valid_flat, n_valid


def _gt2d(gt):
    gt = as_cuda(gt, torch.float32, "gt")
    if gt.dim() == 4 and gt.shape[-1] == 1:
        gt = gt[..., 0]
    if gt.dim() == 3:
        gt = gt.reshape(gt.shape[0], -1)
    if gt.dim() != 2:
        raise ValueError("gt must be [B,H,W], [B,H,W,1] or [B,HW]")
    return gt.contiguous()


def sample_lists_philox(gt, valid_flat, n_valid, K, n, seed, offset=0, image_base=0, want_sel=False,
                        want_rankings=True, ctx=None):
    gt = _gt2d(gt)
    B, HW = gt.shape
    ctx = ctx if ctx is not None else _ctx(gt)
    with torch.cuda.device(gt.device):
        rankings = torch.empty((B, n, K, 2), dtype=torch.float32, device=gt.device) if want_rankings else None
        sel = torch.empty((B, n, K), dtype=torch.int32, device=gt.device) if want_sel else None
        check(ctx.lib.pld_sample_lists_philox(ctx.handle, _p(gt), _p(valid_flat), _p(n_valid), B, HW,
                                              valid_flat.shape[1], int(K), int(n), int(seed), int(offset),
                                              int(image_base), _p(rankings), _p(sel), _stream(gt.device)))
    return rankings, sel


def sample_lists_fed(gt, valid_flat, n_valid, K, sel):
    gt = _gt2d(gt)
    B, HW = gt.shape
    sel = as_cuda(sel, torch.int32, "sel").reshape(B, -1, K)
    n = sel.shape[1]
    ctx = _ctx(gt)
    with torch.cuda.device(gt.device):
        rankings = torch.empty((B, n, K, 2), dtype=torch.float32, device=gt.device)
        check(ctx.lib.pld_sample_lists_fed(ctx.handle, _p(gt), _p(valid_flat), _p(n_valid), B, HW,
                                           valid_flat.shape[1], int(K), int(n), _p(sel), _p(rankings),
                                           _stream(gt.device)))
    return rankings


def sample_lists_mt(gt, valid_flat, n_valid, K, n, raw, consumed):
    """raw u32 words (int32-viewed tensor is fine), consumed i64[1] (device, in/out).
    Returns (rankings, sel).  Caller checks ``Context.raise_on_status`` after syncing."""
    gt = _gt2d(gt)
    B, HW = gt.shape
    ctx = _ctx(gt)
    with torch.cuda.device(gt.device):
        rankings = torch.empty((B, n, K, 2), dtype=torch.float32, device=gt.device)
        sel = torch.empty((B, n, K), dtype=torch.int32, device=gt.device)
        check(ctx.lib.pld_sample_lists_mt(ctx.handle, _p(gt), _p(valid_flat), _p(n_valid), B, HW,
                                          valid_flat.shape[1], int(K), int(n), _p(raw), int(raw.numel()),
                                          _p(consumed), _p(rankings), _p(sel), _stream(gt.device)))
    return rankings, sel


def mt19937_init(seed, device):
    ctx = Context.current(torch.device(device).index or 0)
    with torch.cuda.device(device):
        state = torch.empty(624, dtype=torch.int32, device=device)
        pos = torch.empty(1, dtype=torch.int32, device=device)
        check(ctx.lib.pld_mt19937_init(ctx.handle, int(seed) & 0xFFFFFFFF, _p(state), _p(pos), _stream(device)))
    return state, pos


def mt19937_generate(state, pos, n):
    ctx = _ctx(state)
    with torch.cuda.device(state.device):
        out = torch.empty(int(n), dtype=torch.int32, device=state.device)
        check(ctx.lib.pld_mt19937_generate(ctx.handle, _p(state), _p(pos), _p(out), int(n), _stream(state.device)))
    return out


def gt_minmax(gt, ctx=None):
    gt = _gt2d(gt)
    B, HW = gt.shape
    ctx = ctx if ctx is not None else _ctx(gt)
    with torch.cuda.device(gt.device):
        out = torch.empty((B, 2), dtype=torch.float32, device=gt.device)
        check(ctx.lib.pld_gt_minmax(ctx.handle, _p(gt), B, HW, _p(out), _stream(gt.device)))
    return out


def score_lists(rankings, strategy, threshold=0.03, equality_penalty=-1000, promotion="nep50", minmax=None,
                ctx=None):
    rankings = as_cuda(rankings, torch.float32, "rankings")
    B, n, K, _ = rankings.shape
    ctx = ctx if ctx is not None else _ctx(rankings)
    with torch.cuda.device(rankings.device):
        scores = torch.empty((B, n), dtype=torch.float64, device=rankings.device)
        check(ctx.lib.pld_score_lists(ctx.handle, _p(rankings), _p(minmax), B, n, K, _lib.STRATEGY[strategy],
                                      float(threshold), float(equality_penalty), _lib.PROMOTION[promotion],
                                      _p(scores), _stream(rankings.device)))
    return scores


def select_top(scores, rankings, R, want_order=False, ctx=None):
    rankings = as_cuda(rankings, torch.float32, "rankings")
    scores = as_cuda(scores, torch.float64, "scores")
    B, n, K, _ = rankings.shape
    ctx = ctx if ctx is not None else _ctx(rankings)
    with torch.cuda.device(rankings.device):
        out = torch.empty((B, R, K, 2), dtype=torch.float32, device=rankings.device)
        order = torch.empty((B, R), dtype=torch.int32, device=rankings.device) if want_order else None
        check(ctx.lib.pld_select_top(ctx.handle, _p(scores), _p(rankings), B, n, K, int(R), _p(out), _p(order),
                                     _stream(rankings.device)))
    return out, order


def listmle_fwd_bwd(rankings, pred, B, K, scale, want_grad=True, want_per_list=False, grad_out=None,
                    accumulate=False, ctx=None):
    """rankings f32[B,R,K,2] (any shape reshapable to it), pred f32[B,...].
    Returns (loss f32[1], loss_sum f64[1], grad like pred | None, per_list f32[B*R] | None)."""
    rankings = as_cuda(rankings, torch.float32, "y_true")
    pred = as_cuda(pred, torch.float32, "y_pred")
    if rankings.numel() % (B * K * 2) != 0:
        raise ValueError("y_true with %d elements cannot be viewed as [B=%d, -1, K=%d, 2]" %
                         (rankings.numel(), B, K))
    R = rankings.numel() // (B * K * 2)
    if pred.numel() % B != 0:
        raise ValueError("y_pred cannot be viewed as [B, -1]")
    HW = pred.numel() // B
    ctx = ctx if ctx is not None else _ctx(pred)
    dev = pred.device
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        loss_sum = torch.empty(1, dtype=torch.float64, device=dev)
        grad = None
        if want_grad:
            grad = grad_out if grad_out is not None else torch.empty_like(pred)
        per_list = torch.empty(B * R, dtype=torch.float32, device=dev) if want_per_list else None
        check(ctx.lib.pld_listmle_fwd_bwd(ctx.handle, _p(rankings), _p(pred), B, R, int(K), HW, float(scale),
                                          _p(loss), _p(loss_sum), _p(per_list), _p(grad),
                                          1 if accumulate else 0, _stream(dev)))
    return loss, loss_sum, grad, per_list


def fused_sample_loss_bwd(gt, valid_flat, n_valid, pred, K, n, seed, offset=0, image_base=0, scale=None,
                          want_rankings=True, want_grad=True, want_per_list=False, rankings_out=None,
                          grad_out=None, accumulate=False):
    """One launch: Philox draws -> order by gt -> (rankings) -> gather pred -> NLL -> grad.
    Returns (loss, loss_sum, grad, rankings, per_list)."""
    gt = _gt2d(gt)
    pred = as_cuda(pred, torch.float32, "pred")
    B, HW = gt.shape
    if pred.numel() != B * HW:
        raise ValueError("pred and gt disagree on shape")
    if scale is None:
        scale = 1.0 / float(B * n)
    ctx = _ctx(gt)
    dev = gt.device
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        loss_sum = torch.empty(1, dtype=torch.float64, device=dev)
        rankings = None
        if want_rankings:
            rankings = rankings_out if rankings_out is not None else torch.empty((B, n, K, 2), dtype=torch.float32,
                                                                                 device=dev)
        grad = None
        if want_grad:
            grad = grad_out if grad_out is not None else torch.empty_like(pred)
        per_list = torch.empty(B * n, dtype=torch.float32, device=dev) if want_per_list else None
        check(ctx.lib.pld_fused_sample_loss_bwd(ctx.handle, _p(gt), _p(valid_flat), _p(n_valid), _p(pred), B, HW,
                                                valid_flat.shape[1], int(K), int(n), int(seed), int(offset),
                                                int(image_base), float(scale), _p(rankings), _p(loss),
                                                _p(loss_sum), _p(per_list), _p(grad), 1 if accumulate else 0,
                                                _stream(dev)))
    return loss, loss_sum, grad, rankings, per_list


def fused_step(mask, gt, pred, K, n, seed, offset=0, image_base=0, scale=None, want_rankings=True, want_grad=True,
               want_per_list=False):
    """pld_fused_step: mask f32[B,Hm,Wm], gt f32[B,H,W], pred f32[B,H,W(,1)] -> (loss, loss_sum, grad,
    rankings, per_list, n_valid).  ranking_size <= 16."""
    mask = as_cuda(mask, torch.float32, "mask")
    gt3 = as_cuda(gt, torch.float32, "gt")
    if gt3.dim() == 4:
        gt3 = gt3[..., 0].contiguous()
    pred = as_cuda(pred, torch.float32, "pred")
    B, H, W = gt3.shape
    Hm, Wm = mask.shape[1], mask.shape[2]
    if scale is None:
        scale = 1.0 / float(B * n)
    ctx = _ctx(gt3)
    dev = gt3.device
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        loss_sum = torch.empty(1, dtype=torch.float64, device=dev)
        n_valid = torch.empty(B, dtype=torch.int32, device=dev)
        rankings = torch.empty((B, n, K, 2), dtype=torch.float32, device=dev) if want_rankings else None
        grad = torch.empty_like(pred) if want_grad else None
        per_list = torch.empty(B * n, dtype=torch.float32, device=dev) if want_per_list else None
        check(ctx.lib.pld_fused_step(ctx.handle, _p(mask), _p(gt3), _p(pred), B, Hm, Wm, H, W, int(K), int(n), int(seed),
                                     int(offset), int(image_base), float(scale), _p(n_valid), _p(rankings), _p(loss),
                                     _p(loss_sum), _p(per_list), _p(grad), _stream(dev)))
    return loss, loss_sum, grad, rankings, per_list, n_valid


def fused_step_scored(mask, gt, pred, K, n, R, strategy, threshold=0.03, equality_penalty=-1000, promotion="nep50",
                      seed=0, offset=0, image_base=0, scale=None, want_rankings=True, want_grad=True, want_order=False,
                      want_per_list=False):
    """pld_fused_step_scored: draw n candidates/image, score, keep the best R, (optionally) loss + gradient.
    ``pred`` may be None (sampler only).  ranking_size 1..512 (thread-per-list kernels up to 16, group-per-list above).
    Returns dict(loss, loss_sum, grad, rankings, order, per_list, n_valid)."""
    mask = as_cuda(mask, torch.float32, "mask")
    gt3 = as_cuda(gt, torch.float32, "gt")
    if gt3.dim() == 4:
        gt3 = gt3[..., 0].contiguous()
    B, H, W = gt3.shape
    Hm, Wm = mask.shape[1], mask.shape[2]
    dev = gt3.device
    do_loss = pred is not None
    if do_loss:
        pred = as_cuda(pred, torch.float32, "pred")
    if scale is None:
        scale = 1.0 / float(B * R)
    ctx = _ctx(gt3)
    with torch.cuda.device(dev):
        loss = torch.empty(1, dtype=torch.float32, device=dev) if do_loss else None
        loss_sum = torch.empty(1, dtype=torch.float64, device=dev) if do_loss else None
        n_valid = torch.empty(B, dtype=torch.int32, device=dev)
        rankings = torch.empty((B, R, K, 2), dtype=torch.float32, device=dev) if (want_rankings or not do_loss) else None
        grad = torch.empty_like(pred) if (do_loss and want_grad) else None
        order = torch.empty((B, R), dtype=torch.int32, device=dev) if want_order else None
        per_list = torch.empty(B * R, dtype=torch.float32, device=dev) if (do_loss and want_per_list) else None
        check(ctx.lib.pld_fused_step_scored(ctx.handle, _p(mask), _p(gt3), _p(pred), B, Hm, Wm, H, W, int(K), int(n),
                                            int(R), _lib.STRATEGY[strategy], float(threshold), float(equality_penalty),
                                            _lib.PROMOTION[promotion], int(seed), int(offset), int(image_base),
                                            float(scale), _p(n_valid), _p(order), _p(rankings), _p(loss), _p(loss_sum),
                                            _p(per_list), _p(grad), _stream(dev)))
    return dict(loss=loss, loss_sum=loss_sum, grad=grad, rankings=rankings, order=order, per_list=per_list,
                n_valid=n_valid)


def check_status(device=None):
    """Synchronise the current stream of ``device`` and raise the Python exception the
    reference would have raised for bad data (empty mask, index out of range)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ctx = Context.current(dev.index or 0)
    return ctx.raise_on_status(torch.cuda.current_stream(dev).cuda_stream)
