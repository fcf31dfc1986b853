"""Keras-signature ListMLE / Plackett-Luce loss on the fused CUDA kernel.

Mirror of pldepth/losses/nll_loss.py:32-62 (``HourglassNegativeLogLikelihood`` ->
``FullyFledgedMetaBatchListMLELoss`` -> TF-Ranking 0.3.1 ``ListMLELoss``) with the reshape /
gather contract of pldepth/data/depth_utils.py:39-61:

    loss = HourglassNegativeLogLikelihood(ranking_size=K, batch_size=B)
    value = loss(y_true, y_pred)      # y_true (B,R,K,2) f32, y_pred (B,H,W,1)|(B,H,W) f32
    value.backward()                  # d value / d y_pred, dense, duplicates accumulated

R is inferred from y_true (depth_utils.py:43) so train / validation R may differ
(nll_loss.py:58).  Forward and backward are ONE kernel launch: the gradient is produced
together with the loss and handed to autograd.
"""
import torch

from . import ops

_REDUCTIONS = ("auto", "sum_over_batch_size", "sum", "none")


class _ListMLEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, batch_size, ranking_size, reduction, global_lists):
        R = y_true.numel() // (batch_size * ranking_size * 2)
        L = batch_size * R
        if reduction in ("auto", "sum_over_batch_size"):
            scale = 1.0 / float(global_lists if global_lists else L)
        else:
            scale = 1.0
        need_grad = y_pred.requires_grad
        want_pl = reduction == "none"
        if want_pl and need_grad:
            raise NotImplementedError("reduction='none' is forward-only (per-list NLL); reduce before backward")
        loss, loss_sum, grad, per_list = ops.listmle_fwd_bwd(y_true, y_pred.detach(), batch_size, ranking_size,
                                                             scale, want_grad=need_grad, want_per_list=want_pl)
        ctx.has_grad = need_grad
        if need_grad:
            ctx.save_for_backward(grad)
        ctx.pred_shape = y_pred.shape
        ctx.pred_dtype = y_pred.dtype
        if want_pl:
            return per_list.reshape(L, 1)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.has_grad:
            return None, None, None, None, None, None
        (g,) = ctx.saved_tensors
        g = g.reshape(ctx.pred_shape) * grad_output
        return g.to(ctx.pred_dtype), None, None, None, None, None


class HourglassNegativeLogLikelihood(object):
    """Drop-in for pldepth/losses/nll_loss.py:32-40 (constructor arguments identical).

    ``reduction``: 'auto' / 'sum_over_batch_size' (Keras AUTO, mean over the B*R lists --
    what every reference script uses), 'sum', or 'none' (per-list NLL ``(L, 1)``, forward only).
    ``lambda_weight`` must be None (every call site in the reference passes None).
    ``global_lists`` (extension): total number of lists across all data-parallel ranks, so a
    shard's loss / gradient carry the global 1/L factor (see pldepth_b200.dist).
    """

    def __init__(self, ranking_size, batch_size, reduction="auto", name=None, lambda_weight=None,
                 debug=False, global_lists=None):
        red = getattr(reduction, "name", reduction)   # accept tf.losses.Reduction members
        red = str(red).lower()
        if red not in _REDUCTIONS:
            raise ValueError("unsupported reduction %r" % (reduction,))
        if lambda_weight is not None:
            raise NotImplementedError("lambda_weight is not used anywhere in the reference; only None is supported")
        if int(ranking_size) < 1 or int(ranking_size) > 512:
            raise ValueError("ranking_size must be in [1, 512]")
        self.ranking_size = int(ranking_size)
        self.batch_size = int(batch_size)
        self.reduction = red
        self.name = name
        self.debug = debug
        self.global_lists = global_lists

    def __call__(self, y_true, y_pred, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is never passed by the reference")
        if not isinstance(y_pred, torch.Tensor):
            y_pred = torch.from_dlpack(y_pred)
        if not isinstance(y_true, torch.Tensor):
            y_true = torch.from_dlpack(y_true)
        if self.debug:
            print("point_coords:", y_true.reshape(self.batch_size, -1, self.ranking_size, 2)[..., 0].flatten()[:10])
        return _ListMLEFunction.apply(y_pred, y_true, self.batch_size, self.ranking_size, self.reduction,
                                      self.global_lists)

    def get_config(self):
        return {"ranking_size": self.ranking_size, "batch_size": self.batch_size, "reduction": self.reduction,
                "name": self.name}


class NegativeLogLikelihoodLoss(object):
    """Drop-in for pldepth/losses/nll_loss.py:10-29 (``NegativeLogLikelihoodLoss`` -> ``MetaBatchListMLELoss``):
    labels and logits arrive already gathered, any shape that reshapes to ``[-1, ranking_size]``.  Runs on the
    same kernel by viewing the logits as one flat map indexed 0..L*K-1 (chunked below 2^23 entries)."""

    def __init__(self, ranking_size, reduction="auto", name=None, lambda_weight=None):
        red = str(getattr(reduction, "name", reduction)).lower()
        if red not in ("auto", "sum_over_batch_size", "sum"):
            raise ValueError("unsupported reduction %r" % (reduction,))
        if lambda_weight is not None:
            raise NotImplementedError("lambda_weight is not used anywhere in the reference; only None is supported")
        self.ranking_size, self.reduction, self.name = int(ranking_size), red, name

    def __call__(self, y_true, y_pred, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is never passed by the reference")
        K = self.ranking_size
        labels = y_true.reshape(-1, K).to(torch.float32)
        logits = y_pred.reshape(-1, K)
        L = labels.shape[0]
        scale = 1.0 / L if self.reduction != "sum" else 1.0
        chunk = max(1, (1 << 23) // K)
        total = None
        for lo in range(0, L, chunk):
            hi = min(L, lo + chunk)
            n = hi - lo
            idx = torch.arange(n * K, device=labels.device, dtype=torch.float32).reshape(n, K)
            y = torch.stack([idx, labels[lo:hi]], dim=-1).reshape(1, n, K, 2)
            part = _ListMLEFunction.apply(logits[lo:hi].reshape(1, n * K), y, 1, K, "sum", None) * scale
            total = part if total is None else total + part
        return total


MetaBatchListMLELoss = NegativeLogLikelihoodLoss


class _FusedStepFunction(torch.autograd.Function):
    """``FusedPLStep`` as a differentiable op: forward = sampling + gather + loss + dense gradient in the step's
    kernels; backward hands that gradient (times the incoming scalar) to autograd -- no second pass."""

    @staticmethod
    def forward(ctx, y_pred, gt, mask, step, holder):
        B = gt.shape[0]
        dev = y_pred.device
        pred = y_pred.detach()
        out = dict(grad=torch.empty(pred.numel(), dtype=torch.float32, device=dev),   # fresh: survives later steps
                   loss=torch.empty(1, dtype=torch.float32, device=dev),
                   loss_sum=torch.empty(1, dtype=torch.float64, device=dev),
                   n_valid=torch.empty(B, dtype=torch.int32, device=dev),
                   rankings=(torch.empty((B, step.R, step.K, 2), dtype=torch.float32, device=dev)
                             if step.emit_rankings else None))
        step.run(gt, mask, pred, out=out)
        holder.update(out)
        ctx.save_for_backward(out["grad"])
        ctx.pred_shape, ctx.pred_dtype = y_pred.shape, y_pred.dtype
        return out["loss"].reshape(())

    @staticmethod
    def backward(ctx, grad_output):
        (g,) = ctx.saved_tensors
        return (g.reshape(ctx.pred_shape) * grad_output).to(ctx.pred_dtype), None, None, None, None


class SampledHourglassNLL(object):
    """The reference's training-time pair -- ranking sampler in the tf.data map (hourglass_provider.py:55-62,75-86) +
    ``HourglassNegativeLogLikelihood`` as the compiled Keras loss (PLDepth.py:129-134) -- as ONE differentiable call
    on device-resident tensors:

        criterion = SampledHourglassNLL(ranking_size=5, rankings_per_image=100, strategy="information")
        loss = criterion(gt, mask, y_pred)      # gt [B,H,W], mask [B,Hm,Wm], y_pred [B,H,W,1] | [B,1,H,W] | [B,H,W]
        loss.backward()                         # the dense PL gradient flows into the decoder

    Every call draws fresh lists (Philox stream ``seed``, one offset per call).  Data-parallel shards pass
    ``global_batch`` (all images of the step) and ``image_base`` (index of this rank's first image) so the loss and
    gradient carry the global 1 / (B_global * R) factor and the shards draw disjoint streams; the returned loss is
    then this shard's share of the mean (sum the shares for reporting).  ``last`` holds the buffers of the latest call
    (``rankings`` when ``emit_rankings=True``, ``loss_sum``, ``n_valid``)."""

    def __init__(self, ranking_size, rankings_per_image, strategy="purely", seed=0, emit_rankings=False,
                 global_batch=None, image_base=0, candidate_factor=None, threshold=0.03, equality_penalty=-1000,
                 promotion="nep50", context=None):
        from .step import FusedPLStep
        self.step = FusedPLStep(ranking_size, rankings_per_image, seed=seed, emit_rankings=emit_rankings,
                                global_batch=global_batch, image_base=image_base, strategy=strategy,
                                candidate_factor=candidate_factor, threshold=threshold,
                                equality_penalty=equality_penalty, promotion=promotion, context=context)
        self.last = {}

    def __call__(self, gt, mask, y_pred):
        if not isinstance(y_pred, torch.Tensor):
            y_pred = torch.from_dlpack(y_pred)
        gt = ops.as_cuda(gt, torch.float32, "gt")
        if gt.dim() == 4 and gt.shape[-1] == 1:
            gt = gt[..., 0]
        if gt.dim() == 4 and gt.shape[1] == 1:
            gt = gt[:, 0]
        if y_pred.dim() == 4 and y_pred.shape[1] != 1 and y_pred.shape[-1] != 1:
            raise ValueError("y_pred must have one channel, got %s" % (tuple(y_pred.shape),))
        if not y_pred.is_contiguous():
            y_pred = y_pred.contiguous()
        return _FusedStepFunction.apply(y_pred, gt.contiguous(), mask, self.step, self.last)
