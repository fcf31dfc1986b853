"""Per-image sharding of the hot path over the GPUs of one node (one process per GPU).

The path partitions naturally by image: lists never span images
(hourglass_provider.py:75-86 samples per image; the gather is ``batch_dims=1``,
depth_utils.py:50) and image b's gradient touches only slice ``[b]``.  So every rank samples,
scores and differentiates its own contiguous block of images with the GLOBAL factor
1/(B_global * R), and the only exchange is a SUM all-reduce of one float64 (the loss sum) --
NCCL over NVLink on GPUs, gloo in the CPU tests.  Gradient slices are disjoint: no reduction;
``gather_grad`` all-gathers them only if a replicated map is wanted.

The Philox streams are keyed by the global image index (``image_base``), so a sharded run draws
exactly the lists a single-GPU run over the whole batch would draw.
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world_size):
    """Contiguous block [lo, hi) of images owned by ``rank`` (remainder spread over low ranks)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class LossWindow(object):
    """Per-step loss sums of one rank, reduced once per ``n`` steps instead of once per step.

    The path's only exchange is the all-reduce of one float64 per step, and nothing downstream of a step depends on it
    (a rank's gradient already carries the global 1/L).  Issued per step it is a collective kernel sharing the SMs with
    the list kernel every 0.17 ms; here step i writes its local sum straight into ``slot(i)`` (the fused step takes the
    pointer, no extra launch) and ``reduce()`` all-reduces the whole window -- one collective per ``n`` steps, the
    reduced per-step losses arriving at most ``n - 1`` steps late.
    """

    def __init__(self, n, device, group=None):
        self.n = int(n)
        if self.n < 1:
            raise ValueError("window must hold at least one step")
        self.group = group
        self.buf = torch.zeros(self.n, dtype=torch.float64, device=device)
        self.filled = 0

    def slot(self, i):
        """1-element float64 view the step of index ``i`` (mod n) writes its local loss sum into."""
        k = int(i) % self.n
        return self.buf[k:k + 1]

    def mark(self):
        """Account for one written slot; True when the window is full and should be reduced."""
        self.filled += 1
        return self.filled >= self.n

    def reduce(self, async_op=False):
        """SUM all-reduce of the window (no-op without an initialised process group).  Returns the work handle (or None);
        after it completes ``values()`` holds the global per-step loss sums of the last ``filled`` steps."""
        count, self.filled = self.filled, 0
        self.last_count = count
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            return dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return None

    def values(self):
        return self.buf[: getattr(self, "last_count", self.n)]


class ShardedPLStep(object):
    """Data-parallel wrapper around a per-shard step.

    ``local_step(gt, mask, pred, image_base, global_batch) -> dict(loss_sum f64[1], grad, rankings)``
    defaults to the fused CUDA step; tests inject a CPU stand-in to exercise the sharding /
    reduction logic under gloo.
    """

    def __init__(self, ranking_size, rankings_per_image, global_batch, seed=0, group=None, local_step=None,
                 emit_rankings=True, async_loss=False, loss_every=1):
        """``loss_every`` = n > 1 takes the collective off the step: local loss sums are collected in a ``LossWindow``
        and all-reduced once per n steps; ``run`` then returns ``loss=None`` except on the step that closes a window,
        where ``losses`` holds the n global mean losses (oldest first).  ``flush()`` reduces a partial window."""
        self.K = int(ranking_size)
        self.R = int(rankings_per_image)
        self.global_batch = int(global_batch)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.lo, self.hi = shard_bounds(self.global_batch, self.rank, self.world)
        self._local_step = local_step
        self._fused = None
        self.seed = int(seed)
        self.emit_rankings = emit_rankings
        self.async_loss = bool(async_loss)
        self.loss_every = int(loss_every)
        self._window = None
        self._steps = 0

    def _default_step(self, gt, mask, pred, image_base, global_batch):
        from .step import FusedPLStep
        if self._fused is None:
            self._fused = FusedPLStep(self.K, self.R, seed=self.seed, emit_rankings=self.emit_rankings,
                                      global_batch=global_batch, image_base=image_base)
        return self._fused.run(gt, mask, pred)

    def run(self, gt_local, mask_local, pred_local):
        """Inputs are this rank's images ``[hi-lo, ...]``.  Returns dict with the GLOBAL mean loss
        (identical on every rank), the global loss sum and this rank's gradient slice."""
        n_local = self.hi - self.lo
        if gt_local.shape[0] != n_local:
            raise ValueError("rank %d owns %d images, got %d" % (self.rank, n_local, gt_local.shape[0]))
        step = self._local_step or self._default_step
        out = step(gt_local, mask_local, pred_local, self.lo, self.global_batch)
        if self.loss_every > 1:
            if self._window is None:
                self._window = LossWindow(self.loss_every, out["loss_sum"].device, self.group)
            self._window.slot(self._steps).copy_(out["loss_sum"].reshape(1))
            self._steps += 1
            res = dict(loss=None, loss_sum=None, losses=None, grad=out["grad"], rankings=out.get("rankings"))
            if self._window.mark():
                self._window.reduce()
                res["losses"] = (self._window.values() / float(self.global_batch * self.R)).to(torch.float32).clone()
                res["loss"] = res["losses"][-1:]
            return res
        total = out["loss_sum"].clone()
        work = None
        if self.world > 1:
            # nothing downstream of the step depends on the reduced loss (the local gradient already carries
            # the global 1/L), so with async_loss the reduction overlaps whatever the caller enqueues next
            work = dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group, async_op=self.async_loss)
        if work is not None and self.async_loss:
            return dict(loss=None, loss_sum=total, loss_work=work, grad=out["grad"], rankings=out.get("rankings"))
        loss = (total / float(self.global_batch * self.R)).to(torch.float32)
        return dict(loss=loss, loss_sum=total, grad=out["grad"], rankings=out.get("rankings"))

    def flush(self):
        """Reduce a partially filled window; returns the global mean losses of its steps (oldest first) or None."""
        if self._window is None or self._window.filled == 0:
            return None
        self._window.reduce()
        return (self._window.values() / float(self.global_batch * self.R)).to(torch.float32).clone()

    def finish_loss(self, result):
        """Wait for an asynchronous loss reduction and return the global mean loss."""
        if result.get("loss_work") is not None:
            result["loss_work"].wait()
        return (result["loss_sum"] / float(self.global_batch * self.R)).to(torch.float32)

    def gather_grad(self, grad_local):
        """Replicated dense gradient [B_global, ...] from the disjoint slices (all-gather)."""
        if self.world == 1:
            return grad_local
        sizes = [hi - lo for lo, hi in (shard_bounds(self.global_batch, r, self.world) for r in range(self.world))]
        cap = max(sizes)          # all_gather needs equal shapes: pad ragged shards, trim afterwards
        mine = grad_local.contiguous()
        if mine.shape[0] < cap:
            pad = torch.zeros((cap - mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
            mine = torch.cat([mine, pad], dim=0)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
