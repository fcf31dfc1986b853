"""The fused hot-path step as one object: valid-pixel table -> Philox sampling -> gather ->
ListMLE forward + backward, for a batch of images resident on one GPU.

This is what replaces, per training step, the reference's tf.data sampler map
(hourglass_provider.py:55-62), ``prepare_fully_fledged_loss_input`` (depth_utils.py:39-61) and the
TF-Ranking ListMLE loss + its autodiff (nll_loss.py:32-62).  Output buffers are preallocated and
reused, so a step is three kernel launches on the current stream (core sampler; the score-based strategies add
their scoring pass and top-R selection), no host sync.
"""
import ctypes
import weakref

import torch

from . import ops
from ._lib import Context, check, c_void_p


class FusedPLStep(object):
    def __init__(self, ranking_size, rankings_per_image, seed=0, emit_rankings=True, global_batch=None,
                 image_base=0, strategy="purely", candidate_factor=None, threshold=0.03, equality_penalty=-1000,
                 promotion="nep50", context=None, first_step=0):
        """``strategy``: 'purely' (core sampler, exactly R lists/image -- the headline path) or one of
        'masked' / 'thresholded' / 'information' (R best of int(R * candidate_factor) candidates, default
        factors 1.5 / 1.5 / 5 as in sampling.py:157,190,218).
        ``context``: a private ``_lib.Context`` (own scratch and lookup tables) instead of the thread's; steps of
        independent batches that run concurrently on different streams need one each.  ``first_step`` is the Philox
        offset of the first call (concurrent lanes take disjoint ranges)."""
        self.K = int(ranking_size)
        self.R = int(rankings_per_image)
        self.seed = int(seed)
        self.emit_rankings = bool(emit_rankings)
        self.global_batch = global_batch
        self.image_base = int(image_base)
        self.step_index = int(first_step)
        self._ctx = context
        self._buf = None
        self._validated = {}
        if strategy not in ("purely", "masked", "thresholded", "information"):
            raise ValueError("unknown strategy %r" % (strategy,))
        self.strategy = strategy
        default_f = {"purely": 1.0, "masked": 1.5, "thresholded": 1.5, "information": 5}[strategy]
        self.n_candidates = int(self.R * (default_f if candidate_factor is None else candidate_factor))
        self.threshold, self.equality_penalty, self.promotion = threshold, equality_penalty, promotion
        if strategy != "purely" and self.n_candidates < self.R:
            raise ValueError("scored strategies need candidate_factor >= 1")

    def _buffers(self, B, H, W, Hm, Wm, dev):
        key = (B, H, W, Hm, Wm, dev)
        if self._buf is None or self._buf["key"] != key:
            self._buf = dict(
                key=key,
                valid_flat=torch.empty((B, Hm * Wm), dtype=torch.int32, device=dev),
                n_valid=torch.empty((B,), dtype=torch.int32, device=dev),
                rankings=(torch.empty((B, self.R, self.K, 2), dtype=torch.float32, device=dev)
                          if self.emit_rankings else None),
                grad=torch.empty((B, H, W, 1), dtype=torch.float32, device=dev),
                loss=torch.empty(1, dtype=torch.float32, device=dev),
                loss_sum=torch.empty(1, dtype=torch.float64, device=dev),
            )
        return self._buf

    def run(self, gt, mask, pred, out=None):
        """gt f32[B,H,W], mask f32[B,Hm,Wm], pred f32[B,H,W(,1)] on one CUDA device.
        Returns dict(loss f32[1], loss_sum f64[1], grad f32[B,H,W,1], rankings f32[B,R,K,2]|None).
        ``loss`` carries the factor 1/(global_batch*R) (global_batch defaults to B)."""
        gt = ops.as_cuda(gt, torch.float32, "gt")
        if torch.cuda.current_device() != (gt.device.index or 0):
            with torch.cuda.device(gt.device):
                return self._run(gt, mask, pred, out)
        return self._run(gt, mask, pred, out)

    def _validate(self, gt, mask, pred, out):
        """The C ABI takes raw pointers: everything that reaches it is float32, dense, on ONE CUDA device and of the
        shape the call declares.  Dtype / layout mismatches of the inputs are converted (a copy); anything
        that cannot be fixed by a conversion raises."""
        gt = ops.as_cuda(gt, torch.float32, "gt")
        # masks stay uint8 when they arrive as uint8 / bool (nonzero = valid; pld_fused_step_m8), else float32
        m_dtype = getattr(mask, "dtype", None)
        if m_dtype == torch.bool:
            mask = mask.to(torch.uint8)
        mask = ops.as_cuda(mask, torch.uint8 if m_dtype in (torch.uint8, torch.bool) else torch.float32, "mask")
        if mask.dtype == torch.uint8 and self.strategy != "purely":
            mask = mask.to(torch.float32)          # the scored step takes float masks only
        pred = ops.as_cuda(pred, torch.float32, "pred")
        if gt.dim() == 4 and gt.shape[-1] == 1:
            gt = gt[..., 0].contiguous()
        if mask.dim() == 4 and mask.shape[-1] == 1:
            mask = mask[..., 0].contiguous()
        if gt.dim() != 3:
            raise ValueError("gt must be [B,H,W] (or [B,H,W,1]), got %s" % (tuple(gt.shape),))
        if mask.dim() != 3 or mask.shape[0] != gt.shape[0]:
            raise ValueError("mask must be [B,Hm,Wm] with the batch size of gt, got %s" % (tuple(mask.shape),))
        B, H, W = gt.shape
        if pred.numel() != B * H * W or pred.shape[0] != B:
            raise ValueError("pred must hold B*H*W = %d values ([B,H,W] or [B,H,W,1]), got %s" %
                             (B * H * W, tuple(pred.shape)))
        if mask.device != gt.device or pred.device != gt.device:
            raise ValueError("gt, mask and pred must live on the same device")
        if out is not None:
            want = {"grad": ((B * H * W,), torch.float32), "loss": ((1,), torch.float32),
                    "loss_sum": ((1,), torch.float64), "n_valid": ((B,), torch.int32)}
            if out.get("rankings") is not None:
                want["rankings"] = ((B * self.R * self.K * 2,), torch.float32)
            for name, (shape, dtype) in want.items():
                t = out.get(name)
                if not isinstance(t, torch.Tensor):
                    raise ValueError("out[%r] is missing" % name)
                if t.numel() != shape[0] or t.dtype != dtype or not t.is_contiguous() or t.device != gt.device:
                    raise ValueError("out[%r] must be a contiguous %s tensor of %d elements on %s" %
                                     (name, dtype, shape[0], gt.device))
            if self.emit_rankings and out.get("rankings") is None:
                raise ValueError("out['rankings'] is required when emit_rankings=True")
        return gt, mask, pred

    def _run(self, gt, mask, pred, out=None):
        # validation is cached per set of tensor objects (a training loop passes the same buffers every step; the
        # checks cost more host time than the launches of a step).  Weak references: nothing is kept alive.
        key = (id(gt), id(mask), id(pred), id(out))
        seen = self._validated.get(key)
        if seen is not None and seen[0]() is gt and seen[1]() is mask and seen[2]() is pred and \
                seen[3] == (gt.data_ptr(), mask.data_ptr(), pred.data_ptr()):
            dims, mask_u8 = seen[4], seen[5]
            if out is not None:     # callers may swap the scalar slots between steps (loss windows)
                ls = out.get("loss_sum")
                if not isinstance(ls, torch.Tensor) or ls.dtype != torch.float64 or ls.numel() != 1 or \
                        ls.device != gt.device:
                    raise ValueError("out['loss_sum'] must be a float64 tensor of one element on %s" % gt.device)
        else:
            g0, m0, p0 = gt, mask, pred
            gt, mask, pred = self._validate(gt, mask, pred, out)
            dims = (gt.shape[0], gt.shape[1], gt.shape[2], mask.shape[1], mask.shape[2])
            mask_u8 = mask.dtype == torch.uint8
            if all(isinstance(t, torch.Tensor) for t in (g0, m0, p0)) and gt.data_ptr() == g0.data_ptr() \
                    and pred.data_ptr() == p0.data_ptr() and mask.data_ptr() == m0.data_ptr():
                if len(self._validated) > 64:
                    self._validated.clear()
                self._validated[key] = (weakref.ref(g0), weakref.ref(m0), weakref.ref(p0),
                                        (g0.data_ptr(), m0.data_ptr(), p0.data_ptr()), dims, mask_u8)
        dev = gt.device
        B, H, W, Hm, Wm = dims
        buf = out if out is not None else self._buffers(B, H, W, Hm, Wm, dev)
        ctx = self._ctx if self._ctx is not None else Context.current(dev.index or 0)
        lib = ctx.lib
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        gb = self.global_batch if self.global_batch else B
        scale = 1.0 / (float(gb) * float(self.R))
        p = lambda t: c_void_p(t.data_ptr()) if t is not None else c_void_p(None)
        if self.strategy != "purely":
            from ._lib import STRATEGY, PROMOTION
            check(lib.pld_fused_step_scored(ctx.handle, p(mask), p(gt), p(pred), B, Hm, Wm, H, W, self.K,
                                            self.n_candidates, self.R, STRATEGY[self.strategy], float(self.threshold),
                                            float(self.equality_penalty), PROMOTION[self.promotion], self.seed,
                                            self.step_index, self.image_base, ctypes.c_float(scale), p(buf["n_valid"]),
                                            c_void_p(None), p(buf["rankings"]), p(buf["loss"]), p(buf["loss_sum"]),
                                            c_void_p(None), p(buf["grad"]), stream))
            self.step_index += 1
            return buf
        # one call: mask analysis + zeroed grad, 8-byte lookup tables, fused list kernel
        entry = lib.pld_fused_step_m8 if mask_u8 else lib.pld_fused_step
        check(entry(ctx.handle, p(mask), p(gt), p(pred), B, Hm, Wm, H, W, self.K, self.R, self.seed,
                    self.step_index, self.image_base, ctypes.c_float(scale), p(buf["n_valid"]),
                    p(buf["rankings"]), p(buf["loss"]), p(buf["loss_sum"]), c_void_p(None),
                    p(buf["grad"]), stream))
        self.step_index += 1
        return buf

    def capture(self, gt, mask, pred, out=None):
        """Capture the step for these (static) tensors in a CUDA graph.  The Philox offset moves into device
        memory (``pld_ctx_device_offset``), so every ``graph.replay()`` draws fresh lists; new data is fed by
        copying into the same gt / mask / pred tensors.  Returns (graph, output buffers).  While a captured step
        is in use, other Philox calls on this thread's context share (and advance) the same device counter."""
        for name, t in (("gt", gt), ("mask", mask), ("pred", pred)):
            ok_dtype = (torch.float32, torch.uint8) if name == "mask" else (torch.float32,)
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in ok_dtype and t.is_contiguous()):
                raise ValueError("capture(): %s must already be a contiguous float32 (mask: or uint8) CUDA tensor (a "
                                 "converted copy would be frozen into the graph)" % name)
        dev = gt.device
        ctx = self._ctx if self._ctx is not None else Context.current(dev.index or 0)
        ctx.device_offset(True, self.step_index)
        self.run(gt, mask, pred, out=out)                 # warm-up: scratch allocation happens outside capture
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            buf = self.run(gt, mask, pred, out=out)
        return graph, buf

    @staticmethod
    def new_buffers(B, H, W, Hm, Wm, R, K, dev, emit_rankings=True):
        return dict(
            key=None,
            valid_flat=torch.empty((B, Hm * Wm), dtype=torch.int32, device=dev),
            n_valid=torch.empty((B,), dtype=torch.int32, device=dev),
            rankings=torch.empty((B, R, K, 2), dtype=torch.float32, device=dev) if emit_rankings else None,
            grad=torch.empty((B, H, W, 1), dtype=torch.float32, device=dev),
            loss=torch.empty(1, dtype=torch.float32, device=dev),
            loss_sum=torch.empty(1, dtype=torch.float64, device=dev),
        )

    def check(self, dev):
        """Synchronise the current stream and raise on the status bits of this step's context."""
        if self._ctx is not None:
            return self._ctx.raise_on_status(torch.cuda.current_stream(dev).cuda_stream)
        return ops.check_status(dev)


class HostPipelinedStep(object):
    """The same step fed from HOST buffers (NumPy / pinned tensors), as the reference's tf.data
    pipeline feeds Keras: inputs are copied host->device, the fused step runs, loss and dense
    gradient are copied back.  Two device/host slots and three streams (H2D, compute, D2H) let the
    copies of step i+1 and i-1 overlap the kernels of step i; PCIe is full duplex.

        runner = HostPipelinedStep(K, R, B, H, W)
        for batch in data:                       # gt, mask, pred: float32 host arrays
            ticket = runner.submit(gt, mask, pred)
        loss, grad = runner.result(ticket)       # waits for that step only
    """

    def __init__(self, ranking_size, rankings_per_image, B, H, W, Hm=None, Wm=None, seed=0, device=None,
                 emit_rankings=True, global_batch=None, image_base=0, slots=2, mask_dtype=torch.float32):
        """``mask_dtype``: torch.float32 (what the reference's pipeline hands over) or torch.uint8 (the masks as they
        are stored, nonzero = valid: a quarter of the mask bytes cross PCIe; the kernels read them directly)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        Hm, Wm = Hm or H, Wm or W
        self.dev = dev
        self.shape = (B, H, W, Hm, Wm)
        self.step = FusedPLStep(ranking_size, rankings_per_image, seed, emit_rankings, global_batch, image_base)
        self.slots = []
        for _ in range(slots):
            self.slots.append(dict(
                gt=torch.empty((B, H, W), dtype=torch.float32, device=dev),
                mask=torch.empty((B, Hm, Wm), dtype=mask_dtype, device=dev),
                pred=torch.empty((B, H, W, 1), dtype=torch.float32, device=dev),
                out=FusedPLStep.new_buffers(B, H, W, Hm, Wm, rankings_per_image, ranking_size, dev, emit_rankings),
                h_loss=torch.empty(1, dtype=torch.float32).pin_memory(),
                h_grad=torch.empty((B, H, W, 1), dtype=torch.float32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(), used=False))
        self.mask_dtype = mask_dtype
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.count = 0

    @staticmethod
    def _host(x, dtype=torch.float32):
        if isinstance(x, torch.Tensor):
            return x if x.dtype == dtype else x.to(dtype)
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint8 if dtype == torch.uint8 else np.float32))

    def submit(self, gt, mask, pred):
        i = self.count
        sl = self.slots[i % len(self.slots)]
        compute = torch.cuda.current_stream(self.dev)
        gt, mask, pred = self._host(gt), self._host(mask, self.mask_dtype), self._host(pred)
        with torch.cuda.stream(self.s_in):
            if sl["used"]:
                self.s_in.wait_event(sl["ev_done"])      # the slot's previous kernels have consumed the inputs
            sl["gt"].copy_(gt.reshape(sl["gt"].shape), non_blocking=True)
            sl["mask"].copy_(mask.reshape(sl["mask"].shape), non_blocking=True)
            sl["pred"].copy_(pred.reshape(sl["pred"].shape), non_blocking=True)
            sl["ev_in"].record(self.s_in)
        compute.wait_event(sl["ev_in"])
        if sl["used"]:
            compute.wait_event(sl["ev_out"])             # previous results of this slot have left the device
        out = self.step.run(sl["gt"], sl["mask"], sl["pred"], out=sl["out"])
        sl["ev_done"].record(compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl["ev_done"])
            sl["h_loss"].copy_(out["loss"], non_blocking=True)
            sl["h_grad"].copy_(out["grad"], non_blocking=True)
            sl["ev_out"].record(self.s_out)
        sl["used"] = True
        self.count += 1
        return i

    def result(self, ticket):
        """(loss float, grad pinned tensor [B,H,W,1]) of step ``ticket`` (must be one of the last
        ``slots`` submitted steps); blocks until its copies have landed."""
        if ticket < self.count - len(self.slots) or ticket >= self.count:
            raise ValueError("result of step %d is no longer (or not yet) available" % ticket)
        sl = self.slots[ticket % len(self.slots)]
        sl["ev_out"].synchronize()
        return float(sl["h_loss"][0]), sl["h_grad"]

    def bytes_per_step(self):
        B, H, W, Hm, Wm = self.shape
        mask_bytes = B * Hm * Wm * (1 if self.mask_dtype == torch.uint8 else 4)
        return 2 * B * H * W * 4 + mask_bytes, (B * H * W + 1) * 4
