"""TensorFlow / Keras adapter (import-guarded; TensorFlow is not installable in the build image, so this
module is exercised only where TF exists -- see INTEGRATION.md §2).

``HourglassNegativeLogLikelihood`` here is a ``tf.keras.losses.Loss`` with the reference's constructor
(pldepth/losses/nll_loss.py:32-40) whose ``call`` runs the fused CUDA kernel through ``tf.py_function`` +
``tf.custom_gradient``; tensors cross over DLPack without copies.  The sampler classes of
``pldepth_b200.sampling`` already speak NumPy and drop into ``tf.numpy_function`` unchanged
(hourglass_provider.py:55-58).
"""
try:
    import tensorflow as tf
except Exception as exc:  # pragma: no cover - TF absent in this image
    tf = None
    _IMPORT_ERROR = exc

import torch

from . import ops


def _require_tf():
    if tf is None:
        raise ImportError("pldepth_b200.tf_adapter needs TensorFlow >= 2.2 (%s)" % (_IMPORT_ERROR,))


def _to_torch(t):
    return torch.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t):
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))


def pl_nll(y_true, y_pred, batch_size, ranking_size, global_lists=None, reduction="auto"):
    """Scalar ListMLE loss with a custom gradient w.r.t. ``y_pred``.  ``reduction``: 'auto' /
    'sum_over_batch_size' (mean over the B*R lists, what every reference script uses) or 'sum'."""
    _require_tf()
    if reduction not in ("auto", "sum_over_batch_size", "sum"):
        raise ValueError("unsupported reduction %r (per-list output: use pldepth_b200.losses with reduction='none')"
                         % (reduction,))

    @tf.custom_gradient
    def _op(yt, yp):
        def run(a, b):
            # TF produced a / b on ITS stream and torch launches on its own: py_function hands over tensors whose
            # producers may still be running, so wait for the device before launching and before handing back
            a_t, b_t = _to_torch(a), _to_torch(b)
            torch.cuda.synchronize(b_t.device)
            R = a_t.numel() // (batch_size * ranking_size * 2)
            scale = 1.0 if reduction == "sum" else 1.0 / float(global_lists if global_lists else batch_size * R)
            loss, _, grad, _ = ops.listmle_fwd_bwd(a_t, b_t, batch_size, ranking_size, scale)
            torch.cuda.current_stream(b_t.device).synchronize()
            return _to_tf(loss.reshape(())), _to_tf(grad)
        loss, grad = tf.py_function(run, [yt, yp], [tf.float32, tf.float32])
        loss.set_shape(())

        def backward(upstream):
            return None, upstream * tf.reshape(grad, tf.shape(yp))
        return loss, backward
    return _op(tf.cast(y_true, tf.float32), tf.cast(y_pred, tf.float32))


if tf is not None:  # pragma: no cover
    class HourglassNegativeLogLikelihood(tf.keras.losses.Loss):
        """EXPERIMENTAL (never executed in the build image, which has no TensorFlow).  Keras loss with the
        reference's signature; the reduction is applied inside the kernel: AUTO / SUM_OVER_BATCH_SIZE = mean over
        the B*R lists (SUM_OVER_BATCH_SIZE over the reference's (L, 1) tensor), SUM = plain sum; NONE raises."""

        def __init__(self, ranking_size, batch_size, reduction=tf.keras.losses.Reduction.AUTO, name=None,
                     lambda_weight=None, debug=False):
            super().__init__(reduction=tf.keras.losses.Reduction.NONE, name=name)
            if lambda_weight is not None:
                raise NotImplementedError("lambda_weight is never used by the reference")
            red = str(getattr(reduction, "name", reduction)).lower()
            if red not in ("auto", "sum_over_batch_size", "sum"):
                raise NotImplementedError("reduction %r: only AUTO / SUM_OVER_BATCH_SIZE / SUM are supported" %
                                          (reduction,))
            self.ranking_size, self.batch_size, self.debug = int(ranking_size), int(batch_size), debug
            self._pld_reduction = red

        def call(self, y_true, y_pred):
            return pl_nll(y_true, y_pred, self.batch_size, self.ranking_size, reduction=self._pld_reduction)
