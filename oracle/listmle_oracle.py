"""NumPy restatement of the gather + ListMLE / Plackett-Luce NLL (+ gradient) -- TEST
INFRASTRUCTURE ONLY.  **Parity unpinned** against the real TF-Ranking binary (see
oracle/__init__.py); pinned by known-answer vectors, closed-form properties and an
independent fp64 autograd in tests/test_oracle_listmle.py.

Follows:
* pldepth/data/depth_utils.py:39-61 ``prepare_fully_fledged_loss_input``: labels
  ``(B,R,K,2)`` -> flat indices ``int32(labels[...,0])``, depths ``labels[...,1]``;
  ``selected = gather(reshape(logits,[B,-1]), idx, axis=1, batch_dims=1)``; both -> ``[B*R, K]``.
* pldepth/losses/nll_loss.py:43-62 ``FullyFledgedMetaBatchListMLELoss.compute_unreduced_loss``.
* third-party ``tensorflow_ranking==0.3.1`` (requirements.txt:20),
  ``losses_impl.ListMLELoss.compute_unreduced_loss`` (published source, restated):
    is_valid = labels >= 0; labels' = where(valid, labels, 0);
    logits' = where(valid, logits, log(1e-10));
    key = where(valid, labels', min_row(labels') - 1e-6);
    sort rows by key descending (ties shuffled at random by TF; here: stable, earlier
    position first); s -= max_row(s); sums = log(cumsum(exp(s), reverse)) - s;
    nll = sum_row(sums); weights = 1.
* ``keras.losses._RankingLoss.__call__`` + Keras ``SUM_OVER_BATCH_SIZE`` (reduction AUTO,
  nll_loss.py:33): scalar = mean of the ``(L,1)`` tensor.

The gradient is the closed form of that expression:
``d nll / d s_k = e_k * sum_{i<=k} 1/S_i - 1`` in sorted order, zero for invalid entries,
scaled by ``1/L`` and scatter-added (duplicates accumulate) into a dense ``(B, H*W)`` map.
"""
import numpy as np

LOG_EPS = np.float32(np.log(np.float32(1e-10)))   # tf.math.log(_EPSILON) in float32


def split_rankings(y_true, batch_size, ranking_size):
    """depth_utils.py:43-46,57: -> (idx int32 [B, R*K], labels f32 [B*R, K])."""
    y = np.asarray(y_true, dtype=np.float32).reshape(batch_size, -1, ranking_size, 2)
    idx = y[..., 0].reshape(batch_size, -1).astype(np.int32)
    labels = y[..., 1].reshape(-1, ranking_size)
    return idx, labels


def gather_predictions(y_pred, idx, batch_size, ranking_size):
    """depth_utils.py:45,50,52: batched gather of predictions at the sampled flat indices."""
    pred = np.asarray(y_pred, dtype=np.float32).reshape(batch_size, -1)
    sel = np.take_along_axis(pred, idx.astype(np.int64), axis=1)
    return sel.reshape(-1, ranking_size)


def sort_order(labels):
    """Row-wise order used by ListMLE (valid labels descending, invalid last; stable)."""
    labels = np.asarray(labels, dtype=np.float32)
    valid = labels >= 0
    lab0 = np.where(valid, labels, np.float32(0))
    key = np.where(valid, lab0, lab0.min(axis=1, keepdims=True) - np.float32(1e-6)).astype(np.float32)
    return np.argsort(-key.astype(np.float64), axis=1, kind="stable"), valid


def listmle_per_list(labels, scores, dtype=np.float64):
    """Per-list NLL and d nll / d scores (unsorted positions).  ``dtype`` = arithmetic type
    (float64 = ground truth, float32 = what TF computes)."""
    labels = np.asarray(labels, dtype=np.float32)
    order, valid = sort_order(labels)
    s_in = np.where(valid, np.asarray(scores, dtype=np.float32), LOG_EPS).astype(dtype)
    s = np.take_along_axis(s_in, order, axis=1)
    v_sorted = np.take_along_axis(valid, order, axis=1)
    s = s - s.max(axis=1, keepdims=True)
    e = np.exp(s)
    S = np.cumsum(e[:, ::-1], axis=1, dtype=dtype)[:, ::-1]            # reverse cumsum
    nll = (np.log(S) - s).sum(axis=1, dtype=dtype)
    g_sorted = e * np.cumsum(1.0 / S, axis=1, dtype=dtype) - 1.0
    g_sorted = np.where(v_sorted, g_sorted, 0.0)
    grad = np.zeros_like(g_sorted)
    np.put_along_axis(grad, order, g_sorted, axis=1)
    return nll, grad


def hourglass_nll(y_true, y_pred, batch_size, ranking_size, reduction="auto", dtype=np.float64,
                  global_lists=None):
    """Full loss as Keras calls it: returns (loss, grad wrt y_pred with y_pred's shape,
    per_list_nll[L]).  ``reduction``: auto|sum_over_batch_size (mean over L), sum, none.
    ``global_lists`` overrides L in the mean (multi-GPU shards pass B_global*R)."""
    y_pred = np.asarray(y_pred, dtype=np.float32)
    idx, labels = split_rankings(y_true, batch_size, ranking_size)
    scores = gather_predictions(y_pred, idx, batch_size, ranking_size)
    nll, g = listmle_per_list(labels, scores, dtype)
    L = nll.shape[0]
    if reduction in ("auto", "sum_over_batch_size"):
        scale = 1.0 / float(L if global_lists is None else global_lists)
        loss = nll.sum(dtype=dtype) * scale
    elif reduction == "sum":
        scale = 1.0
        loss = nll.sum(dtype=dtype)
    elif reduction == "none":
        scale = 1.0
        loss = nll.reshape(-1, 1)
    else:
        raise ValueError(reduction)
    dense = np.zeros((batch_size, y_pred.size // batch_size), dtype=dtype)
    rows = np.repeat(np.arange(batch_size), idx.shape[1])
    np.add.at(dense, (rows, idx.reshape(-1).astype(np.int64)), (g * scale).reshape(-1))
    return loss, dense.reshape(y_pred.shape), nll
