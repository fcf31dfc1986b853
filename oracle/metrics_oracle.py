"""NumPy restatement of the reference's evaluation metrics and list producers next to the hot path
-- TEST INFRASTRUCTURE ONLY.  Follows pldepth/active_learning/metrics.py:60-110 and
pldepth/active_learning/active_learning_method.py:59-76.  ``calc_d`` uses OpenCV's
``cv2.normalize(..., NORM_MINMAX)`` in the reference; it is restated as (x - min) / (max - min) and
checked against the real ``cv2`` where it is importable (tests/test_oracle_metrics.py)."""
import numpy as np


def ordinal_error(op, gt, imsize=(448, 448), num=5000):
    rs = np.random.RandomState(10)
    idx = rs.choice(list(range(imsize[0] * imsize[1])), num * 2, replace=False)
    idx0, idx1 = np.split(idx, 2)
    op_flat, gt_flat = np.asarray(op).flatten(), np.asarray(gt).flatten()
    out_order = np.greater(op_flat[idx0], op_flat[idx1])
    gt_order = np.greater(gt_flat[idx0], gt_flat[idx1])
    return 1 - np.equal(out_order, gt_order).sum() / num


def calc_dcg(rel):
    return (rel / np.log2(np.arange(np.shape(rel)[0]) + 2)).sum()


def calc_d(op, gt, imsize=(224, 224), list_size=200, normalize=None):
    op = np.asarray(op, dtype=np.float32)
    if normalize is None:
        mn, mx = np.float64(op.min()), np.float64(op.max())
        op = ((op.astype(np.float64) - mn) * (1.0 / (mx - mn))).astype(np.float32)
    else:
        op = normalize(op)
    op_flat, gt_flat = op.flatten(), np.asarray(gt).flatten()
    rs = np.random.RandomState(69)
    ids = rs.choice(np.arange(imsize[0] * imsize[1]), size=list_size, replace=False)
    rel_d = 1 / (np.sort(op_flat[ids]).astype(np.float64) + 1)
    rel_g = 1 / (np.sort(gt_flat[ids]).astype(np.float64) + 1)
    return calc_dcg(rel_d) / calc_dcg(rel_g)


def oracle_lists(gt, pos_xy, ranking_size, img_size=(224, 224, 3)):
    """active_learning_method.py:59-76 without the shuffle (the caller shuffles)."""
    K = ranking_size
    out = np.zeros([int(pos_xy.shape[0] / K), K, 2], dtype=np.float32)
    buf = np.zeros((K, 2))
    j = 0
    for i in range(0, pos_xy.shape[0] - K, K):
        for k in range(K):
            buf[k, 0] = pos_xy[i + k, 0] * img_size[0] + pos_xy[i + k, 1]
            buf[k, 1] = gt[tuple(pos_xy[i + k])]
        ix = np.argsort(buf[:, 1], kind="stable")[::-1]
        out[j] = buf[ix]
        j += 1
    return out
