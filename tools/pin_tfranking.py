#!/usr/bin/env python
"""Pin the ListMLE oracle against the REAL reference loss, wherever TensorFlow + tensorflow_ranking==0.3.1 import.

The arithmetic of stage 3 lives in the un-vendored third-party package ``tensorflow_ranking==0.3.1``
(/root/reference/requirements.txt:20); this build image has neither TensorFlow nor network access, so
``oracle/listmle_oracle.py`` is "parity unpinned" (DESIGN.md section 2).  This script closes that gap on any machine
that has the reference's dependencies:

    pip install "tensorflow>=2.2,<2.5" tensorflow_ranking==0.3.1
    python tools/pin_tfranking.py --reference /path/to/PLDepth          # writes tests/golden/listmle_*.npz

It imports the reference's OWN ``HourglassNegativeLogLikelihood`` (pldepth/losses/nll_loss.py:32-62) from the given
checkout, runs it on the fixed inputs of ``make_cases()`` and stores, per case: the inputs, the Keras-reduced scalar
loss (``loss(y_true, y_pred)``, nll_loss.py:33 reduction AUTO), the unreduced per-list NLL
(``loss._loss.compute_unreduced_loss``, nll_loss.py:51-62) and the ``tf.GradientTape`` gradient w.r.t. ``y_pred``.
``tests/test_oracle_listmle.py`` and ``tests/test_gpu_listmle.py`` consume every ``tests/golden/listmle_*.npz`` they
find (oracle within 1e-6, CUDA path within 1e-5).  Inputs are tie-free: TF-Ranking breaks label ties at random.

``--dry-run`` lists the cases without TensorFlow (used by the CPU test-suite to keep the generator importable).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def make_cases():
    """name -> dict(y_true (B,R,K,2) f32, y_pred (B,H,W,1) f32, batch_size, ranking_size).  NumPy only, deterministic."""
    cases = {}

    def case(name, B, H, W, K, R, seed, sort=True, dup=True, invalid=0, scale=1.5):
        rs = np.random.RandomState(seed)
        pred = (rs.randn(B, H, W, 1) * scale).astype(np.float32)
        idx = rs.randint(0, H * W, size=(B, R, K))
        if dup and K >= 2:
            idx[:, 0, 1] = idx[:, 0, 0]                   # duplicate pixel inside a list: gradients accumulate
        depth = (rs.permutation(B * R * K).reshape(B, R, K) + 0.5) / (B * R * K)      # all labels distinct
        if sort:
            depth = np.sort(depth, axis=2)[:, :, ::-1]    # what the samplers deliver (sampling.py:121-122)
        y_true = np.stack([idx.astype(np.float32), depth.astype(np.float32)], axis=-1)
        if invalid:
            flat = y_true.reshape(-1, K, 2)
            for j in range(0, flat.shape[0], 3):          # labels < 0 are "invalid" for TF-Ranking
                flat[j, rs.randint(0, K, size=min(invalid, K - 1)), 1] = -1.0
        cases[name] = dict(y_true=y_true.astype(np.float32), y_pred=pred, batch_size=B, ranking_size=K)

    case("k5_sorted", 2, 16, 12, 5, 40, 1)
    case("k5_unsorted", 2, 16, 12, 5, 40, 2, sort=False)
    case("k1", 2, 8, 8, 1, 10, 3)
    case("k2", 3, 8, 8, 2, 25, 4)
    case("k3_readme_default", 4, 20, 20, 3, 100, 5)       # PLDepth.py:34 default ranking_size
    case("k50_long_lists", 2, 24, 24, 50, 12, 6)
    case("k130_sweep_size", 1, 32, 32, 130, 4, 7)         # hyperopt/hyperparams.py sweeps ranking_size to 500
    case("k5_wide_scores", 2, 16, 12, 5, 40, 8, scale=12.0)
    case("k6_invalid_labels", 2, 16, 12, 6, 30, 9, invalid=2)
    return cases


def run_reference(reference_root, cases):
    sys.path.insert(0, reference_root)
    import tensorflow as tf
    import tensorflow_ranking as tfr
    from pldepth.losses.nll_loss import HourglassNegativeLogLikelihood
    meta = dict(tensorflow=tf.__version__, tensorflow_ranking=getattr(tfr, "__version__", "?"),
                numpy=np.__version__, reference=os.path.abspath(reference_root))
    out = {}
    for name, c in cases.items():
        loss_obj = HourglassNegativeLogLikelihood(ranking_size=c["ranking_size"], batch_size=c["batch_size"])
        y_true = tf.constant(c["y_true"])
        y_pred = tf.Variable(c["y_pred"])
        with tf.GradientTape() as tape:
            value = loss_obj(y_true, y_pred)
        grad = tape.gradient(value, y_pred)
        per_list = loss_obj._loss.compute_unreduced_loss(y_true, tf.constant(c["y_pred"]))
        out[name] = dict(c, loss=np.float32(value.numpy()), per_list=np.asarray(per_list.numpy(), np.float32).reshape(-1),
                         grad=np.asarray(grad.numpy(), np.float32), **{"meta_" + k: v for k, v in meta.items()})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of praneeth-b/PLDepth")
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    cases = make_cases()
    if args.dry_run:
        for name, c in cases.items():
            print("%-22s y_true %s y_pred %s" % (name, c["y_true"].shape, c["y_pred"].shape))
        return 0
    try:
        res = run_reference(args.reference, cases)
    except ImportError as exc:
        print("pin_tfranking: cannot import the reference's dependencies (%s); nothing written" % exc, file=sys.stderr)
        return 2
    os.makedirs(args.out, exist_ok=True)
    for name, r in res.items():
        path = os.path.join(args.out, "listmle_%s.npz" % name)
        np.savez_compressed(path, **r)
        print("wrote", path, "loss", float(r["loss"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
