"""CPU oracle for the PLDepth hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with NumPy, the reference's algorithm for the
one path this repository accelerates (ranking sampling -> gather -> ListMLE /
Plackett-Luce NLL forward + backward).  It exists to *check* the CUDA product in
``pldepth_b200``; it is never the thing shipped or measured.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here.  Nothing under
``pldepth_b200/`` imports it, and the product path raises when the CUDA library is
missing instead of falling back to this code.

Parity pins (see DESIGN.md "Oracle"):

* stage 1 (sampler, ``sampler_oracle``): pinned against the UNMODIFIED reference file
  ``/root/reference/pldepth/data/sampling.py`` executed in the build container behind a
  ``tensorflow`` stub (``reference_loader``); its outputs are committed as
  ``tests/golden/sampler_*.npz`` by ``tests/golden/make_golden.py``.
* stages 2-3 (gather + ListMLE + gradient, ``listmle_oracle``): the arithmetic lives in
  the un-vendored third-party package ``tensorflow_ranking==0.3.1`` (requirements.txt:20),
  which is not installable here (no TensorFlow, no network).  The restatement follows the
  published 0.3.1 source of ``losses_impl.ListMLELoss.compute_unreduced_loss`` and is
  anchored on hand-derived Plackett-Luce known-answer vectors, closed-form properties and
  an independent torch-autograd (fp64) gradient -- **parity unpinned** against the real
  TF-Ranking binary.
"""
