"""Import the UNMODIFIED reference sampler from /root/reference behind a TensorFlow stub.

Test infrastructure (see oracle/__init__.py).  Works only where /root/reference exists
(the build container); the GPU box never has it, so nothing run there may call this.

Why a stub: pldepth/data/sampling.py:4 imports pldepth/data/depth_utils.py, which does
``import tensorflow as tf`` and ``import tensorflow.python.keras.backend`` at module top
(depth_utils.py:1-2) and uses ``tf.float32`` as a default argument (depth_utils.py:24).
TensorFlow is not installed here, so four empty modules are registered first.  No
reference source is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pldepth", "data", "sampling.py"))


def load_reference_sampling():
    """Return the reference module ``pldepth.data.sampling`` (unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "tensorflow" not in sys.modules:
        names = ["tensorflow", "tensorflow.python", "tensorflow.python.keras",
                 "tensorflow.python.keras.backend"]
        mods = {n: types.ModuleType(n) for n in names}
        mods["tensorflow"].float32 = "float32"
        mods["tensorflow"].python = mods["tensorflow.python"]
        mods["tensorflow.python"].keras = mods["tensorflow.python.keras"]
        mods["tensorflow.python.keras"].backend = mods["tensorflow.python.keras.backend"]
        mods["tensorflow"].__pld_stub__ = True
        sys.modules.update(mods)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pldepth.data.sampling as ref_sampling  # noqa: E402
    return ref_sampling


class DictModelParams(object):
    """Minimal stand-in for pldepth/models/models_meta.py:27-39 (get_parameter(name, default))."""

    def __init__(self, **kw):
        self.parameters = dict(kw)

    def get_parameter(self, name, default=None):
        return self.parameters.get(name, default)

    def set_parameter(self, name, value):
        self.parameters[name] = value


class ListDataset(object):
    """Stand-in for the ``tf.data.Dataset`` of ``(image, gt)`` elements the evaluation providers iterate
    (generic_ranking_provider.py:86,181): a list with ``as_numpy_iterator``; the stub's
    ``tf.data.experimental.cardinality`` returns its length."""

    def __init__(self, elements):
        self.elements = list(elements)

    def as_numpy_iterator(self):
        return iter(self.elements)

    def __len__(self):
        return len(self.elements)


def load_reference_eval_providers():
    """Return the UNMODIFIED reference module ``pldepth.data.providers.generic_ranking_provider``.

    On top of the sampler's stub it needs ``tf.data.experimental.cardinality`` (generic_ranking_provider.py:83,181)
    and, for ``generate_rankings``, the alias ``np.int`` that NumPy removed in 1.24 (line 189; the reference pins
    numpy~=1.19.5) -- the alias is installed on the numpy module, no reference source is touched."""
    load_reference_sampling()
    import numpy as np
    tf = sys.modules["tensorflow"]
    if getattr(tf, "__pld_stub__", False) and not hasattr(tf, "data"):
        data = types.ModuleType("tensorflow.data")
        exp = types.ModuleType("tensorflow.data.experimental")
        exp.cardinality = lambda ds: len(ds)
        data.experimental = exp
        tf.data = data
        sys.modules["tensorflow.data"] = data
        sys.modules["tensorflow.data.experimental"] = exp
    if not hasattr(np, "int"):
        np.int = int
    import pldepth.data.providers.generic_ranking_provider as ref_eval  # noqa: E402
    return ref_eval
