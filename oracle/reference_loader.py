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
