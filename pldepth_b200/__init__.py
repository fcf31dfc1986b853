"""pldepth_b200 -- B200-native drop-in for PLDepth's sampling -> gather -> ListMLE hot path.

Host-side mirrors of the reference interfaces (same names, arguments and error behaviour):
  losses.HourglassNegativeLogLikelihood      <- pldepth/losses/nll_loss.py:32-40
  sampling.*MaskedRandomSamplingStrategy,
  sampling.InformationScoreBasedSampling      <- pldepth/data/sampling.py:106-243
  models_meta.ModelParameters                <- pldepth/models/models_meta.py:27-70
All compute runs in hand-written sm_100a kernels behind the C ABI in include/pldepth_b200.h.
"""
from ._lib import PLDError, load_library, launch_count  # noqa: F401

__version__ = "0.1.0"
