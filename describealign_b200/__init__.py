"""describealign_b200 - B200-native alignment hot path for julbean/describealign.

Drop-in replacements for the reference's get_energy / get_zero_crossings / get_freq_bands /
align (reference describealign.py:545-1027), implemented as thin Python host code over
hand-written sm_100a CUDA kernels behind a C ABI (include/describealign_b200.h).

Importing this package does not touch CUDA; the first call into the library does.
"""
from .api import (align, align_pcm, get_energy, get_freq_bands, get_zero_crossings,  # noqa: F401
                  track_features)
from .launcher import patch  # noqa: F401

__version__ = "0.1.0"
