"""pysamstats_b200 -- B200-native pileup and binned statistics engine behind the public API
of alimanfoo/pysamstats (``import pysamstats_b200 as pysamstats``).

Importing the package does not touch the GPU; the first ``stat_*`` / ``load_*`` call loads
``csrc/libpss_b200.so`` and creates a CUDA context, and fails loudly if either is missing
(there is no CPU fallback).
"""
__version__ = '1.1.2+b200.0'

from .pileup import *   # noqa: F401,F403
from .binned import *   # noqa: F401,F403
from . import config, opt, util, pileup, binned, bamio, io   # noqa: F401
