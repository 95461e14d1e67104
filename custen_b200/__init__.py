"""cuSten-B200: the cuSten 2D stencil engine re-built for NVIDIA B200 (sm_100a).

The product is the CUDA library under custen_b200/csrc (built to custen_b200/lib/); this package is the thin
Python binding of its C ABI, mirroring the reference's API names.  Importing the package does not load the
library; the first API call does, and fails loudly if it has not been built (there is no CPU path).
"""
from ._lib import LIB_PATH, VARIANTS, EXPORTED, load  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import cuSten_t, Stencil2D, DEVICE, HOST  # noqa: F401

__version__ = "0.1.0"
