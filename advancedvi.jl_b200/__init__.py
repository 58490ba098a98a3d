"""advancedvi.jl_b200 -- B200-native ELBO-gradient path behind AdvancedVI.jl's objective interface.

The directory name carries a dot, so import it as `advancedvi_jl_b200` (a shim package at the
repository root) or through importlib.  Importing needs the built CUDA library
(advancedvi.jl_b200/libavi_b200.so); there is no CPU fallback.
"""
from .api import *          # noqa: F401,F403
from .api import __all__    # noqa: F401
from . import _lib, parallel  # noqa: F401
