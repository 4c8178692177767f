"""poppunk-mod_b200: B200-native (sm_100a) KDE discrepancy for PopPUNK-mod.

The directory name carries a hyphen (it mirrors the upstream repository name), so import it
through the alias module `poppunk_mod_b200` at the repository root, or put this directory on
sys.path and `import KDE_distance` exactly as the reference's callers do.
"""
from . import KDE_distance  # noqa: F401
from .KDE_distance import (KDE_JS_divergence, KDE_JS_divergence_batch, KDE_JS_divergence_batch_text,  # noqa: F401
                           KDE_JS_divergence_pairs,
                           KDE_KL_divergence, default_device, get_grid, kde_density, scale_KDE)
from . import _native  # noqa: F401
from . import distributed, simulator, sweeps  # noqa: F401

__version__ = "0.1.0"
