"""camb_b200 -- B200-native per-k-mode Boltzmann path for CAMB."""
from .host import CAMBError, HostState, default_params  # noqa: F401

__version__ = "0.1.0"
