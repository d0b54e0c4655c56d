"""raft_b200 -- B200-native RAO-solve hot path behind the RAFT API.

Importing the package loads ``csrc/libraftk.so`` (sm_100a).  There is no CPU fallback: a missing
library is an ImportError."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library has not been built)
from . import grid, packer, solver  # noqa: F401

__version__ = "0.1.0"
