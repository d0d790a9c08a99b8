"""nprsph_b200 -- B200-native drop-in for the SPH step of VarunRamakri7/NPR-SPH.

The product is `lib/libnprsph.so` (hand-written sm_100a CUDA behind the C ABI declared in
include/nprsph.h).  This package only holds the ctypes harness used by tests and bench.py.
The directory is named `npr-sph_b200`; import it as `nprsph_b200` (root-level shim).
"""
from .binding import *  # noqa: F401,F403
from . import binding  # noqa: F401
from . import scenes  # noqa: F401
