"""twilight_b200 — B200-native (sm_100a) implementation of TWILIGHT's per-level TALCO-XDrop alignment path.

The product is the CUDA library `libtwilight_b200.so` (C ABI in include/twilight_b200.h); this package holds its
sources (csrc/), the build recipe, the ctypes binding and the host-side mirror of the reference's level-kernel
interface. There is no CPU implementation in this package.
"""
from .api import (Context, LevelOut, LevelPairIn, NodeSideIn, PairOut, ProfilePairIn, TwilightError,  # noqa: F401
                  nucleotide_matrix)
