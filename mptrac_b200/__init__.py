"""mptrac_b200 -- B200-native (sm_100a) engine for MPTRAC's per-particle time-step path.

The product is the C-ABI library ``mptrac_b200/_lib/libmptrac_b200.so`` (``include/mptrac_b200.h``);
this package is the thin ctypes host mirror used by the tests and the benchmark.  There is no CPU
fallback: creating an :class:`Engine` without a CUDA device raises.
"""
from .host import Ctl, Engine, Met, MpbError, Team, load_library, MIX_MAXQ  # noqa: F401

__all__ = ["Ctl", "Engine", "Met", "MpbError", "Team", "load_library", "MIX_MAXQ"]
