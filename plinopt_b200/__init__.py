"""plinopt_b200 -- B200-native candidate-search engine for PLinOpt's hot path.

The product is the C-ABI shared library (include/plinopt_b200.h, built from
plinopt_b200/csrc by plinopt_b200/build.py).  This package is the thin Python
view used by tests and bench.py: `capi` (ctypes), `hm` (SMS / scaling helpers)
and `sharding` (index-range sharding + the single min-allreduce)."""
from . import capi, hm  # noqa: F401

__all__ = ["capi", "hm"]
