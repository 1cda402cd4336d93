"""cora_b200 -- B200-native implementation of the CORA Riemannian-staircase inner loop.

The product is `lib/libcora_b200.so` (hand-written sm_100a CUDA behind the C-ABI of
`include/cora_b200.h`); this package only carries the ctypes binding used by the tests and
bench.py.  There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401
