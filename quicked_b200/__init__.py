"""quicked_b200 — B200-native (sm_100a) implementation of QuickEd's bound-and-align hot path.

The product is the C-ABI shared library `libquicked_b200.so` (include/quicked.h, include/quicked_b200.h);
this package is its Python host-side mirror (ctypes) plus the seeded dataset generator.
"""
from .capi import (BANDED, HIRSCHBERG, QUICKED, WINDOWED, BatchAligner, QuickedAligner, QuickedException,  # noqa: F401
                   generate_pairs_native, load, make_params, pack_pairs)
