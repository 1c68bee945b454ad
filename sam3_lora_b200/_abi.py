"""argtypes/restype declarations for the libsam3b.so entry points beyond sam3b_gemm.

Kept next to include/sam3b.h on purpose: tests/test_abi.py checks that every function the
header declares is exported by the library and declared here.
"""
from __future__ import annotations

import ctypes as C

# name -> (restype, argtypes)
SIGNATURES: dict[str, tuple] = {}


def declare(lib: C.CDLL) -> None:
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
