"""abismal_b200: B200-native (sm_100a) implementation of the `abismal map` hot path.

The product is `libabismal_b200.so` (CUDA kernels behind the C ABI declared in
include/abismal_b200.h) plus the `abismal-b200 map` command-line front end
(abismal_b200/bin/abismal-b200), a drop-in for the reference's `abismal map`.
This Python package is a thin ctypes binding used by tests and bench.py; it
never falls back to a CPU implementation: importing `capi` raises if the CUDA
library has not been built.
"""
from .capi import (  # noqa: F401
    AbgError,
    Index,
    Mapper,
    MODE_A_RICH,
    MODE_PAIRED,
    MODE_RANDOM_PBAT,
    lib_path,
    load_library,
)
from .index_file import IndexFile  # noqa: F401
from .reads import ReadBatch, load_fastq  # noqa: F401
