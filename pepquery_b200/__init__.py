"""pepquery_b200 — B200-native (sm_100a CUDA) implementation of PepQuery's pair-scoring hot path.

The product is the C-ABI shared library ``libpepquery_b200.so`` (include/pepquery_b200.h).  This package is the
thin host-side mirror of the reference's Java entry points on top of that ABI, used by the tests and the bench
(no JVM exists in the build image; INTEGRATION.md shows the JNI / Panama FFM binding a PepQuery maintainer adds).

There is no CPU fallback: importing works anywhere, but creating a search context without the CUDA library or
without a GPU raises.
"""
from . import _abi  # noqa: F401
from .host import (CParameter, PepQueryError, Search, library_path, load_library, parse_mgf,  # noqa: F401
                   scorePeptide2Spectrum)

__all__ = ["CParameter", "PepQueryError", "Search", "library_path", "load_library", "parse_mgf",
           "scorePeptide2Spectrum", "_abi"]
