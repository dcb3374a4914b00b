"""oxli_b200 -- B200-native (sm_100a) k-mer counting behind oxli's KmerCountTable API.

`KmerCountTable` is the compiled drop-in class (csrc/pyoxli.cpp over the C ABI
in include/oxli_b200.h).  There is no CPU fallback: importing without the built
extension raises ImportError, and every call fails without a CUDA device.
"""
from __future__ import annotations

__version__ = "0.3.0"

try:
    from ._oxli import KmerCountTable  # noqa: F401
except ImportError as e:  # the C-ABI layer alone is still usable (bench, tests)
    _binding_error = e

    def __getattr__(name):
        if name == "KmerCountTable":
            raise ImportError(
                "oxli_b200._oxli is not built; run `python -m oxli_b200._build` "
                f"(original error: {_binding_error})"
            )
        raise AttributeError(name)
