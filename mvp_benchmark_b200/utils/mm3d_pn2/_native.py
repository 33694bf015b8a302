"""Locates mvp_benchmark_b200._lib when this package is imported top-level (the reference does
`sys.path.append("../utils")`, completion/model_utils.py:19, so `mvp_benchmark_b200` itself may not be
on sys.path)."""
import os
import sys

try:
    from mvp_benchmark_b200 import _lib
except ImportError as first:  # pragma: no cover - path juggling
    _here = os.path.abspath(__file__)
    for _ in range(4):  # <root>/mvp_benchmark_b200/utils/<package>/_native.py
        _here = os.path.dirname(_here)
    if _here in sys.path:
        raise first
    sys.path.insert(0, _here)
    from mvp_benchmark_b200 import _lib

__all__ = ["_lib"]
