"""mvp_benchmark_b200 — B200-native (sm_100a) point-cloud operators behind the API of
paul007pl/MVP_Benchmark's `utils/metrics` and `utils/mm3d_pn2` packages.

Drop-in use: put `mvp_benchmark_b200/utils` on sys.path ahead of the reference's
`../utils` — `mvp_benchmark_b200.install()` does that — and the reference's
`from metrics import cd, fscore, emd` / `from mm3d_pn2 import furthest_point_sample, ...` resolve here.
"""
import os
import sys

UTILS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "utils")

__version__ = "0.1.0"


def install():
    """Make `metrics` and `mm3d_pn2` importable from this package (ahead of any other copy)."""
    if UTILS_DIR not in sys.path:
        sys.path.insert(0, UTILS_DIR)
    return UTILS_DIR
