"""Drop-in for the reference's utils/metrics/EMD/__init__.py:1-3."""
from .emd_module import emdModule as emd

__all__ = ['emd']
