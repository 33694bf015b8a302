"""Drop-in for the reference's utils/metrics/__init__.py:1-6."""
from .CD import cd, fscore
from .EMD import emd

__all__ = ['cd', 'fscore', 'emd']
