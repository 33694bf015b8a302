"""Drop-in for the reference's utils/metrics/CD/__init__.py:1-4."""
from .chamfer3D.dist_chamfer_3D import chamfer_3DDist as cd
from .fscore import fscore

__all__ = ['cd', 'fscore']
