"""Drop-in import surface of the reference: ``from model import Act3D, DiffusionPlanner``
(reference: model/__init__.py:1-2).  The implementations live in act3d_chained_diffuser_b200."""
from .keypose_optimization.act3d import Act3D  # noqa: F401
from .trajectory_optimization.diffusion_model import DiffusionPlanner  # noqa: F401
