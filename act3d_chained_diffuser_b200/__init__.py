"""act3d_chained_diffuser_b200 -- B200-native hot path of Act3D / ChainedDiffuser.

The product is the CUDA library ``libact3d_b200.so`` (csrc/, C ABI in include/act3d_b200.h) plus
the thin PyTorch host modules that keep the reference's nn.Module boundary
(``model.Act3D`` / ``model.DiffusionPlanner``: same constructor arguments, forward signatures
and state_dict keys).  There is no CPU or eager fallback: without the library, calls raise.
"""
from .keypose import Act3D  # noqa: F401
from .planner import DiffusionPlanner  # noqa: F401

__all__ = ["Act3D", "DiffusionPlanner"]
