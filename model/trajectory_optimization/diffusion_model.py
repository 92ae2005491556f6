from act3d_chained_diffuser_b200.planner import DiffusionPlanner, DiffusionHead  # noqa: F401
