from act3d_chained_diffuser_b200.keypose import Act3D  # noqa: F401
