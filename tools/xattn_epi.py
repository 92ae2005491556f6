"""Launch the fused cross-attention stack with ONE key tile (nk = 64): the fixed, per-launch part (q-proj, first
tile, epilogues) dominates -- for ncu captures of that part."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from tools.xattn_study import run, setup  # noqa: E402

lib.load()
b, nq, nk = 16, 16384, 64
t = setup(b, nq, nk, 1.0)
_, _, ms = run(b, nq, nk, t, iters=4)
print(f"{ms:.3f} ms per launch")
