"""Launch the fused cross-attention stack a few times at the C2 shape (for ncu captures):
    A3D_XATTN_CORE=4 A3D_XATTN_POLY=0 ncu --set full -k regex:xattn -s 2 -c 1 ... python tools/xattn_one.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from tools.xattn_study import run, setup  # noqa: E402

lib.load()
if os.environ.get("A3D_XATTN_POLY"):
    lib.set_option("xattn_poly", int(os.environ["A3D_XATTN_POLY"]))
if os.environ.get("A3D_XATTN6_NP"):
    lib.set_option("xattn6_np", int(os.environ["A3D_XATTN6_NP"]))
b, nq, nk = 16, 16384, 4150
t = setup(b, nq, nk, 1.0)
_, _, ms = run(b, nq, nk, t, iters=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
print(f"{ms:.3f} ms per launch")
