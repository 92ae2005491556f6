import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act3d_chained_diffuser_b200 import lib
from tools.xattn_study import run, setup
lib.load()
b, nq = 16, 16384
for nk in (64, 128, 1088, 4150):
    t = setup(b, nq, nk, 1.0)
    for core in (2, 4):
        lib.set_option("xattn_core", core)
        run(b, nq, nk, t, iters=2)
        _, _, ms = run(b, nq, nk, t, iters=5)
        print("nk", nk, "tiles", (nk + 63) // 64, "core", core, "ms", round(ms, 4), flush=True)
