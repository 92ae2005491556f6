"""Timeline of one CTA of xattn6_kernel (study build: A3D_NVCC_EXTRA=-DA3D_X6_TRACE python -m ...build --force).
Prints, per unit, when the issuing warp issued S, when it got P, and when the row warps waited / computed."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from tools.xattn_study import run, setup  # noqa: E402

L = lib.load()
lib.set_option("xattn_core", 6)
lib.set_option("xattn6_np", int(os.environ.get("A3D_XATTN6_NP", "6")))
b, nq, nk = 16, 16384, 4150
t = setup(b, nq, nk, 1.0)
run(b, nq, nk, t, iters=1)
torch.cuda.synchronize()
buf = np.zeros((3, 4096), dtype=np.uint64)
fn = L.a3d_x6_trace_read
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
ev = {}
t0 = None
for role in range(3):
    for w in buf[role]:
        w = int(w)
        if w == 0:
            continue
        tag, idx, clk = w >> 56, (w >> 40) & 0xffff, w & 0xffffffffff
        ev.setdefault((tag, idx, role), clk)       # first layer only (later layers overwrite nothing: setdefault)
        t0 = clk if t0 is None else min(t0, clk)
print("unit  S_issued  P_wait  P_got | rows0: arrive  S_ready ld_done st_issued | rows1: arrive S_ready ld_done st_issued")
for u in range(8, 60):
    g = lambda tag, role: ev.get((tag, u, role), t0) - t0
    print(f"{u:4d} {g(1,0):8d} {g(2,0):8d} {g(3,0):8d} | {g(4,1):8d} {g(5,1):8d} {g(6,1):8d} {g(7,1):8d} | {g(4,2):8d} {g(5,2):8d} {g(6,2):8d} {g(7,2):8d}")

names = {9: "kernel start", 10: "layer begin", 11: "q GEMM result", 12: "Q written", 13: "first tile done", 14: "key loop done",
         15: "verdict done", 16: "O -> A done", 17: "out-proj result", 18: "LN1 + A done", 19: "FFN1 result", 20: "hid -> A done",
         21: "FFN2 result", 22: "LN2 (+A) done", 23: "teardown"}
ph = []
for w in buf[1][3000:]:
    w = int(w)
    if w:
        ph.append((w & 0xffffffffff, w >> 56, (w >> 40) & 0xffff))
ph.sort()
base = ph[0][0]
prev = base
for clk, tag, layer in ph:
    print(f"{clk - base:9d} (+{clk - prev:7d})  layer {layer}  {names.get(tag, tag)}")
    prev = clk
