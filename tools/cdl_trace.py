"""Timeline of one layer of cd_loop_kernel (study build: A3D_NVCC_EXTRA=-DA3D_CDL_TRACE python -m act3d_chained_diffuser_b200.build --force)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from act3d_chained_diffuser_b200 import lib  # noqa: E402

dev = torch.device("cuda")
w = bench.PLANNER_WORKLOAD
m = bench.build_planner().to(dev)
ins = [t.to(dev) for t in bench.planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
m.compute_trajectory(*ins)
torch.cuda.synchronize()
buf = np.zeros(64, dtype=np.int64)
fn = lib.load().cd_loop_trace_read
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
names = {0: "layer begin", 1: "q_ready passed", 2: "key loop done", 3: "partials sent", 4: "p_ready passed", 5: "merge done", 6: "out-proj + LN12 + planes",
         7: "kv_free passed", 8: "V, K GEMMs", 9: "all-gather + Q GEMM", 10: "kv_ready passed", 11: "MHA done", 12: "O GEMM", 14: "LN1 + FFN",
         15: "LN122 (+ regress)", 16: "make_q done"}
prev = buf[0]
for k in sorted(names):
    if buf[k]:
        print(f"{(buf[k] - buf[0]):8d} (+{buf[k] - prev:6d})  {names[k]}")
        prev = buf[k]
