"""cd_loop_kernel: L2 prefetch distance of the K/V tile stream vs. time of compute_trajectory at C3."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from act3d_chained_diffuser_b200 import lib  # noqa: E402

dev = torch.device("cuda")
w = bench.PLANNER_WORKLOAD
m = bench.build_planner().to(dev)
ins = [t.to(dev) for t in bench.planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
for pf in [int(x) for x in (sys.argv[1:] or ["12", "8", "6", "4", "3", "2", "0", "12"])]:
    lib.set_option("cd_prefetch_tiles", pf)
    for _ in range(2):
        m.compute_trajectory(*ins)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        m.compute_trajectory(*ins)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 4
    print(json.dumps({"prefetch_tiles": pf, "ms_per_trajectory_batch": round(ms, 2),
                      "denoise_steps_per_s": round(w["batch"] * w["steps"] / (ms * 1e-3), 1)}), flush=True)
