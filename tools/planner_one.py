"""One compute_trajectory at C3 through the persistent loop kernel (for ncu captures of cd_loop_kernel)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda")
w = bench.PLANNER_WORKLOAD
m = bench.build_planner().to(dev)
ins = [t.to(dev) for t in bench.planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    m.compute_trajectory(*ins)
torch.cuda.synchronize()
print("done")
