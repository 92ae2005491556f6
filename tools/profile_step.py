"""Run the bench workload's forward a few times, then bracket exactly ONE forward with
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees one step only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "act3d"
dev = torch.device("cuda", 0)
if which == "act3d":
    m = bench.build_act3d().to(dev)
    m.seed_ghost_sampler(1)
    ins = [t.to(dev) for t in bench.act3d_inputs(bench.WORKLOAD["batch"], bench.WORKLOAD["ncam"], 100)]
    fn = lambda: m(*ins)
else:
    fn = bench.planner_step_fn(dev)
with torch.no_grad():
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one step")
