"""Time DiffusionPlanner.compute_trajectory at C3 (batch 32, 50 waypoints, 100 steps, 4 views) for the persistent
cluster kernel and the launch-per-layer path; splits the one-off context encode from the sampling loop."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda")
    w = bench.PLANNER_WORKLOAD
    m = bench.build_planner().to(dev)
    ins = [t.to(dev) for t in bench.planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
    for persistent, graph in ((True, False), (False, True)):
        m.persistent_loop, m.use_cuda_graph = persistent, graph
        for _ in range(2):
            m.compute_trajectory(*ins)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            m.compute_trajectory(*ins)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        print(json.dumps({"persistent_loop": persistent, "cuda_graph": graph, "ms_per_trajectory_batch": round(ms, 2),
                          "denoise_steps_per_s": round(w["batch"] * w["steps"] / (ms * 1e-3), 1)}), flush=True)


if __name__ == "__main__":
    main()
