"""Act3D C2 forward time vs batch (strong-scaling shards): eager launches vs CUDA-graph replay."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda")
m = bench.build_act3d().to(dev)
m.seed_ghost_sampler(1)
full = [t.to(dev) for t in bench.act3d_inputs(16, 4, 100)]
for b in (16, 8, 4, 2, 1):
    ins = [t[:b].contiguous() for t in full]
    row = {"batch": b}
    for graph in (False, True):
        m.use_cuda_graph = graph
        with torch.no_grad():
            for _ in range(4):
                m(*ins)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                m(*ins)
            e.record()
            torch.cuda.synchronize()
        row["graph_ms" if graph else "eager_ms"] = round(s.elapsed_time(e) / 10, 3)
    row["keyframes_per_s_graph"] = round(b / row["graph_ms"] * 1e3, 1)
    print(json.dumps(row), flush=True)
