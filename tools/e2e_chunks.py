"""Act3D C2 end to end from pinned host inputs (model(host tensors) -> action on the host): eager staged path vs.
CUDA-graph replay with the image upload / trunk pipelined in 1, 2, 4, 8 pieces."""
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
host = [t.pin_memory() for t in bench.act3d_inputs(16, 4, 100)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(n):
    tot = 0.0
    with torch.no_grad():
        for i in range(n + 3):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = m(*host)
            res = torch.cat([out["position"], out["rotation"].reshape(16, -1), out["gripper"]], -1).cpu()
            e.record()
            torch.cuda.synchronize()
            if i >= 3:
                tot += s.elapsed_time(e)
    return tot / n


m.use_cuda_graph = False
print(json.dumps({"mode": "eager staged", "ms": round(run(10), 3)}), flush=True)
m.use_cuda_graph = True
for c in (1, 2, 4, 8):
    m.upload_chunks = c
    ms = run(10)
    print(json.dumps({"mode": f"graph, {c} piece(s)", "ms": round(ms, 3), "keyframes_per_s": round(16 / ms * 1e3, 1)}), flush=True)
