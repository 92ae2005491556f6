"""Where does compute_trajectory's time outside the sampling loop go?  Wall clock vs CUDA time per phase."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda")
w = bench.PLANNER_WORKLOAD
m = bench.build_planner().to(dev)
ins = [t.to(dev) for t in bench.planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
for _ in range(3):
    m.compute_trajectory(*ins)
torch.cuda.synchronize()
head = m.prediction_head
orig = head.encode_context
marks = {}


def timed_encode(*a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = orig(*a, **k)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    marks["encode_cpu_ms"] = (t1 - t0) * 1e3
    marks["encode_total_ms"] = (t2 - t0) * 1e3
    return out


head.encode_context = timed_encode
torch.cuda.synchronize()
t0 = time.perf_counter()
m.compute_trajectory(*ins)
torch.cuda.synchronize()
marks["compute_trajectory_ms"] = (time.perf_counter() - t0) * 1e3
print(marks)
from torch.profiler import profile, ProfilerActivity
head.encode_context = orig
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    m.compute_trajectory(*ins)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
