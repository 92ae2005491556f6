"""Act3D / planner training step (bench.TRAIN_WORKLOAD, 1 GPU, no DDP): torch-profiler table of where the step goes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tests.golden import synth  # noqa: E402
from model import Act3D  # noqa: E402
from act3d_chained_diffuser_b200.losses import keypose_loss  # noqa: E402

dev = torch.device("cuda")
w = bench.TRAIN_WORKLOAD
torch.manual_seed(0)
model = Act3D(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], num_attn_heads=w["heads"],
              gripper_loc_bounds=synth.BOUNDS, num_ghost_points=w["ghost_total"], num_sampling_level=3,
              use_instruction=True).to(dev).train()
model.seed_ghost_sampler(99)
rgb, pcd, instr, grip = [t.to(dev) for t in bench.act3d_inputs(w["batch"], w["ncam"], seed=300)]
gt = grip.clone()
gt[:, :3] += 0.02
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)


def step():
    out = model(rgb, pcd, instr, grip, gt_action=gt)
    loss = sum(keypose_loss(out, gt).values())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"step: host {1e3 * (t1 - t0) / 5:.2f} ms, wall {1e3 * (t2 - t0) / 5:.2f} ms ({w['batch'] * 5 / (t2 - t0):.1f} keyframes/s)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ("attn_" in ev.name or "wgrad" in ev.name or "layer_norm" in ev.name):
        name = ev.name
        for key in ("attn_fwd", "attn_bwd_dq", "attn_bwd_dkv", "wgrad_partial", "wgrad_reduce", "layer_norm_grad_input", "GammaBeta", "vectorized_layer_norm"):
            if key in name:
                rows.append((ev.time_range.start, key, ev.device_time))
rows.sort()
print("\nper-launch (in order): kernel, us")
for _, n, t in rows:
    print(f"{n:42s} {t:9.1f}")
