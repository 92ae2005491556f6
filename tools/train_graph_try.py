"""Act3D training step (bench.TRAIN_WORKLOAD): eager vs whole-step CUDA graph, 1 GPU."""
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
from act3d_chained_diffuser_b200.train_graph import GraphedTrainStep, freeze_parameters_without_gradient  # noqa: E402

dev = torch.device("cuda")
w = bench.TRAIN_WORKLOAD


def make():
    torch.manual_seed(0)
    m = Act3D(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], num_attn_heads=w["heads"],
              gripper_loc_bounds=synth.BOUNDS, num_ghost_points=w["ghost_total"], num_sampling_level=3,
              use_instruction=True).to(dev).train()
    m.seed_ghost_sampler(99)
    return m


rgb, pcd, instr, grip = [t.to(dev) for t in bench.act3d_inputs(w["batch"], w["ncam"], seed=300)]
gt = grip.clone()
gt[:, :3] += 0.02


def step_loss(net, rgb, pcd, instr, grip, gt):
    out = net(rgb, pcd, instr, grip, gt_action=gt)
    return sum(keypose_loss(out, gt).values())


ins = (rgb, pcd, instr, grip, gt)
m = make()
print("frozen:", freeze_parameters_without_gradient(m, lambda net: step_loss(net, *ins)))
torch.cuda.set_sync_debug_mode("warn")
step_loss(m, *ins).backward()
torch.cuda.set_sync_debug_mode("default")
g = GraphedTrainStep(m, step_loss, lambda ps: torch.optim.AdamW(ps, lr=1e-4, capturable=True), ins, warmup=3)
losses = []
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    losses.append(g(*ins))
    losses[-1] = losses[-1].clone()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"graphed: {dt * 1e3:.2f} ms/step, {w['batch'] / dt:.1f} keyframes/s; loss {losses[0].item():.4f} -> {losses[-1].item():.4f}")
