"""Run the persistent sampling loop twice on identical inputs / noise and report the difference (study tool)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden import cases, synth  # noqa: E402
from model import DiffusionPlanner  # noqa: E402

bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
length = int(sys.argv[3]) if len(sys.argv) > 3 else 50
ncam = int(sys.argv[4]) if len(sys.argv) > 4 else 4
m = DiffusionPlanner(**dict(cases.PLANNER_KW, diffusion_timesteps=steps)).eval()
cases.install_synth_trunk(m.prediction_head, 120)
synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
m = m.cuda()
inp = cases.planner_inputs(batch=bsz, ncam=ncam, length=length, seed=3)
args = [inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction", "curr_gripper", "goal_gripper")]
outs = []
for i in range(3):
    m._noise_fn = synth.NoiseStream("det")
    outs.append(m.compute_trajectory(*args).clone())
torch.cuda.synchronize()
print("persistent run 0 vs 1 max diff", (outs[0] - outs[1]).abs().max().item(), " 0 vs 2", (outs[0] - outs[2]).abs().max().item())
bad = (outs[0] != outs[1]).any(-1).any(-1).nonzero().flatten().tolist()
print("samples that differ:", bad[:40])
m.persistent_loop = False
m.use_cuda_graph = False
m._noise_fn = synth.NoiseStream("det")
ref = m.compute_trajectory(*args)
print("persistent vs launch-per-layer max diff", (outs[0] - ref).abs().max().item())
