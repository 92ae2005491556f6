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
elif which == "train":
    from tests.golden import synth
    from model import Act3D
    from act3d_chained_diffuser_b200.losses import keypose_loss
    w = bench.TRAIN_WORKLOAD
    torch.manual_seed(0)
    m = Act3D(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], num_attn_heads=w["heads"],
              gripper_loc_bounds=synth.BOUNDS, num_ghost_points=w["ghost_total"], num_sampling_level=3,
              use_instruction=True).to(dev).train()
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    rgb, pcd, instr, grip = [t.to(dev) for t in bench.act3d_inputs(w["batch"], w["ncam"], 300)]

    def fn():
        out = m(rgb, pcd, instr, grip, gt_action=grip)
        loss = sum(keypose_loss(out, grip).values())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
else:
    fn = bench.planner_step_fn(dev)
with torch.set_grad_enabled(which == "train"):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one step")
