"""GPU: ChainedDiffuser denoiser + 100-step sampling loop through the drop-in module, against the
golden vectors (unmodified reference) and the CPU oracle.  Tolerances:
single denoiser call rel-L2 <= 1e-3; full loop with identical injected noise: position max-abs
<= 1e-3 (world metres), quaternion max-abs <= 2e-3."""
import os

import pytest
import torch

from oracle import act3d_ref, planner_ref
from tests.golden import cases, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def build(**over):
    from model import DiffusionPlanner
    kw = dict(cases.PLANNER_KW, **over)
    m = DiffusionPlanner(**kw).eval()
    cases.install_synth_trunk(m.prediction_head, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    return m, kw


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_diffusion_head_matches_golden():
    g = torch.load(os.path.join(G, "diffusion_head.pt"), weights_only=False)
    m, _ = build()
    m = m.cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.7)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([37, 5])
    assert synth.checksum(traj, cur, goal, t) == g["check"]
    pcd_n = m.normalize_pos(inp["pcd_obs"].cuda().permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()
    with torch.no_grad():
        out = m.prediction_head(traj.cuda(), inp["trajectory_mask"].cuda(), t.cuda(), inp["rgb_obs"].cuda(), pcd_n,
                                cur.cuda(), goal.cuda(), inp["instruction"].cuda())[-1].cpu()
    # masked (padded) waypoints are don't-care in the reference too, but they are still computed: compare all rows
    assert torch.isfinite(out).all()
    assert rel(out, g["out"]) <= 1e-3, rel(out, g["out"])
    assert (out - g["out"]).abs().max() <= 3e-3 * g["out"].abs().max()


def test_planner_100_steps_matches_golden():
    g = torch.load(os.path.join(G, "planner_100step.pt"), weights_only=False)
    m, _ = build()
    m = m.cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    assert synth.checksum(inp["curr_gripper"], inp["goal_gripper"]) == g["check"]
    m._noise_fn = synth.NoiseStream("cd")
    traj = m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                         "curr_gripper", "goal_gripper")]).cpu()
    want = g["trajectory"]
    assert traj.shape == want.shape == (2, 12, 7)
    assert (traj[..., :3] - want[..., :3]).abs().max() <= 1e-3, (traj[..., :3] - want[..., :3]).abs().max()
    qd = torch.minimum((traj[..., 3:] - want[..., 3:]).abs().max(-1).values, (traj[..., 3:] + want[..., 3:]).abs().max(-1).values)
    assert qd.max() <= 2e-3, qd.max()
    assert torch.allclose(traj[..., 3:].norm(dim=-1), torch.ones(2, 12), atol=1e-5)


def test_denoiser_vs_oracle_full_length_multicam():
    """L = 50 waypoints, 2 cameras, no padding mask, goal conditioning at test time, untied weights."""
    m, kw = build(weight_tying=False, use_goal_at_test=True)
    inp = cases.planner_inputs(batch=3, ncam=2, length=50, seed=4)
    head_sd = {k[len("prediction_head."):]: v for k, v in m.state_dict().items() if k.startswith("prediction_head.")}
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS, use_goal_at_test=True)
    with torch.no_grad():
        want = planner_ref.compute_trajectory(head_sd, cfg, act3d_ref.trunk_from_module(m.prediction_head),
                                              inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"], inp["instruction"],
                                              inp["curr_gripper"], inp["goal_gripper"], noise_fn=synth.NoiseStream("v"),
                                              n_steps=100)
    m = m.cuda()
    m._noise_fn = synth.NoiseStream("v")
    got = m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                        "curr_gripper", "goal_gripper")]).cpu()
    assert (got[..., :3] - want[..., :3]).abs().max() <= 1e-3
    qd = torch.minimum((got[..., 3:] - want[..., 3:]).abs().max(-1).values, (got[..., 3:] + want[..., 3:]).abs().max(-1).values)
    assert qd.max() <= 2e-3
    # inpainting: first waypoint = current pose, last = goal pose (position part)
    assert (got[:, 0, :3] - inp["curr_gripper"][:, :3]).abs().max() <= 1e-5
    assert (got[:, -1, :3] - inp["goal_gripper"][:, :3]).abs().max() <= 1e-5


@pytest.mark.parametrize("length,ncam,masked_tail,batch", [(12, 1, 3, 2), (50, 2, 0, 3), (64, 1, 5, 1), (17, 1, 0, 5)])
def test_persistent_loop_matches_the_launch_per_layer_path(length, ncam, masked_tail, batch):
    """csrc/cd_loop.cu (one persistent cluster kernel for all 100 steps, trajectory state in shared memory) against the
    launch-per-layer kernels of csrc/cd_denoiser.cu on identical noise: same math, different tiling / summation order."""
    inp = cases.planner_inputs(batch=batch, ncam=ncam, length=length, masked_tail=masked_tail, seed=11)
    outs = []
    for persistent in (True, False):
        m, _ = build()
        m = m.cuda()
        m.persistent_loop = persistent
        m.use_cuda_graph = False
        m._noise_fn = synth.NoiseStream("pl")
        outs.append(m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                                   "curr_gripper", "goal_gripper")]).cpu())
    a, b = outs
    assert torch.isfinite(a).all()
    assert (a[..., :3] - b[..., :3]).abs().max() <= 3e-4, (a[..., :3] - b[..., :3]).abs().max()
    qd = torch.minimum((a[..., 3:] - b[..., 3:]).abs().max(-1).values, (a[..., 3:] + b[..., 3:]).abs().max(-1).values)
    assert qd.max() <= 1e-3, qd.max()


def test_persistent_loop_without_instruction_and_is_deterministic():
    m, _ = build(use_instruction=False)
    m = m.cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=20, seed=2)
    args = [inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction", "curr_gripper", "goal_gripper")]
    m._noise_fn = synth.NoiseStream("pd")
    a = m.compute_trajectory(*args).cpu()
    m._noise_fn = synth.NoiseStream("pd")
    b = m.compute_trajectory(*args).cpu()
    assert torch.equal(a, b)
    m.persistent_loop = False
    m.use_cuda_graph = False
    m._noise_fn = synth.NoiseStream("pd")
    c = m.compute_trajectory(*args).cpu()
    assert (a[..., :3] - c[..., :3]).abs().max() <= 3e-4


def test_training_loss_fused_and_differentiable_paths_agree():
    """Same noise / timesteps: the loss evaluated under no_grad (fused denoiser kernels) equals the loss of the
    differentiable path (csrc/a3d_train.cu behind autograd); eval mode so that dropout is off in both."""
    m, _ = build()
    m = m.cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12)
    gt = torch.cat([synth.points_in_bounds("gt.p", (2, 12)), torch.nn.functional.normalize(synth.normal("gt.q", (2, 12, 4)), dim=-1)], -1)
    args = [gt.cuda()] + [inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction", "curr_gripper", "goal_gripper")]
    torch.manual_seed(3)
    with torch.no_grad():
        fused = m(*args)
    torch.manual_seed(3)
    diff = m(*args)
    assert fused.dim() == 0 and torch.isfinite(fused) and diff.requires_grad
    assert abs(fused.item() - diff.item()) <= 1e-3 * abs(diff.item())


def test_parallel_heads_are_bit_identical_to_the_sequential_order():
    """Position / rotation heads on two streams (and inside the captured CUDA graph) == sequential launches."""
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    outs = []
    for parallel, graph in ((False, False), (True, False), (True, True)):
        m, _ = build()
        m = m.cuda()
        m.prediction_head.parallel_heads = parallel
        m.use_cuda_graph = graph
        m.persistent_loop = False                      # this test is about the launch-per-layer path
        m._noise_fn = synth.NoiseStream("cd")
        outs.append(m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                                   "curr_gripper", "goal_gripper")]).cpu())
        if graph:       # replay of the captured graph on a second call
            m._noise_fn = synth.NoiseStream("cd")
            again = m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                                   "curr_gripper", "goal_gripper")]).cpu()
            assert torch.equal(again, outs[-1])
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], outs[2])


# ------------------------------------------------------------------------------------------------ multi-scale (a17)
def build_ms():
    from model import DiffusionPlanner
    kw = cases.PLANNER_MS_KW
    m = DiffusionPlanner(**kw).eval()
    cases.install_synth_trunk(m.prediction_head, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    return m


def test_multiscale_head_matches_golden():
    """feat_scales_to_use=3: coarse map, then the 64*L and 16*L fine points nearest the running estimate
    (find_traj_nn -> a3d_traj_topk), untied weights per offset.  All three refinements vs the reference."""
    g = torch.load(os.path.join(G, "planner_multiscale.pt"), weights_only=False)
    m = build_ms().cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.4)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([5, 2])
    assert synth.checksum(traj, cur, goal, t, inp["curr_gripper"]) == g["check"]
    pcd_n = m.normalize_pos(inp["pcd_obs"].cuda().permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()
    with torch.no_grad():
        outs = m.prediction_head(traj.cuda(), inp["trajectory_mask"].cuda(), t.cuda(), inp["rgb_obs"].cuda(), pcd_n,
                                 cur.cuda(), goal.cuda(), inp["instruction"].cuda())
    assert len(outs) == 3
    for i, (got, want) in enumerate(zip(outs, g["head_outs"])):
        got = got.cpu()
        assert torch.isfinite(got).all()
        assert rel(got, want) <= 1e-3, (i, rel(got, want))


def test_multiscale_sampling_matches_golden():
    g = torch.load(os.path.join(G, "planner_multiscale.pt"), weights_only=False)
    m = build_ms().cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    m._noise_fn = synth.NoiseStream("cdms")
    traj = m.compute_trajectory(*[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction",
                                                         "curr_gripper", "goal_gripper")]).cpu()
    want = g["trajectory"]
    assert (traj[..., :3] - want[..., :3]).abs().max() <= 2e-3
    qd = torch.minimum((traj[..., 3:] - want[..., 3:]).abs().amax(-1), (traj[..., 3:] + want[..., 3:]).abs().amax(-1))
    assert qd.max() <= 4e-3
