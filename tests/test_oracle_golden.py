"""CPU: pin the oracle (oracle/) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  Same machine class / same torch build, so the
expected agreement is at round-off level; tolerances are written per check."""
import os

import pytest
import torch

from oracle import act3d_ref, planner_ref
from oracle.attention import parallel_attention_stack, relative_cross_attn_stack
from oracle.rope import rope3d_table, sinusoidal_embedding
from tests.golden import cases, synth

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def close(a, b, atol, rtol=0.0):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    lim = atol + rtol * b.abs().max().item()
    assert err <= lim, f"max abs err {err:.3e} > {lim:.3e}"


def our_module_state(ctor, kw, trunk_dim):
    """Build OUR drop-in module (same state_dict keys as the reference) and fill it by name."""
    m = ctor(**kw).eval()
    return m


def test_rope_tables():
    g = load("rope")
    for e in (60, 120):
        xyz = synth.points_in_bounds(f"rope.{e}", (2, 5))
        assert synth.checksum(xyz) == g[f"check_{e}"]
        close(rope3d_table(xyz, e), g[f"table_{e}"], 1e-7)
    close(sinusoidal_embedding(torch.tensor([0, 1, 17, 99]), 120), g["sinus_120"], 1e-7)


def _stack_sd(e, heads, layers, prefix=""):
    """state_dict skeleton of a RelativeCrossAttentionModule, filled by name like make_golden does."""
    sd = {}
    for l in range(layers):
        a, f = f"{prefix}attn_layers.{l}.", f"{prefix}ffw_layers.{l}."
        sd[a + "multihead_attn.in_proj_weight"] = torch.empty(3 * e, e)
        sd[a + "multihead_attn.in_proj_bias"] = torch.empty(3 * e)
        sd[a + "multihead_attn.out_proj.weight"] = torch.empty(e, e)
        sd[a + "multihead_attn.out_proj.bias"] = torch.empty(e)
        sd[a + "norm.weight"] = torch.empty(e)
        sd[a + "norm.bias"] = torch.empty(e)
        sd[f + "linear1.weight"] = torch.empty(e, e)
        sd[f + "linear1.bias"] = torch.empty(e)
        sd[f + "linear2.weight"] = torch.empty(e, e)
        sd[f + "linear2.bias"] = torch.empty(e)
        sd[f + "norm.weight"] = torch.empty(e)
        sd[f + "norm.bias"] = torch.empty(e)
    return synth.fill_state_dict(sd)


def test_attention_stack():
    g = load("attention_stack")
    c = cases.small_attention_case()
    assert synth.checksum(c["query"], c["context"], c["q_xyz"], c["c_xyz"]) == g["check"]
    sd = _stack_sd(c["e"], c["heads"], 2)
    qr, cr = rope3d_table(c["q_xyz"], c["e"]), rope3d_table(c["c_xyz"], c["e"])
    got = relative_cross_attn_stack(sd, "", c["heads"], 2, c["query"], c["context"], qr, cr)
    for a, b in zip(got, g["with_rope"]):
        close(a, b, 2e-5)
    got = relative_cross_attn_stack(sd, "", c["heads"], 2, c["query"], c["context"])
    for a, b in zip(got, g["no_rope"]):
        close(a, b, 2e-5)


def _act3d_state(use_instruction):
    """Name-filled weights for the Act3D fixture config, keyed like the reference's state_dict."""
    from model import Act3D                      # our drop-in module: only used as a key/shape template
    kw = dict(cases.ACT3D_KW, use_instruction=use_instruction)
    m = Act3D(**kw).eval()
    cases.install_synth_trunk(m, kw["embedding_dim"])
    sd = m.state_dict()
    synth.fill_state_dict(sd)
    return m, sd, kw


@pytest.mark.parametrize("use_instruction", [False, True])
def test_act3d_forward(use_instruction):
    g = load(f"act3d_c0_instr{int(use_instruction)}")
    m, sd, kw = _act3d_state(use_instruction)
    inp = cases.act3d_inputs(batch=2, ncam=1)
    assert synth.checksum(inp["visible_rgb"][:, :, :, :8, :8], inp["visible_pcd"][:, :, :, :8, :8],
                          inp["curr_gripper"]) == g["check"]
    cfg = act3d_ref.Act3DConfig(
        embedding_dim=60, num_attn_heads=4, gripper_loc_bounds=synth.BOUNDS,
        use_instruction=use_instruction, ghost_points_per_level=kw["num_ghost_points_val"] // 3)
    sampler = synth.make_ghost_sampler(2, cfg.ghost_points_per_level)
    with torch.no_grad():
        out = act3d_ref.act3d_forward(sd, cfg, act3d_ref.trunk_from_module(m), inp["visible_rgb"],
                                      inp["visible_pcd"], inp["instruction"], inp["curr_gripper"],
                                      ghost_sampler=sampler)
    for lvl in range(3):
        close(out["ghost_pcd_pyramid"][lvl], g["ghost_pcd_pyramid"][lvl], 1e-6)
        close(out["visible_pcd_pyramid"][lvl][:, :64], g["visible_pcd_pyramid"][lvl], 1e-6)
        for j in range(2):
            close(out["ghost_pcd_masks_pyramid"][lvl][j], g["ghost_pcd_masks_pyramid"][lvl][j], 1e-4, 1e-5)
        close(out["position_pyramid"][lvl], g["position_pyramid"][lvl], 1e-6)   # same argmax
    close(out["position"], g["position"], 1e-6)
    close(out["rotation"], g["rotation"], 1e-5)
    close(out["gripper"], g["gripper"], 1e-5)
    close(out["query_features"], g["query_features"], 1e-4)


def test_parallel_attention():
    g = load("parallel_attention")
    e, h, b, s1, s2 = 120, 8, 2, 12, 30
    sd = {}
    for l in range(2):
        p = f"layers.{l}."
        for nm in ("adaln_1", "adaln_12", "adaln_ff1"):
            sd[p + nm + ".modulation.1.weight"] = torch.empty(2 * e, e)
            sd[p + nm + ".modulation.1.bias"] = torch.empty(2 * e)
        for nm in ("sa1", "cross_12"):
            sd[p + nm + ".in_proj_weight"] = torch.empty(3 * e, e)
            sd[p + nm + ".in_proj_bias"] = torch.empty(3 * e)
            sd[p + nm + ".out_proj.weight"] = torch.empty(e, e)
            sd[p + nm + ".out_proj.bias"] = torch.empty(e)
        for nm in ("norm_1", "norm_12", "norm_122"):
            sd[p + nm + ".weight"] = torch.empty(e)
            sd[p + nm + ".bias"] = torch.empty(e)
        sd[p + "ffn_12.0.weight"] = torch.empty(4 * e, e)
        sd[p + "ffn_12.0.bias"] = torch.empty(4 * e)
        sd[p + "ffn_12.3.weight"] = torch.empty(e, 4 * e)
        sd[p + "ffn_12.3.bias"] = torch.empty(e)
    synth.fill_state_dict(sd)
    x = synth.normal("pa.x", (b, s1, e))
    ctx = synth.normal("pa.ctx", (b, s2, e))
    xp = synth.uniform("pa.xp", (b, s1, 3), -1, 1)
    cp = synth.uniform("pa.cp", (b, s2, 3), -1, 1)
    sem = synth.normal("pa.sem", (b, s1, e), 0.5)
    t_emb = synth.normal("pa.t", (b, e))
    assert synth.checksum(x, ctx, xp, cp, sem, t_emb) == g["check"]
    mask = torch.zeros(b, s1, dtype=torch.bool)
    mask[1, -3:] = True
    y = parallel_attention_stack(sd, "", h, 2, x, mask, ctx, rope3d_table(xp, e), rope3d_table(cp, e), sem,
                                 t_emb, self_attention=True, rotary=True, use_adaln=True)
    close(y, g["out"], 2e-5)


def _planner_state():
    from model import DiffusionPlanner           # our drop-in module: key/shape template + trunk holder
    m = DiffusionPlanner(**cases.PLANNER_KW).eval()
    cases.install_synth_trunk(m.prediction_head, cases.PLANNER_KW["embedding_dim"])
    sd = m.state_dict()
    synth.fill_state_dict(sd, skip_prefixes=("prediction_head.backbone.",))
    head = {k[len("prediction_head."):]: v for k, v in sd.items() if k.startswith("prediction_head.")}
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS)
    return m, head, cfg


def test_diffusion_head():
    g = load("diffusion_head")
    m, sd, cfg = _planner_state()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.7)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([37, 5])
    assert synth.checksum(traj, cur, goal, t) == g["check"]
    pcd_n = planner_ref.normalize_pos(cfg, inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    with torch.no_grad():
        ctx = planner_ref.encode_context(sd, cfg, act3d_ref.trunk_from_module(m.prediction_head), inp["rgb_obs"],
                                         pcd_n, inp["instruction"], cur, goal)
        out = planner_ref.denoise_once(sd, cfg, ctx, traj, inp["trajectory_mask"], t)
    close(out, g["out"], 5e-5)


def test_planner_100_steps():
    g = load("planner_100step")
    m, sd, cfg = _planner_state()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    assert synth.checksum(inp["curr_gripper"], inp["goal_gripper"]) == g["check"]
    with torch.no_grad():
        traj = planner_ref.compute_trajectory(sd, cfg, act3d_ref.trunk_from_module(m.prediction_head),
                                              inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"],
                                              inp["instruction"], inp["curr_gripper"], inp["goal_gripper"],
                                              noise_fn=synth.NoiseStream("cd"))
    close(traj[..., :3], g["trajectory"][..., :3], 2e-4)
    # quaternion sign is determined; compare directly
    close(traj[..., 3:], g["trajectory"][..., 3:], 5e-4)


# ------------------------------------------------------------------------------------------------ training objective
def test_keypose_loss_and_its_gradients():
    from oracle.losses import keypose_loss
    g = load("keypose_loss")
    for vi, variant in enumerate(cases.LOSS_VARIANTS):
        want = g[f"v{vi}"]
        pred, gt = cases.keypose_loss_case()
        leaves = [m.requires_grad_(True) for lvl in pred["ghost_pcd_masks_pyramid"] for m in lvl]
        assert synth.checksum(gt, *[m.detach() for m in leaves]) == want["check"]
        for k in ("rotation", "gripper", "fine_ghost_pcd_offsets"):
            pred[k].requires_grad_(True)
        losses = keypose_loss(pred, gt, **variant)
        assert set(losses) == set(want["losses"])
        for k, v in losses.items():
            close(v.detach(), want["losses"][k], 1e-6, 1e-6)
        sum(losses.values()).backward()
        for m, ref in zip(leaves, want["dmasks"]):
            close(m.grad if m.grad is not None else torch.zeros_like(m), ref, 1e-7, 1e-5)
        close(pred["rotation"].grad, want["drotation"], 1e-7, 1e-5)
        close(pred["gripper"].grad, want["dgripper"], 1e-7, 1e-5)
        close(pred["fine_ghost_pcd_offsets"].grad, want["doffsets"], 1e-7, 1e-5)


def leaf_state_dict(module):
    leaves, sd = {}, {}
    for k, v in module.state_dict().items():
        if v.data_ptr() not in leaves:
            leaves[v.data_ptr()] = v.detach().clone().requires_grad_(v.is_floating_point())
        sd[k] = leaves[v.data_ptr()]
    return sd


def test_act3d_training_gradients_of_the_oracle_match_the_reference():
    """Autograd through the oracle restatement + oracle loss == autograd through the unmodified reference
    + its own LossAndMetrics (train mode, gt_action given, 96 ghost points / level)."""
    from model import Act3D
    from oracle.losses import keypose_loss
    g = load("act3d_train_grads")
    kw = dict(cases.ACT3D_KW, use_instruction=True, num_ghost_points=3 * 96)
    m = Act3D(**kw).train()
    cases.install_synth_trunk(m, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict())
    inp = cases.act3d_inputs(batch=2, ncam=1)
    gt = cases.keypose_loss_case(batch=2)[1]
    assert synth.checksum(gt, inp["curr_gripper"]) == g["check"]
    sampler = synth.make_ghost_sampler(2, 96)
    cfg = act3d_ref.Act3DConfig(gripper_loc_bounds=synth.BOUNDS, use_instruction=True, ghost_points_per_level=96)
    sd = leaf_state_dict(m)
    out = act3d_ref.act3d_forward(sd, cfg, act3d_ref.trunk_from_module(m), inp["visible_rgb"], inp["visible_pcd"],
                                  inp["instruction"], inp["curr_gripper"], gt_action=gt, ghost_sampler=sampler)
    for a, b in zip(out["position_pyramid"], g["position_pyramid"]):
        assert torch.equal(a, b)
    losses = keypose_loss(out, gt)
    for k, v in losses.items():
        close(v.detach(), g["losses"][k], 1e-5, 1e-5)
    sum(losses.values()).backward()
    checked = 0
    for name, ref in g["grads"].items():
        got = sd[name].grad
        if ref.abs().max() == 0:
            continue
        assert got is not None, name
        err = ((got - ref).norm() / ref.norm()).item()
        assert err <= 1e-4, (name, err)
        checked += 1
    assert checked >= 60, checked


# ------------------------------------------------------------------------------------------------ multi-scale planner
def test_planner_multiscale_head_and_sampling():
    """feat_scales_to_use=3 (find_traj_nn local refinement, untied weights per offset): all three refinements of one
    DiffusionHead.forward and an 8-step sampling loop against the reference."""
    from model import DiffusionPlanner
    g = load("planner_multiscale")
    kw = cases.PLANNER_MS_KW
    m = DiffusionPlanner(**kw).eval()
    cases.install_synth_trunk(m.prediction_head, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    sd = {k[len("prediction_head."):]: v for k, v in m.state_dict().items() if k.startswith("prediction_head.")}
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS, feat_scales_to_use=3, diffusion_timesteps=8)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.4)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([5, 2])
    assert synth.checksum(traj, cur, goal, t, inp["curr_gripper"]) == g["check"]
    trunk = act3d_ref.trunk_from_module(m.prediction_head)
    pcd_n = planner_ref.normalize_pos(cfg, inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    with torch.no_grad():
        ctx = planner_ref.encode_context(sd, cfg, trunk, inp["rgb_obs"], pcd_n, inp["instruction"], cur, goal)
        outs = planner_ref.denoise_all(sd, cfg, ctx, traj, inp["trajectory_mask"], t)
        assert len(outs) == 3
        for got, want in zip(outs, g["head_outs"]):
            close(got, want, 1e-5, 1e-5)
        sampled = planner_ref.compute_trajectory(sd, cfg, trunk, inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"],
                                                 inp["instruction"], inp["curr_gripper"], inp["goal_gripper"],
                                                 noise_fn=synth.NoiseStream("cdms"))
    close(sampled[..., :3], g["trajectory"][..., :3], 1e-4)
    close(sampled[..., 3:], g["trajectory"][..., 3:], 1e-4)


def check_grad_fingerprints(named_grads, want, tol=1e-3):
    """named_grads: name -> tensor; want: name -> synth.grad_summary dict."""
    checked = 0
    for name, w in want.items():
        got = named_grads.get(name)
        assert got is not None, f"{name}: no gradient"
        got = got.detach().float().cpu()
        r = synth.normal("proj:" + name, tuple(got.shape))
        assert abs(got.norm().item() - w["norm"]) <= tol * w["norm"] + 1e-9, (name, got.norm().item(), w["norm"])
        assert abs((got * r).sum().item() - w["proj"]) <= tol * w["norm"] * w["rnorm"] + 1e-9, name
        if w["full"] is not None:
            assert ((got - w["full"]).norm() / w["full"].norm()).item() <= tol, name
        checked += 1
    return checked


def test_planner_training_loss_and_gradients():
    """Oracle training objective + its autograd vs the reference's DiffusionPlanner.forward + backward."""
    from model import DiffusionPlanner
    g = load("planner_train")
    m = DiffusionPlanner(**cases.PLANNER_KW).eval()
    cases.install_synth_trunk(m.prediction_head, 120)
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    head = m.prediction_head
    sd = leaf_state_dict(head)
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    gt = cases.planner_gt_trajectory(batch=2, length=12)
    assert synth.checksum(gt, inp["curr_gripper"]) == g["check"]
    loss = planner_ref.training_loss(sd, cfg, act3d_ref.trunk_from_module(head), gt, inp["trajectory_mask"], inp["rgb_obs"],
                                     inp["pcd_obs"], inp["instruction"], inp["curr_gripper"], inp["goal_gripper"],
                                     noise_fn=synth.NoiseStream("cdtr"), timesteps=torch.tensor([37, 5]))
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    loss.backward()
    grads = {"prediction_head." + k: v.grad for k, v in sd.items() if v.grad is not None}
    assert check_grad_fingerprints(grads, g["grads"], tol=1e-4) >= 200


def test_act3d_variant_branches():
    """6-D rotation from the top ghost point, offset head, untied stacks, 2 cameras, 333 ghost points per level."""
    from model import Act3D
    g = load("act3d_variant")
    kw = dict(cases.ACT3D_KW, use_instruction=True, rotation_parametrization="6D_from_top_ghost",
              regress_position_offset=True, weight_tying=False, gp_emb_tying=False, num_ghost_points_val=3 * 333)
    m = Act3D(**kw).eval()
    cases.install_synth_trunk(m, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict())
    inp = cases.act3d_inputs(batch=2, ncam=2, seed=3)
    assert synth.checksum(inp["curr_gripper"], inp["instruction"][:, :2]) == g["check"]
    cfg = act3d_ref.Act3DConfig(gripper_loc_bounds=synth.BOUNDS, use_instruction=True, ghost_points_per_level=333,
                                rotation_parametrization="6D_from_top_ghost", regress_position_offset=True)
    with torch.no_grad():
        out = act3d_ref.act3d_forward(m.state_dict(), cfg, act3d_ref.trunk_from_module(m), inp["visible_rgb"],
                                      inp["visible_pcd"], inp["instruction"], inp["curr_gripper"],
                                      ghost_sampler=synth.make_ghost_sampler(2, 333, seed=3))
    for a, b in zip(out["position_pyramid"], g["position_pyramid"]):
        assert torch.equal(a, b)
    for lvl in range(3):
        for j in range(2):
            close(out["ghost_pcd_masks_pyramid"][lvl][j], g["ghost_pcd_masks_pyramid"][lvl][j], 1e-4, 1e-5)
    close(out["position"], g["position"], 1e-5)
    close(out["rotation"], g["rotation"], 1e-4)
    close(out["gripper"], g["gripper"], 1e-5)
    close(out["fine_ghost_pcd_offsets"], g["fine_ghost_pcd_offsets"], 1e-5, 1e-5)
    close(out["query_features"], g["query_features"], 1e-4, 1e-5)
