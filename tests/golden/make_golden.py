"""Generate the golden vectors from the UNMODIFIED reference (build container only).

    python -m tests.golden.make_golden          # from the repo root; needs /root/reference

The reference is imported from /root/reference through oracle/ref_import.py (import stubs for
the two missing packages ``clip`` and ``diffusers``; the latter is our DDPM restatement, see
oracle/ddpm.py "parity unpinned").  Inputs and weights come from tests/golden/{synth,cases}.py
(name-keyed numpy PCG64 streams), so the tests can rebuild them anywhere; the fixture files
hold only the reference's OUTPUTS plus a checksum of the inputs.  CPU, fp32, eval mode.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402
from tests.golden import cases, synth  # noqa: E402


def save(name, payload):
    path = os.path.join(HERE, name + ".pt")
    torch.save(payload, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


@torch.no_grad()
def gen_rope(ref):
    out = {}
    for e in (60, 120):
        xyz = synth.points_in_bounds(f"rope.{e}", (2, 5))
        out[f"table_{e}"] = ref.position_encodings.RotaryPositionEncoding3D(e)(xyz)
        out[f"check_{e}"] = synth.checksum(xyz)
    t = torch.tensor([0, 1, 17, 99])
    out["sinus_120"] = ref.position_encodings.SinusoidalPosEmb(120)(t)
    save("rope", out)


@torch.no_grad()
def gen_attention_stack(ref):
    c = cases.small_attention_case()
    mod = ref.layers.RelativeCrossAttentionModule(c["e"], c["heads"], 2).eval()
    synth.fill_state_dict(mod.state_dict())
    pe = ref.position_encodings.RotaryPositionEncoding3D(c["e"])
    with_rope = mod(c["query"], c["context"], pe(c["q_xyz"]), pe(c["c_xyz"]))
    no_rope = mod(c["query"], c["context"], None, None)
    save("attention_stack", dict(with_rope=with_rope, no_rope=no_rope,
                                 check=synth.checksum(c["query"], c["context"], c["q_xyz"], c["c_xyz"])))


@torch.no_grad()
def gen_act3d(ref, use_instruction):
    kw = dict(cases.ACT3D_KW, use_instruction=use_instruction)
    torch.manual_seed(0)
    model = ref.Act3D(**kw).eval()
    cases.install_synth_trunk(model, kw["embedding_dim"])
    synth.fill_state_dict(model.state_dict())
    inp = cases.act3d_inputs(batch=2, ncam=1)
    sampler = synth.make_ghost_sampler(2, model.num_ghost_points_val, diameter=kw["fine_sampling_ball_diameter"])
    model._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor)
    out = model(inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"])
    keep = dict(
        position=out["position"], rotation=out["rotation"], gripper=out["gripper"],
        position_pyramid=out["position_pyramid"],
        ghost_pcd_masks_pyramid=out["ghost_pcd_masks_pyramid"],
        ghost_pcd_pyramid=out["ghost_pcd_pyramid"],
        query_features=out["query_features"],
        visible_pcd_pyramid=[p[:, :64].clone() for p in out["visible_pcd_pyramid"]],
        check=synth.checksum(inp["visible_rgb"][:, :, :, :8, :8], inp["visible_pcd"][:, :, :, :8, :8],
                             inp["curr_gripper"]),
    )
    save(f"act3d_c0_instr{int(use_instruction)}", keep)


@torch.no_grad()
def gen_parallel_attention(ref):
    e, h, b, s1, s2 = 120, 8, 2, 12, 30
    mod = ref.layers.ParallelAttention(num_layers=2, d_model=e, n_heads=h, self_attention1=True,
                                       self_attention2=False, cross_attention1=True, cross_attention2=False,
                                       rotary_pe=True, use_adaln=True).eval()
    synth.fill_state_dict(mod.state_dict())
    pe = ref.position_encodings.RotaryPositionEncoding3D(e)
    x = synth.normal("pa.x", (b, s1, e))
    ctx = synth.normal("pa.ctx", (b, s2, e))
    xp = synth.uniform("pa.xp", (b, s1, 3), -1, 1)
    cp = synth.uniform("pa.cp", (b, s2, 3), -1, 1)
    sem = synth.normal("pa.sem", (b, s1, e), 0.5)
    t_emb = synth.normal("pa.t", (b, e))
    mask = torch.zeros(b, s1, dtype=torch.bool)
    mask[1, -3:] = True
    y, _ = mod(seq1=x, seq1_key_padding_mask=mask, seq2=ctx, seq2_key_padding_mask=None,
               seq1_pos=pe(xp), seq2_pos=pe(cp), seq1_sem_pos=sem, seq2_sem_pos=None, ada_sgnl=t_emb)
    save("parallel_attention", dict(out=y, check=synth.checksum(x, ctx, xp, cp, sem, t_emb)))


def build_planner(ref):
    torch.manual_seed(0)
    model = ref.DiffusionPlanner(**cases.PLANNER_KW).eval()
    cases.install_synth_trunk(model.prediction_head, cases.PLANNER_KW["embedding_dim"])
    synth.fill_state_dict(model.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    return model


@torch.no_grad()
def gen_diffusion_head(ref):
    model = build_planner(ref)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.7)
    # normalised-frame inputs straight into the head (diffusion_head.py:200)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    pcd_n = model.normalize_pos(inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    t = torch.tensor([37, 5])
    out = model.prediction_head(traj, inp["trajectory_mask"], t, visible_rgb=inp["rgb_obs"], visible_pcd=pcd_n,
                                curr_gripper=cur, goal_gripper=goal, instruction=inp["instruction"])
    save("diffusion_head", dict(out=out[-1], check=synth.checksum(traj, cur, goal, t)))


@torch.no_grad()
def gen_planner(ref):
    model = build_planner(ref)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    with synth.patched_randn(synth.NoiseStream("cd")):
        traj = model.compute_trajectory(inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"],
                                        inp["instruction"], inp["curr_gripper"], inp["goal_gripper"])
    save("planner_100step", dict(trajectory=traj, check=synth.checksum(inp["curr_gripper"], inp["goal_gripper"])))


@torch.no_grad()
def gen_act3d_variant(ref):
    """Act3D with the branches the two main fixtures do not take: 6-D rotation from the top ghost point, offset
    regression head, untied weights / ghost embeddings, two cameras, ragged ghost count."""
    kw = dict(cases.ACT3D_KW, use_instruction=True, rotation_parametrization="6D_from_top_ghost",
              regress_position_offset=True, weight_tying=False, gp_emb_tying=False, num_ghost_points_val=3 * 333)
    torch.manual_seed(0)
    model = ref.Act3D(**kw).eval()
    cases.install_synth_trunk(model, kw["embedding_dim"])
    synth.fill_state_dict(model.state_dict())
    inp = cases.act3d_inputs(batch=2, ncam=2, seed=3)
    sampler = synth.make_ghost_sampler(2, 333, seed=3)
    model._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor)
    out = model(inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"])
    save("act3d_variant", dict(
        position=out["position"], rotation=out["rotation"], gripper=out["gripper"],
        position_pyramid=out["position_pyramid"], ghost_pcd_masks_pyramid=out["ghost_pcd_masks_pyramid"],
        fine_ghost_pcd_offsets=out["fine_ghost_pcd_offsets"], query_features=out["query_features"],
        check=synth.checksum(inp["curr_gripper"], inp["instruction"][:, :2])))


def build_planner_ms(ref):
    torch.manual_seed(0)
    model = ref.DiffusionPlanner(**cases.PLANNER_MS_KW).eval()
    cases.install_synth_trunk(model.prediction_head, cases.PLANNER_MS_KW["embedding_dim"])
    synth.fill_state_dict(model.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    return model


@torch.no_grad()
def gen_planner_multiscale(ref):
    """feat_scales_to_use=3: one DiffusionHead.forward (all three refinements) and an 8-step sampling loop."""
    model = build_planner_ms(ref)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.4)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    pcd_n = model.normalize_pos(inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    t = torch.tensor([5, 2])
    outs = model.prediction_head(traj, inp["trajectory_mask"], t, visible_rgb=inp["rgb_obs"], visible_pcd=pcd_n,
                                 curr_gripper=cur, goal_gripper=goal, instruction=inp["instruction"])
    with synth.patched_randn(synth.NoiseStream("cdms")):
        sampled = model.compute_trajectory(inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"],
                                           inp["instruction"], inp["curr_gripper"], inp["goal_gripper"])
    save("planner_multiscale", dict(head_outs=[o.clone() for o in outs], trajectory=sampled,
                                    check=synth.checksum(traj, cur, goal, t, inp["curr_gripper"])))


def gen_planner_train(ref):
    """DiffusionPlanner.forward (training objective, diffusion_model.py:253-324) with autograd: loss value and
    parameter gradients.  eval() mode so that dropout is off; torch.randn / torch.randint are pinned."""
    model = build_planner(ref)
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    gt = cases.planner_gt_trajectory(batch=2, length=12)
    orig_randint = torch.randint
    torch.randint = lambda *a, **k: torch.tensor([37, 5])
    try:
        with synth.patched_randn(synth.NoiseStream("cdtr")):
            loss = model(gt, inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"], inp["instruction"],
                         inp["curr_gripper"], inp["goal_gripper"])
    finally:
        torch.randint = orig_randint
    loss.backward()
    # 3 M gradient entries: keep (norm, projection on a name-keyed random direction) per tensor, full tensors if small
    grads = {n: synth.grad_summary(n, p.grad) for n, p in model.named_parameters()
             if p.grad is not None and p.grad.abs().max() > 0}
    save("planner_train", dict(loss=loss.detach(), grads=grads, check=synth.checksum(gt, inp["curr_gripper"])))


def reference_loss_class():
    """LossAndMetrics straight from the reference's main_keypose.py source (the module itself cannot be
    imported here: tap / blosc / datasets are missing).  Only the class statement is executed."""
    import ast
    from oracle.ref_import import REFERENCE_ROOT
    src = open(os.path.join(REFERENCE_ROOT, "main_keypose.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "LossAndMetrics")
    ns = {"torch": torch, "F": torch.nn.functional, "np": __import__("numpy")}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "main_keypose.py", "exec"), ns)
    return ns["LossAndMetrics"]


def _loss_obj(cls, **kw):
    args = dict(position_loss="ce", rotation_parametrization="quat_from_query", ground_truth_gaussian_spread=0.01,
                compute_loss_at_all_layers=False, label_smoothing=0.0, position_loss_coeff=1.0,
                position_offset_loss_coeff=10000.0, rotation_loss_coeff=10.0, gripper_loss_coeff=1.0,
                symmetric_rotation_loss=False)
    args.update(kw)
    import inspect
    params = inspect.signature(cls.__init__).parameters
    return cls(**{k: v for k, v in args.items() if k in params})


def gen_keypose_loss(ref):
    cls = reference_loss_class()
    out = {}
    for vi, variant in enumerate(cases.LOSS_VARIANTS):
        pred, gt = cases.keypose_loss_case()
        leaves = [m.requires_grad_(True) for lvl in pred["ghost_pcd_masks_pyramid"] for m in lvl]
        pred["rotation"].requires_grad_(True)
        pred["gripper"].requires_grad_(True)
        pred["fine_ghost_pcd_offsets"].requires_grad_(True)
        losses = _loss_obj(cls, **variant).compute_loss(pred, {"action": gt})
        sum(losses.values()).backward()
        out[f"v{vi}"] = dict(losses={k: v.detach() for k, v in losses.items()},
                             dmasks=[m.grad if m.grad is not None else torch.zeros_like(m) for m in leaves],
                             drotation=pred["rotation"].grad, dgripper=pred["gripper"].grad,
                             doffsets=pred["fine_ghost_pcd_offsets"].grad, check=synth.checksum(gt, *[m.detach() for m in leaves]))
    save("keypose_loss", out)


def gen_act3d_train_grads(ref):
    """Reference Act3D with autograd + the reference's own loss: parameter gradients (pins the oracle's autograd)."""
    cls = reference_loss_class()
    kw = dict(cases.ACT3D_KW, use_instruction=True, num_ghost_points=3 * 96)
    torch.manual_seed(0)
    model = ref.Act3D(**kw).train()
    cases.install_synth_trunk(model, kw["embedding_dim"])
    synth.fill_state_dict(model.state_dict())
    inp = cases.act3d_inputs(batch=2, ncam=1)
    gt = cases.keypose_loss_case(batch=2)[1]
    sampler = synth.make_ghost_sampler(2, model.num_ghost_points, diameter=kw["fine_sampling_ball_diameter"])
    model._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor)
    out = model(inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"], gt_action=gt)
    losses = _loss_obj(cls).compute_loss(out, {"action": gt})
    sum(losses.values()).backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    save("act3d_train_grads", dict(grads=grads, losses={k: v.detach() for k, v in losses.items()},
                                   position_pyramid=[p.detach() for p in out["position_pyramid"]],
                                   check=synth.checksum(gt, inp["curr_gripper"])))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference()
    gen_rope(ref)
    gen_attention_stack(ref)
    gen_act3d(ref, False)
    gen_act3d(ref, True)
    gen_parallel_attention(ref)
    gen_diffusion_head(ref)
    gen_planner(ref)
    gen_keypose_loss(ref)
    gen_act3d_train_grads(ref)
    gen_planner_multiscale(ref)
    gen_planner_train(ref)
    gen_act3d_variant(ref)


if __name__ == "__main__":
    main()
