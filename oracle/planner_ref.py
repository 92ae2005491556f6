"""Oracle: ChainedDiffuser trajectory denoiser + sampling loop (test infrastructure only).

Functional CPU restatement of
  model/trajectory_optimization/diffusion_head.py:200-363   (DiffusionHead.forward / _one_attention_round)
  model/utils/encoder.py:81-203                              (token encoders)
  model/trajectory_optimization/diffusion_model.py:64-324    (DiffusionPlanner)
for rotation_parametrization='6D' (the only runnable one), any attn_rounds, and
feat_scales_to_use in {1, 2, 3} (coarse-to-fine local refinement around the trajectory, find_traj_nn).  ``sd`` holds the keys under ``prediction_head.`` with that
prefix stripped.
"""
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .attention import parallel_attention_stack
from .ddpm import DDPMScheduler
from .geometry import (find_traj_nn, matrix_to_ortho6d, matrix_to_quat, normalise_quat, ortho6d_to_matrix,
                       pcd_level, quat_to_matrix)
from .rope import rope3d_table, sinusoidal_embedding


@dataclass
class PlannerConfig:
    embedding_dim: int = 120
    num_attn_heads: int = 8
    num_vis_ins_attn_layers: int = 2
    num_query_cross_attn_layers: int = 6
    use_instruction: bool = True
    use_goal: bool = True
    use_goal_at_test: bool = False
    gripper_loc_bounds: object = None
    diffusion_timesteps: int = 100
    feat_scales_to_use: int = 1
    attn_rounds: int = 1


def _mlp2(sd, p, x, i0="0", i1="3"):
    return F.linear(F.relu(F.linear(x, sd[f"{p}{i0}.weight"], sd[f"{p}{i0}.bias"])),
                    sd[f"{p}{i1}.weight"], sd[f"{p}{i1}.bias"])


FEATURE_MAP_PYRAMID = ["res3", "res1", "res1", "res1"]        # encoder.py:52-53
DOWNSCALING_PYRAMID = [8, 2, 2, 2]


def encode_context(sd, cfg: PlannerConfig, trunk: Callable, rgb, pcd, instruction, curr_gripper, goal_gripper):
    """Everything in DiffusionHead.forward that does not depend on the trajectory or the
    timestep: trunk features and point pyramid of every scale, instruction tokens,
    gripper / goal tokens, and -- for the scale-0 offsets, whose context is the whole coarse map --
    the context after the vision->language attention."""
    e, h = cfg.embedding_dim, cfg.num_attn_heads
    b, ncam = rgb.shape[:2]
    fpn = trunk(rgb.reshape(b * ncam, *rgb.shape[2:]))                       # encoder.py:133-139
    pcd_flat = pcd.reshape(b * ncam, *pcd.shape[2:])
    feats, pts = [], []
    for s in range(cfg.feat_scales_to_use):                                  # encoder.py:141-167
        fm = fpn[FEATURE_MAP_PYRAMID[s]]
        feats.append(fm.view(b, ncam, *fm.shape[1:]).permute(0, 1, 3, 4, 2).reshape(b, -1, e))   # diffusion_head.py:290-293
        pts.append(pcd_level(pcd_flat, DOWNSCALING_PYRAMID[s], ncam))

    instr, instr_rope = None, None
    if cfg.use_instruction:                                                  # encoder.py:169-187
        instr = F.linear(instruction, sd["instruction_encoder.weight"], sd["instruction_encoder.bias"])
        instr_rope = rope3d_table(torch.zeros(b, instr.shape[1], 3), e)

    cur = F.linear(curr_gripper, sd["curr_gripper_encoder.weight"], sd["curr_gripper_encoder.bias"])[:, None]
    cur = cur + sd["curr_gripper_embed.weight"].repeat(b, 1).unsqueeze(1)    # diffusion_head.py:231-237
    goal = None
    if cfg.use_goal:                                                         # diffusion_head.py:239-247
        goal = F.linear(goal_gripper, sd["goal_gripper_encoder.weight"], sd["goal_gripper_encoder.bias"])[:, None]
        goal = goal + sd["goal_gripper_embed.weight"].repeat(b, 1).unsqueeze(1)
    out = {"feats": feats, "pts": pts, "instr": instr, "instr_rope": instr_rope, "cur": cur, "goal": goal,
           "cur_xyz": curr_gripper[:, :3], "goal_xyz": goal_gripper[:, :3], "pcd": pts[0], "static": {}}
    for off in range(cfg.attn_rounds * cfg.feat_scales_to_use):
        if off % cfg.feat_scales_to_use == 0:
            out["static"][off] = _offset_context(sd, cfg, out, off, None)
    out["ctx"], out["ctx_rope"] = out["static"][0]
    return out


def _offset_context(sd, cfg: PlannerConfig, context, off, p_inds):
    """Context tokens + rotary table of one (round, scale) offset: visual tokens (all of them, or the
    p_inds subset), vision->language attention, current / goal gripper tokens.  diffusion_head.py:289-323."""
    e, h = cfg.embedding_dim, cfg.num_attn_heads
    scale = off % cfg.feat_scales_to_use
    feats, pts = context["feats"][scale], context["pts"][scale]
    if p_inds is not None:                                                   # diffusion_head.py:295-302
        feats = torch.stack([f[i] for f, i in zip(feats, p_inds)])
        pts = torch.stack([f[i] for f, i in zip(pts, p_inds)])
    rope = rope3d_table(pts, e)
    if cfg.use_instruction:                                                  # diffusion_head.py:305-314
        feats = parallel_attention_stack(sd, f"vl_attention.{off}.", h, cfg.num_vis_ins_attn_layers,
                                         feats, None, context["instr"])
    feats = torch.cat([feats, context["cur"]], dim=1)
    rope = torch.cat([rope, rope3d_table(context["cur_xyz"][:, None], e)], dim=1)
    if cfg.use_goal:                                                         # diffusion_head.py:320-323
        feats = torch.cat([feats, context["goal"]], dim=1)
        rope = torch.cat([rope, rope3d_table(context["goal_xyz"][:, None], e)], dim=1)
    return feats, rope


def denoise_all(sd, cfg: PlannerConfig, context, trajectory, trajectory_mask, timestep):
    """One DiffusionHead.forward given the step-invariant context: the list of refined trajectories, one per
    (attention round, feature scale) offset.  diffusion_head.py:215-277 (token prep, loop) and :325-363."""
    e, h = cfg.embedding_dim, cfg.num_attn_heads
    b, length, _ = trajectory.shape
    x0 = _mlp2(sd, "traj_encoder.", trajectory)                              # computed once from the input trajectory
    traj_rope = rope3d_table(trajectory[..., :3], e)
    t_emb = sinusoidal_embedding(timestep, e)                                # encoder.py:199
    wp_pe = sinusoidal_embedding(torch.arange(0, length), e)[None].repeat(b, 1, 1)   # diffusion_head.py:326-328
    outs = []
    for off in range(cfg.attn_rounds * cfg.feat_scales_to_use):
        scale = off % cfg.feat_scales_to_use
        if cfg.use_goal and scale > 0:                                       # diffusion_head.py:253-259
            p_inds = find_traj_nn(outs[-1][..., :3], context["pts"][scale], 64 if scale == 1 else 16)
            ctx, ctx_rope = _offset_context(sd, cfg, context, off, p_inds)
        elif off in context["static"]:
            ctx, ctx_rope = context["static"][off]
        else:
            ctx, ctx_rope = _offset_context(sd, cfg, context, off, None)
        x = x0
        if cfg.use_instruction:                                              # diffusion_head.py:330-336
            x = parallel_attention_stack(sd, f"traj_lang_attention.{off}.", h, 1, x, trajectory_mask, context["instr"],
                                         seq1_sem_pos=wp_pe, apply_ffn=False)
        common = dict(seq1_mask=trajectory_mask, seq2=ctx, seq1_rope=traj_rope, seq2_rope=ctx_rope,
                      seq1_sem_pos=wp_pe, ada_signal=t_emb, self_attention=True, rotary=True, use_adaln=True)
        x = parallel_attention_stack(sd, f"traj_attention.{off}.", h, cfg.num_query_cross_attn_layers - 2, x, **common)
        pos_f = parallel_attention_stack(sd, f"pos_attention.{off}.", h, 2, x, **common)
        rot_f = parallel_attention_stack(sd, f"rot_attention.{off}.", h, 2, x, **common)
        upd = torch.cat((_mlp2(sd, f"pos_regressor.{off}.", pos_f), _mlp2(sd, f"rot_regressor.{off}.", rot_f)), -1)
        base = trajectory if not outs else outs[-1]
        outs.append(torch.cat((base[..., :3] + upd[..., :3], upd[..., 3:]), -1))   # diffusion_head.py:271-274
    return outs


def denoise_once(sd, cfg: PlannerConfig, context, trajectory, trajectory_mask, timestep):
    """Last refinement of denoise_all (what the sampling loop keeps, diffusion_model.py:104)."""
    return denoise_all(sd, cfg, context, trajectory, trajectory_mask, timestep)[-1]


# ---------------------------------------------------------------- planner wrapper

def normalize_pos(cfg, pos):
    lo = torch.tensor(cfg.gripper_loc_bounds[0]).float()
    hi = torch.tensor(cfg.gripper_loc_bounds[1]).float()
    return (pos - lo) / (hi - lo) * 2.0 - 1.0                               # diffusion_model.py:187-190


def unnormalize_pos(cfg, pos):
    lo = torch.tensor(cfg.gripper_loc_bounds[0]).float()
    hi = torch.tensor(cfg.gripper_loc_bounds[1]).float()
    return (pos + 1.0) / 2.0 * (hi - lo) + lo                               # diffusion_model.py:192-195


def quat_signal_to_6d(signal):
    """(..., 7+) [xyz, quat(taken as real-first), rest] -> (..., 9+).  diffusion_model.py:197-213."""
    signal = signal.clone()
    signal[..., 3:7] = normalise_quat(signal[..., 3:7])
    rot = quat_to_matrix(signal[..., 3:7])
    lead = rot.shape[:-2]
    r6 = matrix_to_ortho6d(rot.reshape(-1, 3, 3)).reshape(*lead, 6)
    return torch.cat([signal[..., :3], r6, signal[..., 7:]], dim=-1)


def sixd_signal_to_quat(signal):
    """(B, L, 9+) -> (B, L, 7+).  diffusion_model.py:215-230."""
    b, length, _ = signal.shape
    quat = matrix_to_quat(ortho6d_to_matrix(signal[..., 3:9].reshape(b * length, 6))).reshape(b, length, 4)
    return torch.cat([signal[..., :3], quat, signal[..., 9:]], dim=-1)


def compute_trajectory(sd, cfg: PlannerConfig, trunk, trajectory_mask, rgb, pcd, instruction,
                       curr_gripper, goal_gripper, noise_fn: Optional[Callable] = None,
                       hoist_context: bool = True, n_steps: Optional[int] = None):
    """DiffusionPlanner.compute_trajectory (diffusion_model.py:121-185) + conditional_sample (:86-119).

    ``noise_fn(shape)`` supplies the Gaussian draws in the reference's call order
    ((B,L,9) once, then per step (B,L,3) and (B,L,6)); default torch.randn.
    ``hoist_context=False`` re-encodes the context every step like the reference does
    (bit-identical results) -- used when timing the CPU baseline.
    """
    if noise_fn is None:
        noise_fn = lambda shape: torch.randn(shape)
    steps = n_steps or cfg.diffusion_timesteps
    pos_s = DDPMScheduler(cfg.diffusion_timesteps, "scaled_linear", "sample")
    rot_s = DDPMScheduler(cfg.diffusion_timesteps, "squaredcos_cap_v2", "sample")

    pcd_n = normalize_pos(cfg, pcd.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    cur = curr_gripper.clone()
    goal = goal_gripper.clone()
    cur[:, :3] = normalize_pos(cfg, cur[:, :3])
    goal[:, :3] = normalize_pos(cfg, goal[:, :3])
    cur = quat_signal_to_6d(cur)
    goal = quat_signal_to_6d(goal)

    b, d = cur.shape
    length = trajectory_mask.size(1)
    cond = torch.zeros(b, length, d)
    cmask = torch.zeros_like(cond)
    cond[:, 0] = cur
    cmask[:, 0] = 1
    if cfg.use_goal_at_test:                                                 # diffusion_model.py:162-167
        for i in range(b):
            neg = -int(trajectory_mask[i].sum())
            cond[i][neg - 1] = goal[i]
            cmask[i][neg - 1:] = 1
    cmask = cmask.bool()

    pos_s.set_timesteps(steps)
    rot_s.set_timesteps(steps)
    traj = noise_fn(cond.shape) + cond                                       # diffusion_model.py:91-96
    ctx = encode_context(sd, cfg, trunk, rgb, pcd_n, instruction, cur, goal) if hoist_context else None
    last = pos_s.timesteps[-1]
    for t in pos_s.timesteps:
        if not hoist_context:
            ctx = encode_context(sd, cfg, trunk, rgb, pcd_n, instruction, cur, goal)
        out = denoise_once(sd, cfg, ctx, traj, trajectory_mask, t * torch.ones(b).long()).clone()
        out[cmask] = cond[cmask]
        if t == last:
            traj = out
        else:
            c0, c1, sg = pos_s.step_coefficients(int(t))
            p = c0 * out[..., :3].clamp(-1, 1) + c1 * traj[..., :3]
            if int(t) > 0:
                p = p + sg * noise_fn(p.shape)
            c0, c1, sg = rot_s.step_coefficients(int(t))
            r = c0 * out[..., 3:9].clamp(-1, 1) + c1 * traj[..., 3:9]
            if int(t) > 0:
                r = r + sg * noise_fn(r.shape)
            traj = torch.cat((p, r), -1)

    traj = sixd_signal_to_quat(traj)
    traj[:, :, :3] = unnormalize_pos(cfg, traj[:, :, :3])
    return traj


def training_loss(sd, cfg: PlannerConfig, trunk, gt_trajectory, trajectory_mask, rgb, pcd, instruction,
                  curr_gripper, goal_gripper, noise_fn: Callable, timesteps: torch.Tensor):
    """DiffusionPlanner.forward with run_inference=False (diffusion_model.py:253-324): normalise, 6-D rotations,
    one random timestep per sample, noise added by the two schedulers, one denoiser evaluation, and
    100 * L1(position) + 10 * L1(rotation) summed over the returned refinements.  ``noise_fn(shape)`` and
    ``timesteps`` stand for the reference's torch.randn / torch.randint draws."""
    pos_s = DDPMScheduler(cfg.diffusion_timesteps, "scaled_linear", "sample")
    rot_s = DDPMScheduler(cfg.diffusion_timesteps, "squaredcos_cap_v2", "sample")
    gt = gt_trajectory.clone()
    gt[:, :, :3] = normalize_pos(cfg, gt[:, :, :3])
    pcd_n = normalize_pos(cfg, pcd.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    cur, goal = curr_gripper.clone(), goal_gripper.clone()
    cur[:, :3] = normalize_pos(cfg, cur[:, :3])
    goal[:, :3] = normalize_pos(cfg, goal[:, :3])
    gt, cur, goal = quat_signal_to_6d(gt), quat_signal_to_6d(cur), quat_signal_to_6d(goal)
    noise = noise_fn(gt.shape)
    noisy = torch.cat((pos_s.add_noise(gt[..., :3], noise[..., :3], timesteps),
                       rot_s.add_noise(gt[..., 3:9], noise[..., 3:9], timesteps)), -1)
    ctx = encode_context(sd, cfg, trunk, rgb, pcd_n, instruction, cur, goal)
    total = 0
    for pred in denoise_all(sd, cfg, ctx, noisy, trajectory_mask, timesteps):
        total = total + (100 * F.l1_loss(pred[..., :3], gt[..., :3], reduction="mean")
                         + 10 * F.l1_loss(pred[..., 3:9], gt[..., 3:9], reduction="mean"))
    return total
