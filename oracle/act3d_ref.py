"""Oracle: Act3D keypose forward (test infrastructure only, see oracle/__init__.py).

Functional CPU restatement of model/keypose_optimization/act3d.py:176-535 of the
reference, driven by a ``state_dict`` (same keys as the reference module) and a
small config dict.  The frozen backbone + FPN ("trunk") are third-party
torchvision code on both sides and are passed in as a callable.
"""
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .attention import relative_cross_attn_stack
from .geometry import (local_topk, normalise_quat, ortho6d_to_matrix, pcd_level,
                       sample_ball, sample_cube)
from .rope import rope3d_table


@dataclass
class Act3DConfig:
    embedding_dim: int = 60
    num_attn_heads: int = 4
    num_ghost_point_cross_attn_layers: int = 2
    num_query_cross_attn_layers: int = 2
    num_vis_ins_attn_layers: int = 2
    rotation_parametrization: str = "quat_from_query"
    gripper_loc_bounds: object = None
    num_sampling_level: int = 3
    fine_sampling_ball_diameter: float = 0.16
    regress_position_offset: bool = False
    use_instruction: bool = False
    ghost_points_per_level: int = 1000          # already divided by num_sampling_level
    feature_map_pyramid: List[str] = field(default_factory=lambda: ["res3", "res1", "res1", "res1"])
    downscaling_factor_pyramid: List[int] = field(default_factory=lambda: [8, 2, 2, 2])


def trunk_from_module(model) -> Callable:
    """normalize -> backbone -> FPN of a reference/our module (act3d.py:364-369)."""
    def run(rgb_flat):
        return model.feature_pyramid(model.backbone(model.normalize(rgb_flat)))
    return run


def default_ghost_sampler(cfg: Act3DConfig, batch: int):
    """Host sampler with the reference's numpy call order (act3d.py:394-440)."""
    bounds_all = np.array(cfg.gripper_loc_bounds)
    diam = [None, cfg.fine_sampling_ball_diameter, cfg.fine_sampling_ball_diameter / 4.0,
            cfg.fine_sampling_ball_diameter / 16.0]

    def sample(level, anchor):
        n = cfg.ghost_points_per_level
        if level == 0:
            pts = np.stack([sample_cube(bounds_all, n) for _ in range(batch)])
        else:
            a = anchor[:, 0].cpu().numpy()
            r = diam[level] / 2
            lo = np.clip(a - r, bounds_all[0], bounds_all[1])
            hi = np.clip(a + r, bounds_all[0], bounds_all[1])
            pts = np.stack([sample_ball(a[i], r, np.stack([lo[i], hi[i]]), n) for i in range(batch)])
        return torch.from_numpy(pts).float()
    return sample


def act3d_forward(sd, cfg: Act3DConfig, trunk: Callable, visible_rgb, visible_pcd, instruction,
                  curr_gripper, gt_action=None, ghost_sampler: Optional[Callable] = None,
                  teacher_positions: Optional[list] = None):
    """Returns the same dict as Act3D.forward (act3d.py:340-357).

    ``teacher_positions`` (list of (B,1,3), one per level) replaces the argmax-picked
    position fed to the *next* level (anchor, top-k centre, query rotary) -- used by
    parity tests to decouple levels; outputs still report the free argmax.
    """
    e, h = cfg.embedding_dim, cfg.num_attn_heads
    bsz, ncam = visible_rgb.shape[:2]
    dev = visible_rgb.device
    gt_pos = gt_action[:, :3].unsqueeze(1).detach() if gt_action is not None else None
    grip_xyz = curr_gripper[:, :3]
    if ghost_sampler is None:
        ghost_sampler = default_ghost_sampler(cfg, bsz)

    # ---- visual trunk + point pyramid + rotary tables (act3d.py:359-392)
    rgb_flat = visible_rgb.reshape(bsz * ncam, *visible_rgb.shape[2:])
    pcd_flat = visible_pcd.reshape(bsz * ncam, *visible_pcd.shape[2:])
    fpn = trunk(rgb_flat)
    feats_pyr, rope_pyr, pcd_pyr = [], [], []
    for i in range(cfg.num_sampling_level):
        fm = fpn[cfg.feature_map_pyramid[i]]
        pts = pcd_level(pcd_flat, cfg.downscaling_factor_pyramid[i], ncam)
        feats_pyr.append(fm.view(bsz, ncam, *fm.shape[1:]))
        rope_pyr.append(rope3d_table(pts, e))
        pcd_pyr.append(pts)

    # ---- instruction tokens (act3d.py:198-216; ins_pos_emb is off in shipped configs)
    if cfg.use_instruction:
        instr_feats = F.linear(instruction, sd["instruction_encoder.weight"], sd["instruction_encoder.bias"])
        instr_feats = instr_feats.transpose(0, 1)                           # (53, B, E)
        instr_rope = rope3d_table(torch.zeros(bsz, instr_feats.shape[0], 3, device=dev), e)
    else:
        instr_feats, instr_rope = None, None

    # ---- gripper token (act3d.py:219-220)
    grip_rope = rope3d_table(grip_xyz.unsqueeze(1), e)
    grip_feat = sd["curr_gripper_embed.weight"].repeat(bsz, 1).unsqueeze(0)  # (1, B, E)

    out_ghost_pcd, out_ghost_feats, out_pos, out_masks = [], [], [], []
    carried = []                                                             # positions fed forward
    query = None
    for i in range(cfg.num_sampling_level):
        anchor = None if i == 0 else (gt_pos if gt_pos is not None else carried[-1])
        ghost = ghost_sampler(i, anchor).to(dev)                             # (B, Ng, 3)

        if i == 0:
            ctx = feats_pyr[i].permute(1, 3, 4, 0, 2).reshape(-1, bsz, e)    # (ncam h w, B, E)
            ctx_rope = rope_pyr[i]
        else:
            idx = local_topk(carried[-1], pcd_pyr[i], 32 * 32 * ncam)        # act3d.py:244-245
            flat = feats_pyr[i].permute(0, 1, 3, 4, 2).reshape(bsz, -1, e)   # (B, ncam h w, E)
            ctx = torch.stack([f[j] for f, j in zip(flat, idx)]).transpose(0, 1)
            ctx_rope = torch.stack([f[j] for f, j in zip(rope_pyr[i], idx)])

        ctx = torch.cat([ctx, grip_feat], dim=0)                             # act3d.py:258-260
        ctx_rope = torch.cat([ctx_rope, grip_rope], dim=1)
        if cfg.use_instruction:                                              # act3d.py:261-270
            ctx = relative_cross_attn_stack(sd, f"vis_ins_attn_pyramid.{i}.", h, cfg.num_vis_ins_attn_layers,
                                            ctx, instr_feats)[-1]
            ctx = torch.cat([ctx, instr_feats], dim=0)
            ctx_rope = torch.cat([ctx_rope, instr_rope], dim=1)

        # ghost points cross-attend to the context (act3d.py:442-465)
        ghost_rope = rope3d_table(ghost, e)
        g0 = sd[f"ghost_points_embed_pyramid.{i}.weight"].unsqueeze(0).repeat(ghost.shape[1], bsz, 1)
        ghost_feats = relative_cross_attn_stack(sd, f"ghost_point_cross_attn_pyramid.{i}.", h,
                                                cfg.num_ghost_point_cross_attn_layers, g0, ctx,
                                                ghost_rope, ctx_rope)[-1]

        # query token (act3d.py:281-301): no rotary at level 0
        if i == 0:
            query = sd["query_embed.weight"].unsqueeze(1).repeat(1, bsz, 1)
            q_rope, c_rope = None, None
        else:
            q_rope, c_rope = rope3d_table(carried[-1], e), ctx_rope
        q_layers = relative_cross_attn_stack(sd, f"query_cross_attn_pyramid.{i}.", h,
                                             cfg.num_query_cross_attn_layers, query, ctx, q_rope, c_rope)

        # mask logits + top ghost point (act3d.py:493-494, 312-314)
        masks = [torch.einsum("bc,nbc->bn", f.squeeze(0), ghost_feats) for f in q_layers]
        query = q_layers[-1]
        top = torch.max(masks[-1], dim=-1).indices
        ghost_cn = ghost.transpose(1, 2)                                     # (B, 3, Ng)
        pos_i = ghost_cn[torch.arange(bsz), :, top].unsqueeze(1)

        out_ghost_pcd.append(ghost_cn)
        out_ghost_feats.append(ghost_feats)
        out_pos.append(pos_i)
        out_masks.append(masks)
        carried.append(teacher_positions[i] if teacher_positions is not None else pos_i)

    # ---- offsets + action head (act3d.py:323-337, 507-535)
    offsets = None
    if cfg.regress_position_offset:
        o = F.linear(F.relu(F.linear(out_ghost_feats[-1], sd["ghost_point_offset_predictor.0.weight"],
                                     sd["ghost_point_offset_predictor.0.bias"])),
                     sd["ghost_point_offset_predictor.2.weight"], sd["ghost_point_offset_predictor.2.bias"])
        offsets = o.permute(1, 2, 0)                                         # (B, 3, Ng)

    top = torch.max(out_masks[-1][-1], dim=-1).indices
    position = out_ghost_pcd[-1][torch.arange(bsz), :, top]
    if offsets is not None:
        position = position + offsets[torch.arange(bsz), :, top]
    if cfg.rotation_parametrization.endswith("from_top_ghost"):
        feats = out_ghost_feats[-1].transpose(0, 1)[torch.arange(bsz), top]
    else:
        feats = query.squeeze(0)
    pred = F.linear(F.relu(F.linear(feats, sd["gripper_state_predictor.0.weight"], sd["gripper_state_predictor.0.bias"])),
                    sd["gripper_state_predictor.2.weight"], sd["gripper_state_predictor.2.bias"])
    rot_dim = 4 if "quat" in cfg.rotation_parametrization else 6
    rotation = normalise_quat(pred[:, :rot_dim]) if rot_dim == 4 else ortho6d_to_matrix(pred[:, :rot_dim])
    gripper = torch.sigmoid(pred[:, rot_dim:])

    return {
        "position": position, "rotation": rotation, "gripper": gripper,
        "position_pyramid": out_pos,
        "visible_rgb_mask_pyramid": [None] * cfg.num_sampling_level,
        "ghost_pcd_masks_pyramid": out_masks,
        "ghost_pcd_pyramid": out_ghost_pcd,
        "fine_ghost_pcd_offsets": offsets,
        "visible_rgb_features_pyramid": feats_pyr,
        "visible_pcd_pyramid": pcd_pyr,
        "query_features": query,
        "instruction_features": instr_feats,
        "instruction_dummy_pos": instr_rope,
        "ghost_pcd_features_pyramid": out_ghost_feats,      # extra (not in the reference dict): for parity tests
    }
