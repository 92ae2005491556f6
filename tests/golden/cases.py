"""Case builders shared by make_golden.py (drives the real reference) and the tests (drive the
oracle and the CUDA path).  A case builder returns plain tensors / dicts only; it never
imports the reference, the oracle or the product."""
import torch

from . import synth

ACT3D_KW = dict(
    backbone="resnet", image_size=(256, 256), embedding_dim=60, num_attn_heads=4,
    num_ghost_point_cross_attn_layers=2, num_query_cross_attn_layers=2, num_vis_ins_attn_layers=2,
    rotation_parametrization="quat_from_query", gripper_loc_bounds=synth.BOUNDS,
    num_ghost_points=768, num_ghost_points_val=768, weight_tying=True, gp_emb_tying=True,
    num_sampling_level=3, fine_sampling_ball_diameter=0.16, regress_position_offset=False,
)

PLANNER_KW = dict(
    backbone="resnet", image_size=(256, 256), embedding_dim=120, output_dim=7,
    num_vis_ins_attn_layers=2, num_query_cross_attn_layers=6, use_instruction=True, use_goal=True,
    use_goal_at_test=False, feat_scales_to_use=1, attn_rounds=1, weight_tying=True,
    gripper_loc_bounds=synth.BOUNDS, rotation_parametrization="6D", diffusion_timesteps=100,
)


# multi-scale refinement (feat_scales_to_use=3: coarse map, then the 64*L / 16*L fine points nearest the trajectory),
# untied weights so that every (round, scale) offset has its own parameters; short schedule for the sampling fixture
PLANNER_MS_KW = dict(PLANNER_KW, feat_scales_to_use=3, weight_tying=False, diffusion_timesteps=8)


def act3d_inputs(batch=2, ncam=1, seed=0):
    rgb, pcd = synth.rgbd("a3d", batch, ncam, 256, seed)
    return dict(
        visible_rgb=rgb, visible_pcd=pcd,
        instruction=synth.normal("a3d.instr", (batch, 53, 512), 1.0, seed),
        curr_gripper=synth.gripper_pose("a3d.grip", batch, seed, with_open=True),
    )


def planner_inputs(batch=2, ncam=1, length=12, seed=0, masked_tail=0):
    rgb, pcd = synth.rgbd("cd", batch, ncam, 256, seed)
    mask = torch.zeros(batch, length, dtype=torch.bool)
    if masked_tail:
        mask[-1, length - masked_tail:] = True
    return dict(
        trajectory_mask=mask, rgb_obs=rgb, pcd_obs=pcd,
        instruction=synth.normal("cd.instr", (batch, 53, 512), 1.0, seed),
        curr_gripper=synth.gripper_pose("cd.cur", batch, seed, with_open=False),
        goal_gripper=synth.gripper_pose("cd.goal", batch, seed, with_open=False),
    )


def install_synth_trunk(module, embed_dim, seed=0):
    """Replace backbone + normalize + FPN of a (reference or our) module by the deterministic
    SynthTrunk so that fixtures do not depend on ResNet-50's 25M random weights."""
    module.backbone = torch.nn.Identity()
    module.normalize = torch.nn.Identity()
    module.feature_pyramid = synth.SynthTrunk(embed_dim, seed)
    return module


def small_attention_case(e=60, heads=4, lq=40, lk=70, batch=2, seed=0):
    return dict(
        query=synth.normal("att.q", (lq, batch, e), 1.0, seed),
        context=synth.normal("att.c", (lk, batch, e), 1.0, seed),
        q_xyz=synth.points_in_bounds("att.qx", (batch, lq), seed),
        c_xyz=synth.points_in_bounds("att.cx", (batch, lk), seed),
        e=e, heads=heads,
    )


def keypose_loss_case(batch=3, ng=97, levels=3, layers=2, seed=0, with_offsets=True):
    """A synthetic Act3D output dict + ground truth for the keypose objective (main_keypose.py:353-429)."""
    gt_pos = synth.points_in_bounds("loss.gt", (batch,), seed)
    quat = synth.normal("loss.q", (batch, 4), seed=seed)
    gt = torch.cat([gt_pos, quat / quat.norm(dim=-1, keepdim=True), (synth.uniform("loss.o", (batch, 1), seed=seed) > 0.5).float()], -1)
    pred = dict(ghost_pcd_masks_pyramid=[], ghost_pcd_pyramid=[])
    for i in range(levels):
        # ghost points in a shrinking ball around the ground truth so that the Gaussian label is not one-hot
        off = synth.normal(f"loss.g{i}", (batch, ng, 3), 0.03 / (2 ** i), seed)
        pred["ghost_pcd_pyramid"].append((gt_pos[:, None] + off).transpose(1, 2))
        pred["ghost_pcd_masks_pyramid"].append([synth.normal(f"loss.m{i}{j}", (batch, ng), 2.0, seed) for j in range(layers)])
    rot = synth.normal("loss.r", (batch, 4), seed=seed)
    pred["rotation"] = rot / rot.norm(dim=-1, keepdim=True)
    pred["gripper"] = torch.sigmoid(synth.normal("loss.gr", (batch, 1), seed=seed))
    pred["position"] = synth.points_in_bounds("loss.p", (batch,), seed)
    pred["fine_ghost_pcd_offsets"] = synth.normal("loss.off", (batch, 3, ng), 0.002, seed) if with_offsets else None
    return pred, gt


LOSS_VARIANTS = [
    dict(),
    dict(compute_loss_at_all_layers=True, label_smoothing=0.1, symmetric_rotation_loss=True, ground_truth_gaussian_spread=0.02),
]


def planner_gt_trajectory(batch=2, length=12, seed=0):
    """Ground-truth trajectory (B, L, 7): xyz in the workspace + unit quaternion (main_trajectory.py:177-199 input)."""
    q = synth.normal("cd.gtq", (batch, length, 4), seed=seed)
    return torch.cat([synth.points_in_bounds("cd.gt", (batch, length), seed), q / q.norm(dim=-1, keepdim=True)], -1)
