"""ChainedDiffuser trajectory planner on the B200 kernels -- drop-in for the reference's
model/trajectory_optimization/diffusion_model.py (+ diffusion_head.py, model/utils/encoder.py):
same ctor kwargs, forward / compute_trajectory signatures and state_dict keys.
"""
import torch
from torch import nn
from torchvision.ops import FeaturePyramidNetwork

from .params import ParallelStackParams, mlp
from .trunk import build_backbone


def _repeat(n, tie, factory):
    if tie:
        one = factory()
        return nn.ModuleList(one for _ in range(n))
    return nn.ModuleList(factory() for _ in range(n))


class DiffusionHead(nn.Module):
    """Parameter owner with the reference's key names (diffusion_head.py:12-198, encoder.py:14-79)."""

    def __init__(self, backbone="clip", image_size=(256, 256), embedding_dim=60, output_dim=7, num_attn_heads=8,
                 num_vis_ins_attn_layers=2, num_query_cross_attn_layers=6, use_instruction=False, use_goal=False,
                 use_sigma=False, feat_scales_to_use=1, attn_rounds=1, weight_tying=False,
                 rotation_parametrization="quat"):
        super().__init__()
        if tuple(image_size) != (256, 256):
            raise AssertionError("image_size must be (256, 256)")
        if rotation_parametrization != "6D":
            # traj_encoder is Linear(9, E) in the reference: only '6D' can run
            raise NotImplementedError("only rotation_parametrization='6D' is runnable (as in the reference)")
        if feat_scales_to_use != 1 or attn_rounds != 1:
            raise NotImplementedError("feat_scales_to_use > 1 / attn_rounds > 1 are not built yet (DESIGN.md 'next')")
        if use_sigma:
            raise NotImplementedError("use_sigma is never enabled by the reference's entry points")
        self.image_size = tuple(image_size)
        self.embedding_dim, self.num_attn_heads = embedding_dim, num_attn_heads
        self.use_instruction, self.use_goal = use_instruction, use_goal
        self.attn_rounds, self.feat_scales = attn_rounds, feat_scales_to_use
        self.rotation_parametrization = rotation_parametrization
        output_dim += 2
        e, h = embedding_dim, num_attn_heads
        n = attn_rounds * feat_scales_to_use

        self.backbone, self.normalize = build_backbone(backbone)
        for p in self.backbone.parameters():
            p.requires_grad = False
        self.feature_pyramid = FeaturePyramidNetwork([64, 256, 512, 1024, 2048], e)
        self.feature_map_pyramid = ["res3", "res1", "res1", "res1"]
        self.downscaling_factor_pyramid = [8, 2, 2, 2]
        self.curr_gripper_embed = nn.Embedding(1, e)
        self.goal_gripper_embed = nn.Embedding(1, e)
        self.instruction_encoder = nn.Linear(512, e)

        self.traj_encoder = mlp(9, e, e, dropout=0.1)
        self.curr_gripper_encoder = nn.Linear(output_dim, e)
        if use_goal:
            self.goal_gripper_encoder = nn.Linear(output_dim, e)
        if use_instruction:
            self.vl_attention = _repeat(n, weight_tying, lambda: ParallelStackParams(num_vis_ins_attn_layers, e, h))
        self.traj_lang_attention = _repeat(n, weight_tying, lambda: ParallelStackParams(1, e, h, apply_ffn=False))
        adaln = dict(self_attention=True, rotary_pe=True, use_adaln=True)
        self.traj_attention = _repeat(n, weight_tying,
                                      lambda: ParallelStackParams(num_query_cross_attn_layers - 2, e, h, **adaln))
        self.pos_attention = _repeat(n, weight_tying, lambda: ParallelStackParams(2, e, h, **adaln))
        self.rot_attention = _repeat(n, weight_tying, lambda: ParallelStackParams(2, e, h, **adaln))
        self.pos_regressor = nn.ModuleList(mlp(e, e, 3, dropout=0.1) for _ in range(n))
        self.rot_regressor = nn.ModuleList(mlp(e, e, output_dim - 3, dropout=0.1) for _ in range(n))


class DiffusionPlanner(nn.Module):

    def __init__(self, backbone="clip", image_size=(256, 256), embedding_dim=60, output_dim=7,
                 num_vis_ins_attn_layers=2, num_query_cross_attn_layers=8, use_instruction=False, use_goal=False,
                 use_goal_at_test=True, feat_scales_to_use=1, attn_rounds=1, weight_tying=False,
                 gripper_loc_bounds=None, rotation_parametrization="quat", diffusion_timesteps=100):
        super().__init__()
        self._use_goal = use_goal
        self._use_goal_at_test = use_goal_at_test
        self._rotation_parametrization = rotation_parametrization
        self.prediction_head = DiffusionHead(
            backbone=backbone, image_size=image_size, embedding_dim=embedding_dim, output_dim=output_dim,
            num_vis_ins_attn_layers=num_vis_ins_attn_layers, num_query_cross_attn_layers=num_query_cross_attn_layers,
            use_instruction=use_instruction, use_goal=use_goal, feat_scales_to_use=feat_scales_to_use,
            attn_rounds=attn_rounds, weight_tying=weight_tying, rotation_parametrization=rotation_parametrization)
        self.n_steps = diffusion_timesteps
        self.gripper_loc_bounds = torch.tensor(gripper_loc_bounds)

    def compute_trajectory(self, trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper):
        raise NotImplementedError("trajectory kernels land in the next commit")

    def forward(self, gt_trajectory, trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper,
                run_inference=False):
        if run_inference:
            return self.compute_trajectory(trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper)
        raise NotImplementedError("training forward (one denoiser call + L1 loss) lands with the backward kernels")
