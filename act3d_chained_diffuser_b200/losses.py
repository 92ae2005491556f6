"""Training objectives of the two entry points, on the device.

keypose_loss mirrors LossAndMetrics.compute_loss of the reference's keypose trainer
(main_keypose.py:353-429, position_loss="ce"): a soft cross-entropy between the mask logits of every
pyramid level and a Gaussian label built from the ghost-point coordinates (one kernel per level,
csrc/a3d_train.cu soft_ce_kernel, forward and gradient in the same pass), quaternion / 6D MSE and
gripper MSE.  The planner's objective lives in DiffusionPlanner.forward (diffusion_model.py:253-324).
"""
import torch
import torch.nn.functional as F

from . import lib


class _SoftCE(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, ghost, gt, spread, label_smoothing):
        loss, dlog = lib.soft_ce(logits.contiguous(), ghost, gt, spread, label_smoothing,
                                 want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(dlog if dlog is not None else logits.new_empty(0))
        return loss

    @staticmethod
    def backward(ctx, grad):
        (dlog,) = ctx.saved_tensors
        return dlog * grad[:, None], None, None, None, None


def soft_cross_entropy(logits, ghost_bn3, gt_position, spread, label_smoothing=0.0):
    """(B, Ng) logits, (B, Ng, 3) ghost points, (B, 3) ground truth -> (B,) losses."""
    return _SoftCE.apply(logits, ghost_bn3.detach().contiguous().float(), gt_position.detach().contiguous().float(),
                         float(spread), float(label_smoothing))


def keypose_loss(pred, gt_action, position_loss_coeff=1.0, position_offset_loss_coeff=10000.0, rotation_loss_coeff=10.0,
                 gripper_loss_coeff=1.0, ground_truth_gaussian_spread=0.01, label_smoothing=0.0,
                 compute_loss_at_all_layers=False, symmetric_rotation_loss=False):
    """dict of scalar losses with the reference's key names; sum(losses.values()) is what the trainer
    back-propagates (main_keypose.py:223-226).  gt_action (B, 8) = xyz, quaternion, gripper open."""
    gt_action = gt_action.to(pred["rotation"].device).float()
    gt_pos = gt_action[:, :3]
    losses = {}
    levels = pred["ghost_pcd_masks_pyramid"]
    layer_ids = range(len(levels[0])) if compute_loss_at_all_layers else [-1]
    for j in layer_ids:                        # later layers overwrite earlier ones, exactly like the reference's dict
        for i, masks in enumerate(levels):
            ghost = pred["ghost_pcd_pyramid"][i].transpose(1, 2)                       # (B, Ng, 3)
            ce = soft_cross_entropy(masks[j], ghost, gt_pos, ground_truth_gaussian_spread, label_smoothing)
            losses[f"position_ce_level{i}"] = ce.mean() * position_loss_coeff / len(levels)
    if pred.get("fine_ghost_pcd_offsets") is not None:                                 # main_keypose.py:405-417
        with_off = pred["ghost_pcd_pyramid"][-1] + pred["fine_ghost_pcd_offsets"]
        if pred["ghost_pcd_pyramid"][-1].shape[-1] != pred["ghost_pcd_pyramid"][0].shape[-1]:
            with_off = with_off[:, :, -(pred["ghost_pcd_pyramid"][-1].shape[-1] // len(levels)):]
        losses["position_offset"] = (F.mse_loss(with_off, gt_pos.unsqueeze(-1).expand_as(with_off))
                                     * (position_offset_loss_coeff * position_loss_coeff))
    pred["position"] = pred["position"].detach()                                       # main_keypose.py:419-424
    rot = pred["rotation"]
    if rot.dim() == 2:                                                                 # quaternion heads
        gt_quat = gt_action[:, 3:7]
        if symmetric_rotation_loss:
            a = F.mse_loss(rot, gt_quat, reduction="none").mean(1)
            b = F.mse_loss(rot, -gt_quat, reduction="none").mean(1)
            sel = (a < b).float()
            losses["rotation"] = (sel * a + (1 - sel) * b).mean() * rotation_loss_coeff
        else:
            losses["rotation"] = F.mse_loss(rot, gt_quat) * rotation_loss_coeff
    else:
        # the reference's LossAndMetrics supervises quaternions only (main_keypose.py:426-444 indexes gt_action[:, 3:7]
        # against pred["rotation"] and fails on a (B, 3, 3) matrix); training without rotation supervision would be silent
        raise NotImplementedError("keypose_loss: rotation supervision is implemented for quaternion heads "
                                  "(rotation_parametrization='quat_from_query' / 'quat_from_top_ghost'), "
                                  f"got a rotation of shape {tuple(rot.shape)}")
    losses["gripper"] = F.mse_loss(pred["gripper"], gt_action[:, 7:8]) * gripper_loss_coeff
    return losses
