"""Oracle: training objective of the keypose trainer (test infrastructure only, see oracle/__init__.py).

Plain-torch CPU restatement of LossAndMetrics.compute_loss with position_loss="ce"
(main_keypose.py:353-429).  Pinned by tests/golden/keypose_loss.pt, which tests/golden/make_golden.py
produces by executing the reference's own LossAndMetrics class source on seeded inputs.
"""
import torch
import torch.nn.functional as F


def keypose_loss(pred, gt_action, position_loss_coeff=1.0, position_offset_loss_coeff=10000.0,
                 rotation_loss_coeff=10.0, gripper_loss_coeff=1.0, ground_truth_gaussian_spread=0.01,
                 label_smoothing=0.0, compute_loss_at_all_layers=False, symmetric_rotation_loss=False,
                 rotation_parametrization="quat_from_query"):
    gt_position = gt_action[:, :3]
    losses = {}
    # Gaussian ball around the ground truth as soft label (main_keypose.py:387-396)
    labels = []
    for ghost in pred["ghost_pcd_pyramid"]:                              # (B, 3, Ng)
        l2 = ((ghost - gt_position.unsqueeze(-1)) ** 2).sum(1).sqrt()
        labels.append(torch.softmax(-l2 / ground_truth_gaussian_spread, dim=-1).detach())
    n_layers = len(pred["ghost_pcd_masks_pyramid"][0])
    for j in (range(n_layers) if compute_loss_at_all_layers else [-1]):  # :398-403
        for i, masks in enumerate(pred["ghost_pcd_masks_pyramid"]):
            losses[f"position_ce_level{i}"] = F.cross_entropy(
                masks[j], labels[i], label_smoothing=label_smoothing).mean() * position_loss_coeff / len(labels)
    if pred.get("fine_ghost_pcd_offsets") is not None:                   # :405-417 (equal ghost counts per level)
        with_off = pred["ghost_pcd_pyramid"][-1] + pred["fine_ghost_pcd_offsets"]
        losses["position_offset"] = F.mse_loss(with_off, gt_position.unsqueeze(-1).repeat(1, 1, with_off.shape[-1])) \
            * (position_offset_loss_coeff * position_loss_coeff)
    if "quat" in rotation_parametrization:                               # :369-380
        gt_quat = gt_action[:, 3:7]
        if symmetric_rotation_loss:
            a = F.mse_loss(pred["rotation"], gt_quat, reduction="none").mean(1)
            b = F.mse_loss(pred["rotation"], -gt_quat, reduction="none").mean(1)
            sel = (a < b).float()
            losses["rotation"] = (sel * a + (1 - sel) * b).mean()
        else:
            losses["rotation"] = F.mse_loss(pred["rotation"], gt_quat)
        losses["rotation"] = losses["rotation"] * rotation_loss_coeff
    losses["gripper"] = F.mse_loss(pred["gripper"], gt_action[:, 7:8]) * gripper_loss_coeff   # :364-365
    return losses
