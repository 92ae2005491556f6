"""Small rotation helpers used on the host side (B x few elements; elementwise torch ops)."""
import torch
import torch.nn.functional as F


def normalise_quat(x):
    """x / max(||x||, 1e-10)   (reference: model/utils/utils.py:51-52)."""
    return x / torch.clamp(x.square().sum(dim=-1).sqrt().unsqueeze(-1), min=1e-10)


def _unit(v):
    return v / torch.sqrt(v.pow(2).sum(1)).clamp_min(1e-8)[:, None]


def ortho6d_to_matrix(o6):
    """(N,6) -> (N,3,3) Gram-Schmidt with columns (x, y, z)   (reference: utils.py:117-130)."""
    x = _unit(o6[:, 0:3])
    z = _unit(torch.linalg.cross(x, o6[:, 3:6], dim=1))
    y = torch.linalg.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


def matrix_to_ortho6d(m):
    """first two columns of the rotation matrix   (reference: utils.py:133-139)."""
    return m[:, :, :2].permute(0, 2, 1).flatten(-2)


def quat_to_matrix(q):
    """real-part-first quaternion -> matrix   (reference: utils/pytorch3d_transforms.py:44-73)."""
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quat(m):
    """matrix -> real-part-first quaternion, best-conditioned candidate
    (reference: utils/pytorch3d_transforms.py:105-164)."""
    lead = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(lead + (9,)), dim=-1)
    t = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                     1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1)
    q_abs = torch.sqrt(t.clamp_min(0)) * (t > 0)
    rows = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    rows = rows / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return rows[best, :].reshape(lead + (4,))
