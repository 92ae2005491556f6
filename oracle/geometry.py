"""Oracle: geometry helpers (test infrastructure only)."""
import numpy as np
import torch
import torch.nn.functional as F


def pcd_level(pcd_flat: torch.Tensor, factor: int, num_cameras: int) -> torch.Tensor:
    """(B*ncam, 3, H, W) -> (B, ncam*h*w, 3) point pyramid level.

    Reference: act3d.py:379-383 / encoder.py:147-158 (bilinear F.interpolate with
    scale 1/factor, align_corners=False == mean of the two centre pixels per axis).
    """
    lvl = F.interpolate(pcd_flat, scale_factor=1.0 / factor, mode="bilinear")
    bn, c, h, w = lvl.shape
    b = bn // num_cameras
    return lvl.view(b, num_cameras, c, h, w).permute(0, 1, 3, 4, 2).reshape(b, num_cameras * h * w, c)


def pcd_level_closed_form(pcd_flat: torch.Tensor, factor: int, num_cameras: int) -> torch.Tensor:
    """Same level written as the explicit 4-tap mean the CUDA kernel computes: the two centre pixels
    (f*i + f/2 - 1, f*i + f/2) per axis, all weights 0.25, accumulated in raster order
    (((p00 + p01) + p10) + p11) * 0.25 -- bit-identical to torch's CPU bilinear kernel."""
    f = factor
    lo, hi = f // 2 - 1, f // 2
    p00, p01 = pcd_flat[:, :, lo::f, lo::f], pcd_flat[:, :, lo::f, hi::f]
    p10, p11 = pcd_flat[:, :, hi::f, lo::f], pcd_flat[:, :, hi::f, hi::f]
    lvl = (((p00 + p01) + p10) + p11) * 0.25
    bn, c, h, w = lvl.shape
    b = bn // num_cameras
    return lvl.view(b, num_cameras, c, h, w).permute(0, 1, 3, 4, 2).reshape(b, num_cameras * h * w, c)


def local_topk(center: torch.Tensor, points: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k nearest points, ascending distance.  center (B,1,3), points (B,N,3).

    Reference: act3d.py:244-245  (sqrt of the fp32 sum of squares, topk largest=False).
    """
    d = ((center - points) ** 2).sum(-1).sqrt()
    return d.topk(k=k, dim=-1, largest=False).indices


def local_topk_exact(center: np.ndarray, points: np.ndarray, k: int) -> np.ndarray:
    """Bit-exact integer oracle for the CUDA selection kernel: fp32 distance
    ((dx*dx + dy*dy) + dz*dz, IEEE sqrt), ties broken toward the lower index,
    output sorted by (distance, index).  center (B,3), points (B,N,3) float32."""
    c = center.astype(np.float32)[:, None, :]
    p = points.astype(np.float32)
    diff = c - p
    sq = diff * diff
    d = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2], dtype=np.float32)
    order = np.argsort(d, axis=-1, kind="stable")          # stable => lower index first on ties
    return order[:, :k].astype(np.int64), d


def find_traj_nn(trajectory: torch.Tensor, points: torch.Tensor, nn_: int) -> torch.Tensor:
    """Indices of the nn_ * L points closest to ANY waypoint (squared distance to the nearest waypoint,
    ascending).  trajectory (B, L, 3), points (B, P, 3) -> (B, nn_ * L).  Reference: utils.py:38-48."""
    d = ((trajectory[:, :, None] - points[:, None]) ** 2).sum(-1)
    return d.min(1).values.topk(k=nn_ * trajectory.shape[1], dim=-1, largest=False).indices


def normalise_quat(x: torch.Tensor) -> torch.Tensor:
    """Reference: model/utils/utils.py:51-52."""
    return x / torch.clamp(x.square().sum(dim=-1).sqrt().unsqueeze(-1), min=1e-10)


def ortho6d_to_matrix(o6: torch.Tensor) -> torch.Tensor:
    """(N, 6) -> (N, 3, 3), columns (x, y, z).  Reference: utils.py:98-130."""
    def unit(v):
        mag = torch.sqrt(v.pow(2).sum(1)).clamp_min(1e-8)
        return v / mag[:, None]
    xr, yr = o6[:, 0:3], o6[:, 3:6]
    x = unit(xr)
    z = unit(torch.linalg.cross(x, yr, dim=1))
    y = torch.linalg.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


def matrix_to_ortho6d(m: torch.Tensor) -> torch.Tensor:
    """First two columns, concatenated.  Reference: utils.py:133-139."""
    return m[:, :, :2].permute(0, 2, 1).flatten(-2)


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Real-part-first quaternion -> rotation matrix.
    Reference: utils/pytorch3d_transforms.py:44-73."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
    ), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quat(m: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> real-part-first quaternion (best-conditioned candidate).
    Reference: utils/pytorch3d_transforms.py:105-164."""
    lead = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(lead + (9,)), dim=-1)
    raw = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                       1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1)
    q_abs = torch.where(raw > 0, torch.sqrt(raw.clamp_min(0)), torch.zeros_like(raw))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
    ], dim=-2)
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    pick = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(lead + (1, 4))).squeeze(-2)


def sample_cube(bounds: np.ndarray, n: int) -> np.ndarray:
    """U(bounds) per axis with three consecutive np.random.uniform calls.
    Reference: utils.py:68-73 (global numpy RNG, call order x, y, z)."""
    x = np.random.uniform(bounds[0][0], bounds[1][0], n)
    y = np.random.uniform(bounds[0][1], bounds[1][1], n)
    z = np.random.uniform(bounds[0][2], bounds[1][2], n)
    return np.stack([x, y, z], axis=1)


def sample_ball(center: np.ndarray, radius: float, bounds: np.ndarray, n: int) -> np.ndarray:
    """Rejection sampling of the ball inside the clipped box.  Reference: utils.py:76-84."""
    kept = np.empty((0, 3))
    while kept.shape[0] < n:
        pts = sample_cube(bounds, n)
        kept = np.concatenate([kept, pts[np.linalg.norm(pts - center, axis=1) < radius]])
    return kept[:n]
