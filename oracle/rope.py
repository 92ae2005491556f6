"""Oracle: positional encodings (test infrastructure only, see oracle/__init__.py).

Follows model/utils/position_encodings.py of the reference.
"""
import math

import torch


def rope3d_table(xyz: torch.Tensor, embed_dim: int) -> torch.Tensor:
    """3-D rotary table. (..., 3) -> (..., E, 2) with [..., 0]=cos, [..., 1]=sin.

    Reference: RotaryPositionEncoding3D.forward, position_encodings.py:58-97.
    The E channels are split in three axis blocks (x | y | z) of E/3 channels;
    inside a block channel pair (2j, 2j+1) shares the angle p_axis * w_j with
    w_j = exp(-2j * ln(1e4) / (E/3)).
    """
    third = embed_dim // 3
    freq = torch.exp(
        torch.arange(0, third, 2, dtype=torch.float32, device=xyz.device)
        * (-math.log(10000.0) / third)
    ).to(xyz.dtype)
    lead = xyz.shape[:-1]
    cos_blocks, sin_blocks = [], []
    for axis in range(3):
        ang = xyz[..., axis:axis + 1] * freq          # (..., E/6)
        c = torch.cos(ang)
        s = torch.sin(ang)
        # duplicate every angle for the two channels of its pair
        cos_blocks.append(torch.stack([c, c], dim=-1).reshape(*lead, -1))
        sin_blocks.append(torch.stack([s, s], dim=-1).reshape(*lead, -1))
    cos_all = torch.cat(cos_blocks, dim=-1)
    sin_all = torch.cat(sin_blocks, dim=-1)
    return torch.stack([cos_all, sin_all], dim=-1).detach()


def rotate_pairs(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """Apply the rotary rotation on channel pairs (2i, 2i+1) of the full E vector.

    Reference: RotaryPositionEncoding.embed_rotary, position_encodings.py:31-34.
    out[2i] = x[2i] c - x[2i+1] s ; out[2i+1] = x[2i+1] c + x[2i] s.
    """
    even = x[..., 0::2]
    odd = x[..., 1::2]
    swapped = torch.stack([-odd, even], dim=-1).reshape(x.shape)
    return x * cos + swapped * sin


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """(N,) -> (N, dim): [sin(t w_k) | cos(t w_k)], w_k = 1e4^(-k/(dim/2-1)).

    Reference: SinusoidalPosEmb.forward, position_encodings.py:13-20.
    """
    half = dim // 2
    step = math.log(10000) / (half - 1)
    w = torch.exp(torch.arange(half, device=t.device) * -step)
    arg = t[:, None] * w[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)
