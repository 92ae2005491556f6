"""Deterministic synthetic inputs / weights shared by the golden-vector generator (runs in the
build container against the real reference) and by the tests (run anywhere).

Everything is drawn from ``numpy.random.Generator(PCG64(crc32(name) ^ seed))`` so that a
tensor depends only on its *name*, shape and the seed -- never on construction order or on
torch's global RNG.  Fixture files store a checksum of what was generated so that a change
in numpy's stream would be detected rather than silently shifting the inputs.
"""
import hashlib
import zlib

import numpy as np
import torch

# union of tasks/18_peract_tasks_location_bounds.json +-0.04
WORKSPACE_LO = [-0.110, -0.556, 0.713]
WORKSPACE_HI = [0.648, 0.518, 1.512]
BOUNDS = [WORKSPACE_LO, WORKSPACE_HI]


def rng_for(name: str, seed: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF))


def normal(name, shape, scale=1.0, seed=0) -> torch.Tensor:
    return torch.from_numpy((rng_for(name, seed).standard_normal(shape) * scale).astype(np.float32))


def uniform(name, shape, lo=0.0, hi=1.0, seed=0) -> torch.Tensor:
    return torch.from_numpy(rng_for(name, seed).uniform(lo, hi, shape).astype(np.float32))


def points_in_bounds(name, shape_prefix, seed=0, lo=None, hi=None) -> torch.Tensor:
    lo = np.array(WORKSPACE_LO if lo is None else lo)
    hi = np.array(WORKSPACE_HI if hi is None else hi)
    u = rng_for(name, seed).uniform(0.0, 1.0, tuple(shape_prefix) + (3,))
    return torch.from_numpy((lo + u * (hi - lo)).astype(np.float32))


def gripper_pose(name, batch, seed=0, with_open=True) -> torch.Tensor:
    """[xyz in bounds, unit quaternion, open flag] -> (B, 8) or (B, 7)."""
    xyz = points_in_bounds(name + ".xyz", (batch,), seed)
    q = normal(name + ".quat", (batch, 4), seed=seed)
    q = q / q.norm(dim=-1, keepdim=True)
    parts = [xyz, q]
    if with_open:
        parts.append((uniform(name + ".open", (batch, 1), seed=seed) > 0.5).float())
    return torch.cat(parts, dim=-1)


def rgbd(name, batch, ncam, hw=256, seed=0):
    """rgb ~ U[0,1), pcd ~ U[lo,hi] per channel (unstructured: worst case for gather locality)."""
    rgb = uniform(name + ".rgb", (batch, ncam, 3, hw, hw), seed=seed)
    pcd = points_in_bounds(name + ".pcd", (batch, ncam, hw, hw), seed).permute(0, 1, 4, 2, 3).contiguous()
    return rgb, pcd


def fill_state_dict(sd: dict, seed=0, skip_prefixes=("backbone.",), gain=1.0) -> dict:
    """Overwrite every floating tensor of ``sd`` (in place, by key name) with deterministic
    values.  Matrices ~ N(0, gain/sqrt(fan_in)), LayerNorm weights ~ 1 + 0.1 N, biases and
    everything 1-D ~ 0.1 N -- so that zero-initialised paths (adaLN, biases) are exercised
    (SURVEY.md App. B.3).  Tied modules share storage in the module, so writing the same
    name-derived values through one alias keeps them tied: we key on the *first* alias."""
    seen = {}
    for k in sorted(sd.keys()):
        t = sd[k]
        if not torch.is_floating_point(t) or any(k.startswith(p) for p in skip_prefixes):
            continue
        ptr = t.data_ptr()
        if ptr in seen:
            continue
        seen[ptr] = k
        if t.dim() >= 2:
            fan_in = int(np.prod(t.shape[1:]))
            v = normal("w:" + k, tuple(t.shape), gain / np.sqrt(fan_in), seed)
        elif k.endswith("norm.weight") or ".norm_" in k and k.endswith(".weight"):
            v = 1.0 + normal("w:" + k, tuple(t.shape), 0.1, seed)
        else:
            v = normal("w:" + k, tuple(t.shape), 0.1, seed)
        with torch.no_grad():
            t.copy_(v)
    return sd


def grad_summary(name, grad):
    """Compact fingerprint of a gradient tensor: (l2 norm, <grad, r>, full tensor or None) with r ~ N(0,1) keyed on
    the parameter name; small tensors (<= 2048 entries) are kept in full."""
    g = grad.detach().float().cpu()
    r = normal("proj:" + name, tuple(g.shape))
    return dict(norm=g.norm().item(), proj=(g * r).sum().item(), rnorm=r.norm().item(),
                full=g.clone() if g.numel() <= 2048 else None)


def checksum(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()[:16]


class SynthTrunk(torch.nn.Module):
    """Stand-in for backbone+FPN used by the *fixture* configs: returns deterministic feature
    maps (a fixed random projection of box-averaged rgb) with the FPN's shapes, so fixtures do
    not depend on 25M random ResNet weights.  res1: 1/2 resolution, res3: 1/8 resolution."""

    def __init__(self, embed_dim, seed=0):
        super().__init__()
        self.embed_dim = embed_dim
        self.register_buffer("proj1", normal("trunk.proj1", (embed_dim, 12), 0.6, seed))
        self.register_buffer("proj3", normal("trunk.proj3", (embed_dim, 48), 0.3, seed))

    def forward(self, rgb_flat):
        n = rgb_flat.shape[0]
        x1 = torch.nn.functional.pixel_unshuffle(rgb_flat, 2)                  # (n, 12, H/2, W/2)
        x3 = torch.nn.functional.pixel_unshuffle(torch.nn.functional.avg_pool2d(rgb_flat, 2), 4)  # (n, 48, H/8, W/8)
        f1 = torch.einsum("ec,nchw->nehw", self.proj1, x1 - 0.5)
        f3 = torch.einsum("ec,nchw->nehw", self.proj3, x3 - 0.5)
        f1 = torch.tanh(f1) + 0.25 * f1
        f3 = torch.tanh(f3) + 0.25 * f3
        z = rgb_flat.new_zeros(n, self.embed_dim, 1, 1)
        return {"res1": f1.contiguous(), "res2": z, "res3": f3.contiguous(), "res4": z, "res5": z}


def make_ghost_sampler(batch, n_per_level, bounds=None, diameter=0.16, seed=0):
    """Deterministic stand-in for Act3D._sample_ghost_points (act3d.py:394-440): level 0 uniform in
    the workspace box, level >= 1 rejection-sampled in the ball (radius d/2, d = diameter/4^(level-1))
    around the anchor, clipped to the box.  Depends only on (seed, level, sample index, anchor)."""
    b = np.array(BOUNDS if bounds is None else bounds, dtype=np.float64)
    diam = [None, diameter, diameter / 4.0, diameter / 16.0]

    def sample(level, anchor):
        out = np.empty((batch, n_per_level, 3), dtype=np.float64)
        for i in range(batch):
            g = rng_for(f"ghost.l{level}.s{i}", seed)
            if level == 0:
                out[i] = b[0] + g.uniform(0, 1, (n_per_level, 3)) * (b[1] - b[0])
                continue
            c = anchor[i, 0].detach().cpu().double().numpy()
            r = diam[level] / 2
            lo = np.clip(c - r, b[0], b[1])
            hi = np.clip(c + r, b[0], b[1])
            kept = np.empty((0, 3))
            while kept.shape[0] < n_per_level:
                pts = lo + g.uniform(0, 1, (n_per_level, 3)) * (hi - lo)
                kept = np.concatenate([kept, pts[np.linalg.norm(pts - c, axis=1) < r]])
            out[i] = kept[:n_per_level]
        return torch.from_numpy(out).float()
    return sample


class NoiseStream:
    """Gaussian draws in call order; the k-th call gets normal('noise.<tag>.<k>', shape)."""

    def __init__(self, tag="n", seed=0):
        self.tag, self.seed, self.k = tag, seed, 0

    def __call__(self, shape):
        t = normal(f"noise.{self.tag}.{self.k}", tuple(shape), 1.0, self.seed)
        self.k += 1
        return t


class patched_randn:
    """Context manager: torch.randn(...) -> NoiseStream draws (used to drive the unmodified
    reference, whose sampling loop calls torch.randn directly, diffusion_model.py:91-95)."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        self._orig = torch.randn

        def fake(*size, **kw):
            shape = kw.pop("size", None)
            if shape is None:
                shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
            out = self.stream(tuple(shape))
            dev = kw.get("device")
            dt = kw.get("dtype")
            if dt is not None:
                out = out.to(dt)
            if dev is not None:
                out = out.to(dev)
            return out
        torch.randn = fake
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False
