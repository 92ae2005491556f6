"""CPU: host-side logic of the drop-in modules that needs no kernel -- the planner's bounded per-shape sampler cache,
sync-free goal inpainting (diffusion_model.py:163-167), the unreachable-parameter freeze used for DDP without
find_unused_parameters, device-resident image normalisation, and the supported-configuration errors."""
import pytest
import torch

from tests.golden import cases


def _planner():
    from model import DiffusionPlanner
    return DiffusionPlanner(**cases.PLANNER_KW)


def test_planner_sampler_cache_is_lru_bounded():
    m = _planner()
    dev = torch.device("cpu")
    m.max_sampler_shapes = 3
    states = [m._sampler_state(2, length, (1, False), dev) for length in (5, 6, 7)]
    assert m._sampler_state(2, 5, (1, False), dev) is states[0]           # hit: becomes most recently used
    m._sampler_state(2, 8, (1, False), dev)                               # evicts length 6 (least recently used)
    keys = [k[1] for k in m._samplers]
    assert keys == [7, 5, 8]
    assert m._sampler_state(2, 6, (1, False), dev) is not states[1]
    m.clear_samplers()
    assert len(m._samplers) == 0


def test_goal_inpainting_matches_the_reference_loop():
    """Vectorised form of `cond[i][-n_pad-1] = goal[i]; mask[i][-n_pad-1:] = 1` (diffusion_model.py:163-167)."""
    g = torch.Generator().manual_seed(0)
    b, length, d = 5, 12, 9
    goal = torch.randn(b, d, generator=g)
    pad = torch.tensor([0, 3, 1, 7, 11])
    trajectory_mask = torch.arange(length)[None, :] >= (length - pad)[:, None]
    want_c, want_m = torch.zeros(b, length, d), torch.zeros(b, length, d)
    for i in range(b):                                                     # the reference's loop
        neg = -trajectory_mask[i].sum().long()
        want_c[i][neg - 1] = goal[i]
        want_m[i][neg - 1:] = 1
    cond, cmask = torch.zeros(b, length, d), torch.zeros(b, length, d)
    last = length - 1 - trajectory_mask.sum(1).long()                      # planner.compute_trajectory
    cond[torch.arange(b), last] = goal
    cmask[torch.arange(length)[None, :] >= last[:, None]] = 1
    assert torch.equal(cond, want_c) and torch.equal(cmask, want_m)


def test_freeze_parameters_without_gradient():
    from act3d_chained_diffuser_b200.train_graph import freeze_parameters_without_gradient

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.used = torch.nn.Linear(4, 3)
            self.unused = torch.nn.Linear(4, 3)
            self.frozen = torch.nn.Linear(3, 1)
            self.frozen.weight.requires_grad_(False)

        def forward(self, x):
            return self.frozen(self.used(x)).sum()

    net = Net()
    x = torch.randn(7, 4)
    names = freeze_parameters_without_gradient(net, lambda m: m(x))
    assert sorted(names) == ["unused.bias", "unused.weight"]
    assert not net.unused.weight.requires_grad and net.used.weight.requires_grad and net.frozen.bias.requires_grad
    assert all(p.grad is None for p in net.parameters())                  # left clean for the optimizer / DDP


def test_normalize_images_matches_torchvision():
    from torchvision import transforms
    from act3d_chained_diffuser_b200.trunk import normalize_images
    norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    x = torch.rand(3, 3, 8, 8)
    assert torch.equal(normalize_images(norm, x), norm(x))
    assert torch.equal(normalize_images(torch.nn.Identity(), x), x)


def test_unsupported_configurations_say_what_is_accepted():
    from model import Act3D, DiffusionPlanner
    with pytest.raises(NotImplementedError, match="embedding_dim=60 with 4 heads"):
        Act3D(**dict(cases.ACT3D_KW, embedding_dim=120, num_attn_heads=8))
    with pytest.raises(NotImplementedError, match="embedding_dim=120 with 8 heads"):
        DiffusionPlanner(**dict(cases.PLANNER_KW, embedding_dim=60))
