"""DDPM scheduler: closed-form known answers for the oracle restatement (oracle/ddpm.py), the product's posterior
table (act3d_chained_diffuser_b200/ddpm.py) against it, and -- when the third-party package the reference takes the
scheduler from is installed -- the oracle against ``diffusers.DDPMScheduler`` itself
(reference call sites: model/trajectory_optimization/diffusion_model.py:51-60, 87-88, 99, 111-116, 296-303).
``diffusers`` is absent from this image, so that last test skips here and the oracle stays "parity unpinned"."""
import math

import pytest
import torch

from act3d_chained_diffuser_b200.ddpm import PosteriorTable
from oracle import ddpm as oracle_ddpm

SCHEDULES = ["scaled_linear", "squaredcos_cap_v2"]      # positions, rotations (diffusion_model.py:51-60)


def _oracle(schedule, n=100):
    s = oracle_ddpm.DDPMScheduler(num_train_timesteps=n, beta_schedule=schedule, prediction_type="sample")
    s.set_timesteps(n)
    return s


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_oracle_known_answers(schedule):
    s = _oracle(schedule)
    n = 100
    if schedule == "scaled_linear":
        assert math.isclose(s.betas[0].item(), 1e-4, rel_tol=1e-5) and math.isclose(s.betas[-1].item(), 0.02, rel_tol=1e-5)
        mid = ((1e-4 ** 0.5 + 0.02 ** 0.5) / 2) ** 2                    # linear in sqrt(beta)
        assert math.isclose(((s.betas[49] ** 0.5 + s.betas[50] ** 0.5) / 2).item() ** 2, mid, rel_tol=1e-4)
    else:
        abar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        assert math.isclose(s.alphas_cumprod[9].item(), abar(10 / n) / abar(0), rel_tol=1e-4)   # telescoping product
        assert s.betas.max().item() <= 0.999 + 1e-7
    assert s.timesteps.tolist() == list(range(n - 1, -1, -1))
    for t in range(n):
        c0, ct, sg = (float(v) for v in s.step_coefficients(t))
        a_t = s.alphas_cumprod[t].item()
        a_p = s.alphas_cumprod[t - 1].item() if t else 1.0
        # posterior q(x_{t-1} | x_t, x_0) of Ho et al. 2020 eq. 7: with x_t = sqrt(a_t) x0 + sqrt(1 - a_t) eps the
        # step must reproduce the marginal of x_{t-1}: mean sqrt(a_prev) x0, variance 1 - a_prev
        # (fp32 tables: 1 - a_t is formed by cancellation, so the tolerance scales with eps / (1 - a_t))
        tol = 2e-6 + 2e-7 / (1 - a_t)
        assert math.isclose(c0 + ct * math.sqrt(a_t), math.sqrt(a_p), rel_tol=0, abs_tol=tol)
        if t:
            assert math.isclose(sg * sg + ct * ct * (1 - a_t), 1 - a_p, rel_tol=0, abs_tol=tol)
        else:
            assert (c0, ct, sg) == (1.0, 0.0, 0.0)                        # the last step returns the clipped prediction


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("n_infer", [100, 50, 10])
def test_product_table_matches_oracle(schedule, n_infer):
    table = PosteriorTable(schedule, 100, n_infer)
    s = _oracle(schedule)
    s.set_timesteps(n_infer)
    assert table.timesteps == s.timesteps.tolist()
    assert torch.equal(table.alphas_cumprod, s.alphas_cumprod)
    for t in table.timesteps:
        want = torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in s.step_coefficients(t)])
        assert torch.equal(table.coef[t], want), t
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(7, 5, 3, generator=g), torch.randn(7, 5, 3, generator=g)
    t = torch.randint(0, 100, (7,), generator=g)
    assert torch.equal(table.add_noise(x0, eps, t), s.add_noise(x0, eps, t))
    again = table.coef
    table.set_timesteps(n_infer)                                           # cached: same table object, no recomputation
    assert table.coef is again


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_oracle_matches_diffusers(schedule):
    diffusers = pytest.importorskip("diffusers", reason="diffusers (the reference's scheduler package) is not installed")
    theirs = diffusers.schedulers.scheduling_ddpm.DDPMScheduler(num_train_timesteps=100, beta_schedule=schedule,
                                                              prediction_type="sample")
    theirs.set_timesteps(100)
    ours = _oracle(schedule)
    assert torch.allclose(theirs.alphas_cumprod, ours.alphas_cumprod, rtol=0, atol=0)
    assert theirs.timesteps.tolist() == ours.timesteps.tolist()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 6, 3, generator=g)
    for t in ours.timesteps.tolist():
        pred = torch.randn(4, 6, 3, generator=g) * 1.5                     # exercises clip_sample
        a = theirs.step(pred, t, x, generator=torch.Generator().manual_seed(t)).prev_sample
        b = ours.step(pred, t, x, generator=torch.Generator().manual_seed(t)).prev_sample
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6), t
        x = b
    tt = torch.tensor([0, 17, 99, 50])
    assert torch.allclose(theirs.add_noise(x, pred, tt), ours.add_noise(x, pred, tt), rtol=1e-6, atol=1e-6)
