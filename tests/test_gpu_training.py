"""GPU: the training path (csrc/a3d_train.cu behind autograd) against the CPU oracle's autograd.

Kernel level: attention core forward / backward incl. key padding, ragged sizes, multi-chunk dK/dV and
dropout (the kernel's own mask is recovered with a probe call, then a dense torch reference with that
mask must match forward and backward); rotary apply and its transpose; token-gather backward.
Model level: Act3D and the denoiser produce the same outputs as the forward-only inference kernels and
the same parameter gradients as the oracle differentiated on the CPU (fp32; tolerance 1e-3 rel-L2 per
tensor, gradients are fp32 end to end so they land around 1e-5)."""
import pytest
import torch

from oracle import act3d_ref, planner_ref
from oracle.rope import rope3d_table, rotate_pairs
from tests.golden import cases, synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def leaf_state_dict(module):
    """state_dict as autograd leaves for the oracle; tied entries (weight_tying / gp_emb_tying register one
    module under several keys) share ONE leaf so that their gradient contributions are summed like in the module."""
    leaves, sd = {}, {}
    for k, v in module.state_dict().items():
        if v.data_ptr() not in leaves:
            leaves[v.data_ptr()] = v.detach().clone().requires_grad_(v.is_floating_point())
        sd[k] = leaves[v.data_ptr()]
    return sd


def dense_attention(q, k, v, heads, key_mask=None, keep=None, keep_scale=1.0):
    """fp64 reference: q (B,Nq,E), k/v (B,Nk,E), key_mask (B,Nk) bool, keep (B,H,Nq,Nk) 0/1."""
    b, nq, e = q.shape
    hd = e // heads
    qh = q.view(b, nq, heads, hd).transpose(1, 2)
    kh = k.view(b, -1, heads, hd).transpose(1, 2)
    vh = v.view(b, -1, heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    if keep is not None:
        p = p * keep * keep_scale
    return (p @ vh).transpose(1, 2).reshape(b, nq, e)


# the two attention cores of the training path: 0 = tensor cores, error-compensated fp16 pairs (a3d_train_mma.cu,
# default); 1 = fp32 CUDA cores (a3d_train.cu).  Tolerances (rel-L2 against fp64): forward, gradients.
CORE_TOL = {0: (2e-6, 2e-5), 1: (2e-6, 2e-5)}


@pytest.fixture(params=[0, 1], ids=["mma", "simt"])
def train_core(request):
    from act3d_chained_diffuser_b200 import lib
    lib.set_option("train_attn_core", request.param)
    yield request.param
    lib.set_option("train_attn_core", 0)


@pytest.mark.parametrize("b,heads,nq,nk,masked,gscale", [
    (2, 4, 70, 150, False, 1.0), (2, 8, 50, 50, True, 1.0), (3, 4, 1, 333, False, 1.0), (2, 4, 300, 4150, False, 1.0),
    (1, 4, 600, 53, False, 1.0), (2, 8, 33, 129, True, 1.0),
    (2, 4, 333, 300, False, 1e-7), (2, 4, 130, 70, True, 3e4)])        # gradients far outside fp16's normal range
def test_attention_core_forward_backward(train_core, b, heads, nq, nk, masked, gscale):
    from act3d_chained_diffuser_b200.autograd_ops import attention_core
    e = 15 * heads
    q = synth.normal("tr.q", (b, nq, e), 0.6).double()
    k = synth.normal("tr.k", (b, nk, e), 1.0).double()
    v = synth.normal("tr.v", (b, nk, e), 1.0).double()
    g = synth.normal("tr.g", (b, nq, e), 1.0).double() * gscale
    if gscale != 1.0:                                  # rows of very different gradient magnitude inside one tile
        g = g * torch.logspace(0, -4, nq, dtype=torch.float64)[None, :, None]
    mask = None
    if masked:
        mask = torch.zeros(b, nk, dtype=torch.bool)
        mask[0, nk - 7:] = True
        mask[-1, ::3] = True
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    want = dense_attention(qr, kr, vr, heads, mask)
    want.backward(g)
    qc, kc, vc = (t.float().cuda().requires_grad_(True) for t in (q, k, v))
    got = attention_core(qc, kc, vc, heads, mask.cuda() if masked else None)
    got.backward(g.float().cuda())
    torch.cuda.synchronize()
    tol_f, tol_g = CORE_TOL[train_core]
    assert rel(got.detach().cpu().double(), want.detach()) <= tol_f
    for name, a, r in (("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)):
        assert rel(a.cpu().double(), r) <= tol_g, (name, rel(a.cpu().double(), r))
    if gscale != 1.0:                                  # per-row accuracy of dq does not depend on the row's gradient scale
        err = (qc.grad.cpu().double() - qr.grad).norm(dim=-1) / qr.grad.norm(dim=-1).clamp_min(1e-300)
        assert err.max().item() <= 20 * tol_g, err.max().item()


def test_attention_dropout_mask_is_consistent_between_forward_and_backward(train_core):
    from act3d_chained_diffuser_b200 import lib
    b, heads, nq, nk, p = 2, 4, 40, 15, 0.3
    e = 15 * heads
    seed = 1234567
    # probe: q = 0 -> uniform probabilities 1/nk; v = one-hot per key -> o[., d] = keep[., key d] * scale / nk
    q0 = torch.zeros(b, nq, e, device="cuda")
    k0 = torch.zeros(b, nk, e, device="cuda")
    v0 = torch.zeros(b, nk, e, device="cuda")
    for h in range(heads):
        v0[:, torch.arange(nk), h * 15 + torch.arange(nk)] = 1.0
    o, _ = lib.attn_fwd(q0, k0, v0, None, heads, p, seed)
    keep = (o.view(b, nq, heads, 15).transpose(1, 2) * nk * (1 - p)).round().cpu().double()     # (B,H,Nq,Nk)
    assert set(keep.unique().tolist()) <= {0.0, 1.0}
    frac = keep.mean().item()
    assert abs(frac - (1 - p)) < 0.05, frac
    q = synth.normal("dr.q", (b, nq, e), 0.6).double()
    k = synth.normal("dr.k", (b, nk, e), 1.0).double()
    v = synth.normal("dr.v", (b, nk, e), 1.0).double()
    g = synth.normal("dr.g", (b, nq, e), 1.0).double()
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    want = dense_attention(qr, kr, vr, heads, None, keep, 1.0 / (1 - p))
    want.backward(g)
    qc, kc, vc = (t.float().cuda() for t in (q, k, v))
    got, lse = lib.attn_fwd(qc, kc, vc, None, heads, p, seed)
    dq, dk, dv = lib.attn_bwd(qc, kc, vc, None, got, g.float().cuda(), lse, heads, p, seed)
    tol_f, tol_g = CORE_TOL[train_core]
    assert rel(got.cpu().double(), want.detach()) <= tol_f
    for name, a, r in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        assert rel(a.cpu().double(), r) <= tol_g, (name, rel(a.cpu().double(), r))
    # another seed gives another mask
    o2, _ = lib.attn_fwd(q0, k0, v0, None, heads, p, seed + 1)
    assert not torch.equal(o, o2)


@pytest.mark.parametrize("e", [60, 120])
def test_rope_apply_and_transpose(e):
    from act3d_chained_diffuser_b200 import lib
    x = synth.normal("rp.x", (3, 77, e), 1.0)
    pos = synth.points_in_bounds("rp.p", (3, 77))
    tab = rope3d_table(pos, e)
    want = rotate_pairs(x, tab[..., 0], tab[..., 1])
    got = lib.rope_apply(x.cuda(), pos.cuda())
    assert (got.cpu() - want).abs().max() <= 2e-6
    back = lib.rope_apply(got, pos.cuda(), transpose=True)
    assert (back.cpu() - x).abs().max() <= 2e-6
    # the transpose is the adjoint: <R x, y> = <x, R^T y>
    y = synth.normal("rp.y", (3, 77, e), 1.0).cuda()
    lhs = (got * y).sum()
    rhs = (x.cuda() * lib.rope_apply(y, pos.cuda(), transpose=True)).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-3 * abs(lhs.item()) + 1e-3


@pytest.mark.parametrize("channels_last", [False, True])
def test_gather_tokens_backward(channels_last):
    from act3d_chained_diffuser_b200.autograd_ops import gather_tokens
    b, ncam, e, h, w, k = 2, 2, 60, 16, 16, 100
    feat = synth.normal("gt.f", (b * ncam, e, h, w), 1.0)
    pcd = synth.points_in_bounds("gt.p", (b, ncam * h * w))
    idx = torch.stack([torch.randperm(ncam * h * w, generator=torch.Generator().manual_seed(i))[:k] for i in range(b)])
    g = synth.normal("gt.g", (b, k, e), 1.0)
    fr = feat.clone().requires_grad_(True)
    flat = fr.view(b, ncam, e, h * w).permute(0, 1, 3, 2).reshape(b, ncam * h * w, e)
    want = torch.stack([flat[i][idx[i]] for i in range(b)])
    want.backward(g)
    fc = feat.cuda()
    if channels_last:
        fc = fc.contiguous(memory_format=torch.channels_last)
    fc.requires_grad_(True)
    tok, pos = gather_tokens(fc, pcd.cuda(), idx.int().cuda(), b, ncam)
    tok.backward(g.cuda())
    assert torch.equal(tok.detach().cpu(), want.detach())
    assert torch.equal(pos.cpu(), torch.stack([pcd[i][idx[i]] for i in range(b)]))
    assert torch.equal(fc.grad.cpu(), fr.grad)
    # identity gather (level 0): gradient is the transposed copy
    fc2 = feat.cuda().requires_grad_(True)
    tok2, _ = gather_tokens(fc2, pcd.cuda(), None, b, ncam)
    g2 = synth.normal("gt.g2", (b, ncam * h * w, e), 1.0)
    tok2.backward(g2.cuda())
    want2 = g2.view(b, ncam, h * w, e).permute(0, 1, 3, 2).reshape(b * ncam, e, h, w)
    assert torch.equal(fc2.grad.cpu(), want2)


@pytest.mark.parametrize("rows,o,i", [(70000, 60, 60), (5000, 120, 60), (3000, 480, 120), (33, 7, 5), (513, 64, 64)])
def test_linear_weight_gradient_kernel(rows, o, i):
    from act3d_chained_diffuser_b200 import lib
    from act3d_chained_diffuser_b200.autograd_ops import _Linear
    dy = synth.normal("wg.dy", (rows, o), 1.0)
    x = synth.normal("wg.x", (rows, i), 1.0)
    dw, db = lib.linear_wgrad(dy.cuda(), x.cuda())
    want_w, want_b = dy.double().t() @ x.double(), dy.double().sum(0)
    assert rel(dw.cpu().double(), want_w) <= 1e-5 and rel(db.cpu().double(), want_b) <= 1e-5
    # through autograd, on a 3-D input, with and without bias
    w = synth.normal("wg.w", (o, i), 0.1).cuda().requires_grad_(True)
    b = synth.normal("wg.b", (o,), 0.1).cuda().requires_grad_(True)
    xin = x.cuda().view(1, rows, i).requires_grad_(True)
    y = _Linear.apply(xin, w, b, False)
    y.backward(dy.cuda().view(1, rows, o))
    assert rel(w.grad.cpu().double(), want_w) <= 1e-5 and rel(b.grad.cpu().double(), want_b) <= 1e-5
    assert rel(xin.grad.cpu().double()[0], dy.double() @ w.detach().cpu().double()) <= 1e-4
    assert rel(y.detach().cpu()[0], torch.nn.functional.linear(x, w.detach().cpu(), b.detach().cpu())) <= 1e-5


# ------------------------------------------------------------------------------------------------ Act3D
def _act3d(use_instruction, **over):
    from model import Act3D
    kw = dict(cases.ACT3D_KW, use_instruction=use_instruction, **over)
    m = Act3D(**kw).eval()           # eval: the synthetic trunk has no BN, dropout is 0 in Act3D anyway
    cases.install_synth_trunk(m, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict())
    return m, kw


def _act3d_loss(out, tag):
    """Fixed random linear functional of everything the reference's loss reads (main_keypose.py:353-429)."""
    loss = 0.0
    for lvl, masks in enumerate(out["ghost_pcd_masks_pyramid"]):
        for j, mk in enumerate(masks):
            w = synth.normal(f"{tag}.m{lvl}{j}", tuple(mk.shape), 1.0).to(mk.device)
            loss = loss + (torch.log_softmax(mk, dim=-1) * torch.softmax(w, dim=-1)).sum() + 0.05 * (mk * w).sum()
    loss = loss + (out["rotation"] * synth.normal(f"{tag}.r", tuple(out["rotation"].shape), 1.0).to(loss.device)).sum()
    loss = loss + (out["gripper"] * synth.normal(f"{tag}.g", tuple(out["gripper"].shape), 1.0).to(loss.device)).sum()
    if out.get("fine_ghost_pcd_offsets") is not None:
        off = out["fine_ghost_pcd_offsets"]
        loss = loss + (off * synth.normal(f"{tag}.o", tuple(off.shape), 1.0).to(loss.device)).sum()
    return loss


@pytest.mark.parametrize("use_instruction,over", [
    (True, {}),
    (False, dict(rotation_parametrization="6D_from_top_ghost", regress_position_offset=True, weight_tying=False,
                 gp_emb_tying=False))])
def test_act3d_training_gradients_match_oracle(use_instruction, over):
    n_ghost = 200
    m, kw = _act3d(use_instruction, num_ghost_points=3 * n_ghost, num_ghost_points_val=3 * n_ghost, **over)
    inp = cases.act3d_inputs(batch=2, ncam=2, seed=4)
    sampler = synth.make_ghost_sampler(2, n_ghost, seed=4)
    cfg = act3d_ref.Act3DConfig(gripper_loc_bounds=synth.BOUNDS, use_instruction=use_instruction,
                                ghost_points_per_level=n_ghost,
                                rotation_parametrization=kw["rotation_parametrization"],
                                regress_position_offset=kw["regress_position_offset"])
    gt_action = torch.cat([synth.points_in_bounds("tr.gt", (2,), 4), torch.zeros(2, 5)], -1)

    # ---- oracle: autograd on the CPU restatement, parameters = leaf copies of the state dict
    sd = leaf_state_dict(m)
    rgb_ref = inp["visible_rgb"].clone().requires_grad_(True)
    want = act3d_ref.act3d_forward(sd, cfg, act3d_ref.trunk_from_module(m), rgb_ref, inp["visible_pcd"],
                                   inp["instruction"], inp["curr_gripper"], gt_action=gt_action, ghost_sampler=sampler)
    teacher = [p.detach().clone() for p in want["position_pyramid"]]
    _act3d_loss(want, "a3dloss").backward()

    # ---- ours on the GPU
    m = m.cuda()
    m._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor).to(device)
    m._teacher_positions = teacher
    rgb = inp["visible_rgb"].cuda().requires_grad_(True)
    out = m(rgb, inp["visible_pcd"].cuda(), inp["instruction"].cuda(), inp["curr_gripper"].cuda(),
            gt_action=gt_action.cuda())
    for lvl in range(3):
        for j in range(2):
            got, ref = out["ghost_pcd_masks_pyramid"][lvl][j].detach().cpu(), want["ghost_pcd_masks_pyramid"][lvl][j].detach()
            assert rel(got, ref) <= 1e-4, (lvl, j, rel(got, ref))
    _act3d_loss(out, "a3dloss").backward()
    torch.cuda.synchronize()

    checked = 0
    for name, p in m.named_parameters():
        ref = sd[name].grad
        if ref is None or ref.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max() <= 1e-6, name
            continue
        assert p.grad is not None, f"{name}: no gradient reached this parameter"
        assert rel(p.grad.cpu(), ref) <= 1e-3, (name, rel(p.grad.cpu(), ref))
        checked += 1
    assert checked >= (40 if use_instruction else 30), checked
    # gradient through the token gather into the feature maps (and on to the images)
    assert rel(rgb.grad.cpu(), rgb_ref.grad) <= 1e-3

    # ---- the differentiable path and the fused inference kernels agree on the forward
    with torch.no_grad():
        fused = m(inp["visible_rgb"].cuda(), inp["visible_pcd"].cuda(), inp["instruction"].cuda(),
                  inp["curr_gripper"].cuda(), gt_action=gt_action.cuda())
    for lvl in range(3):
        a, b_ = fused["ghost_pcd_masks_pyramid"][lvl][-1], out["ghost_pcd_masks_pyramid"][lvl][-1].detach()
        assert rel(a, b_) <= 2e-3


def test_act3d_training_step_reduces_loss():
    """A few AdamW steps on one fixed batch through the reference's parameter groups shape (engine.py:89-101)."""
    m, kw = _act3d(True, num_ghost_points=3 * 128)
    m = m.cuda().train()
    sampler = synth.make_ghost_sampler(2, 128, seed=8)
    m._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor).to(device)
    inp = {k: v.cuda() for k, v in cases.act3d_inputs(batch=2, ncam=1, seed=8).items()}
    gt = torch.cat([synth.points_in_bounds("tr.gt2", (2,), 8), torch.zeros(2, 5)], -1).cuda()
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    losses = []
    for _ in range(6):
        out = m(inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"], gt_action=gt)
        loss = 0.0
        for lvl in range(3):                                   # soft cross-entropy toward the ghost points nearest gt
            gp = out["ghost_pcd_pyramid"][lvl].transpose(1, 2)
            label = torch.softmax(-(gp - gt[:, None, :3]).norm(dim=-1) / 0.01, dim=-1)
            for mk in out["ghost_pcd_masks_pyramid"][lvl]:
                loss = loss - (label * torch.log_softmax(mk, dim=-1)).sum(-1).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0], losses


def test_graphed_train_step_replays_a_whole_step():
    """train_graph.GraphedTrainStep: forward + keypose loss + backward + capturable AdamW captured in one CUDA graph.
    Every replay advances the device-side sampler counter (fresh ghost points), updates the parameters and lowers the
    loss on a fixed batch; parameters unreachable from the loss are frozen first."""
    from act3d_chained_diffuser_b200.losses import keypose_loss
    from act3d_chained_diffuser_b200.train_graph import GraphedTrainStep, freeze_parameters_without_gradient
    m, kw = _act3d(True, num_ghost_points=3 * 128)
    m = m.cuda().train()
    m.seed_ghost_sampler(3)
    inp = {k: v.cuda() for k, v in cases.act3d_inputs(batch=2, ncam=1, seed=8).items()}
    gt = torch.cat([synth.points_in_bounds("tr.gt3", (2,), 8), torch.nn.functional.normalize(torch.ones(2, 4), dim=-1),
                    torch.ones(2, 1)], -1).cuda()
    batch = (inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"], gt)

    def step_loss(net, rgb, pcd, instr, grip, action):
        return sum(keypose_loss(net(rgb, pcd, instr, grip, gt_action=action), action).values())

    frozen = freeze_parameters_without_gradient(m, lambda net: step_loss(net, *batch))
    assert all(not dict(m.named_parameters())[n].requires_grad for n in frozen)
    step = GraphedTrainStep(m, step_loss, lambda ps: torch.optim.AdamW(ps, lr=1e-3, capturable=True), batch, warmup=3)
    m.seed_ghost_sampler(3)
    before = [p.detach().clone() for p in m.parameters() if p.requires_grad]
    losses = [step(*batch).item() for _ in range(8)]
    assert int(m._graph_counter_buf.item()) == 8 * m.num_sampling_level          # the counter is advanced by the graph itself
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    assert len(set(losses)) == len(losses)                                        # no two replays saw the same ghost points
    changed = sum(int(not torch.equal(a, b)) for a, b in zip(before, (p for p in m.parameters() if p.requires_grad)))
    assert changed >= 0.9 * len(before)


# ------------------------------------------------------------------------------------------------ planner
def _planner(**over):
    from model import DiffusionPlanner
    kw = dict(cases.PLANNER_KW, **over)
    m = DiffusionPlanner(**kw).eval()
    cases.install_synth_trunk(m.prediction_head, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    return m, kw


def test_denoiser_training_gradients_match_oracle():
    m, kw = _planner()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.7)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([37, 5])
    gw = synth.normal("cd.gw", (b, length, 9), 1.0)
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS)
    head = m.prediction_head
    pcd_n = m.normalize_pos(inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()

    sd = leaf_state_dict(head)
    rgb_ref = inp["rgb_obs"].clone().requires_grad_(True)
    ctx = planner_ref.encode_context(sd, cfg, act3d_ref.trunk_from_module(head), rgb_ref, pcd_n, inp["instruction"],
                                     cur, goal)
    want = planner_ref.denoise_once(sd, cfg, ctx, traj, inp["trajectory_mask"], t)
    (want * gw).sum().backward()

    m = m.cuda()
    rgb = inp["rgb_obs"].cuda().requires_grad_(True)
    got = m.prediction_head(traj.cuda(), inp["trajectory_mask"].cuda(), t.cuda(), rgb, pcd_n.cuda(), cur.cuda(),
                            goal.cuda(), inp["instruction"].cuda())[-1]
    assert rel(got.detach().cpu(), want.detach()) <= 1e-4
    (got * gw.cuda()).sum().backward()
    torch.cuda.synchronize()
    checked = 0
    for name, p in m.prediction_head.named_parameters():
        ref = sd[name].grad
        if ref is None or ref.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max() <= 1e-6, name
            continue
        assert p.grad is not None, f"{name}: no gradient reached this parameter"
        assert rel(p.grad.cpu(), ref) <= 1e-3, (name, rel(p.grad.cpu(), ref))
        checked += 1
    assert checked >= 100, checked
    assert rel(rgb.grad.cpu(), rgb_ref.grad) <= 1e-3


def test_planner_training_loss_backward_with_dropout():
    """DiffusionPlanner.forward in train mode (dropout on, incl. the in-kernel attention-weight dropout):
    finite scalar loss, gradients on every trainable tensor the loss depends on, and a step reduces it."""
    m, kw = _planner()
    m = m.cuda().train()
    inp = {k: v.cuda() for k, v in cases.planner_inputs(batch=2, ncam=1, length=12).items()}
    gt = torch.cat([synth.points_in_bounds("cd.gt", (2, 12)), torch.nn.functional.normalize(
        synth.normal("cd.gtq", (2, 12, 4)), dim=-1)], -1).cuda()
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=3e-4)
    torch.manual_seed(0)
    losses = []
    for _ in range(5):
        torch.manual_seed(1)          # same noise / timestep / dropout masks every iteration
        loss = m(gt, inp["trajectory_mask"], inp["rgb_obs"], inp["pcd_obs"], inp["instruction"], inp["curr_gripper"],
                 inp["goal_gripper"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert loss.dim() == 0 and all(torch.isfinite(torch.tensor(losses)))
    with_grad = sum(p.grad is not None and bool(p.grad.abs().max() > 0) for p in params)
    assert with_grad >= 100, with_grad
    assert losses[-1] < losses[0], losses


# ------------------------------------------------------------------------------------------------ objective + golden
def test_keypose_loss_kernel_matches_reference_golden():
    """losses.keypose_loss (soft-CE kernel, forward + gradient in one pass) against the values and gradients the
    reference's own LossAndMetrics produced (tests/golden/keypose_loss.pt)."""
    import os
    from act3d_chained_diffuser_b200.losses import keypose_loss
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "keypose_loss.pt"), weights_only=False)
    for vi, variant in enumerate(cases.LOSS_VARIANTS):
        want = g[f"v{vi}"]
        pred, gt = cases.keypose_loss_case()
        pred["ghost_pcd_pyramid"] = [p.transpose(1, 2).contiguous().cuda().transpose(1, 2) for p in pred["ghost_pcd_pyramid"]]
        pred["ghost_pcd_masks_pyramid"] = [[m.cuda().requires_grad_(True) for m in lvl] for lvl in pred["ghost_pcd_masks_pyramid"]]
        for k in ("rotation", "gripper", "fine_ghost_pcd_offsets"):
            pred[k] = pred[k].cuda().requires_grad_(True)
        pred["position"] = pred["position"].cuda()
        losses = keypose_loss(pred, gt.cuda(), **variant)
        assert set(losses) == set(want["losses"])
        for k, v in losses.items():
            assert abs(v.item() - want["losses"][k].item()) <= 2e-5 * max(1.0, abs(want["losses"][k].item())), k
        sum(losses.values()).backward()
        leaves = [m for lvl in pred["ghost_pcd_masks_pyramid"] for m in lvl]
        for m, ref in zip(leaves, want["dmasks"]):
            got = m.grad.cpu() if m.grad is not None else torch.zeros_like(ref)
            assert (got - ref).abs().max() <= 1e-6 + 1e-4 * ref.abs().max()
        assert rel(pred["rotation"].grad.cpu(), want["drotation"]) <= 1e-5
        assert rel(pred["fine_ghost_pcd_offsets"].grad.cpu(), want["doffsets"]) <= 1e-5


def test_act3d_training_gradients_match_reference_golden():
    """End to end on the GPU: train-mode Act3D + keypose_loss + backward against the parameter gradients of the
    unmodified reference (tests/golden/act3d_train_grads.pt; levels teacher-forced with its positions)."""
    import os
    from model import Act3D
    from act3d_chained_diffuser_b200.losses import keypose_loss
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "act3d_train_grads.pt"), weights_only=False)
    kw = dict(cases.ACT3D_KW, use_instruction=True, num_ghost_points=3 * 96)
    m = Act3D(**kw).train()
    cases.install_synth_trunk(m, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict())
    m = m.cuda()
    inp = cases.act3d_inputs(batch=2, ncam=1)
    gt = cases.keypose_loss_case(batch=2)[1]
    assert synth.checksum(gt, inp["curr_gripper"]) == g["check"]
    sampler = synth.make_ghost_sampler(2, 96)
    m._sample_ghost_points = lambda total_timesteps, device, level, anchor=None: sampler(level, anchor).to(device)
    m._teacher_positions = [p.clone() for p in g["position_pyramid"]]
    out = m(inp["visible_rgb"].cuda(), inp["visible_pcd"].cuda(), inp["instruction"].cuda(), inp["curr_gripper"].cuda(),
            gt_action=gt.cuda())
    losses = keypose_loss(out, gt.cuda())
    for k, v in losses.items():
        assert abs(v.item() - g["losses"][k].item()) <= 1e-4 * max(1.0, abs(g["losses"][k].item())), k
    sum(losses.values()).backward()
    params = dict(m.named_parameters())
    checked = 0
    # gradients that vanish analytically (the last LayerNorm bias of the ghost stack under softmax-CE: the logit
    # gradients sum to zero) are fp32 noise ~1e-8 on both sides: absolute floor relative to the largest gradient
    floor = 1e-6 * max(r.norm().item() for r in g["grads"].values())
    for name, ref in g["grads"].items():
        if ref.abs().max() == 0:
            continue
        got = params[name].grad
        assert got is not None, name
        err = (got.cpu() - ref).norm().item()
        assert err <= 1e-3 * ref.norm().item() + floor, (name, err, ref.norm().item())
        checked += 1
    assert checked >= 60, checked


def test_multiscale_denoiser_training_gradients_match_oracle():
    """feat_scales_to_use=3 through the differentiable path: every refinement and the parameter gradients of the
    summed objective (diffusion_model.py:315-324) against the oracle's autograd."""
    from model import DiffusionPlanner
    kw = cases.PLANNER_MS_KW
    m = DiffusionPlanner(**kw).eval()
    cases.install_synth_trunk(m.prediction_head, kw["embedding_dim"])
    synth.fill_state_dict(m.state_dict(), skip_prefixes=("prediction_head.backbone.",))
    head = m.prediction_head
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    b, length = inp["trajectory_mask"].shape
    traj = synth.normal("cd.traj", (b, length, 9), 0.4)
    cur = synth.normal("cd.cur9", (b, 9), 0.5)
    goal = synth.normal("cd.goal9", (b, 9), 0.5)
    t = torch.tensor([5, 2])
    gws = [synth.normal(f"cd.gw{i}", (b, length, 9), 1.0) for i in range(3)]
    cfg = planner_ref.PlannerConfig(gripper_loc_bounds=synth.BOUNDS, feat_scales_to_use=3, diffusion_timesteps=8)
    pcd_n = m.normalize_pos(inp["pcd_obs"].permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()
    sd = leaf_state_dict(head)
    ctx = planner_ref.encode_context(sd, cfg, act3d_ref.trunk_from_module(head), inp["rgb_obs"], pcd_n, inp["instruction"],
                                     cur, goal)
    want = planner_ref.denoise_all(sd, cfg, ctx, traj, inp["trajectory_mask"], t)
    sum((w_ * g_).sum() for w_, g_ in zip(want, gws)).backward()

    m = m.cuda()
    got = m.prediction_head(traj.cuda(), inp["trajectory_mask"].cuda(), t.cuda(), inp["rgb_obs"].cuda(), pcd_n.cuda(),
                            cur.cuda(), goal.cuda(), inp["instruction"].cuda())
    assert len(got) == 3
    for a_, w_ in zip(got, want):
        assert rel(a_.detach().cpu(), w_.detach()) <= 1e-4
    sum((a_ * g_.cuda()).sum() for a_, g_ in zip(got, gws)).backward()
    torch.cuda.synchronize()
    checked = 0
    floor = 1e-6 * max(v.grad.norm().item() for v in sd.values() if v.grad is not None)
    for name, p in m.prediction_head.named_parameters():
        ref = sd[name].grad
        if ref is None or ref.abs().max() == 0:
            continue
        assert p.grad is not None, f"{name}: no gradient reached this parameter"
        err = (p.grad.cpu() - ref).norm().item()
        assert err <= 1e-3 * ref.norm().item() + floor, (name, err, ref.norm().item())
        checked += 1
    assert checked >= 300, checked


def test_planner_training_matches_reference_golden():
    """DiffusionPlanner.forward (training objective) + backward on the GPU against the loss and the parameter-gradient
    fingerprints of the unmodified reference (tests/golden/planner_train.pt; eval mode = dropout off, pinned noise
    and timesteps)."""
    import os
    from tests.test_oracle_golden import check_grad_fingerprints
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "planner_train.pt"), weights_only=False)
    m, _ = _planner()
    m = m.cuda()
    inp = cases.planner_inputs(batch=2, ncam=1, length=12, masked_tail=3)
    gt = cases.planner_gt_trajectory(batch=2, length=12)
    assert synth.checksum(gt, inp["curr_gripper"]) == g["check"]
    m._noise_fn = synth.NoiseStream("cdtr")
    m._timestep_fn = lambda n: torch.tensor([37, 5])
    loss = m(gt.cuda(), *[inp[k].cuda() for k in ("trajectory_mask", "rgb_obs", "pcd_obs", "instruction", "curr_gripper",
                                                   "goal_gripper")])
    assert abs(loss.item() - g["loss"].item()) <= 1e-4 * abs(g["loss"].item())
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert check_grad_fingerprints(grads, g["grads"], tol=1e-3) >= 200


# ------------------------------------------------------------------------------------------------ row-wise kernels
@pytest.mark.parametrize("rows,o,i", [(5000, 60, 60), (66000, 120, 60), (3001, 480, 120), (3001, 120, 480), (4099, 240, 120)])
def test_linear_rows_kernel_forward_and_data_gradient(rows, o, i):
    """a3d_linear_fwd against fp64: y = x W^T + b [ReLU] and the transposed product dx = dy W."""
    from act3d_chained_diffuser_b200 import lib
    x = synth.normal("lr.x", (rows, i), 1.0)
    w = synth.normal("lr.w", (o, i), 0.3)
    b = synth.normal("lr.b", (o,), 0.5)
    dy = synth.normal("lr.g", (rows, o), 1.0)
    want = x.double() @ w.double().t() + b.double()
    assert lib.linear_supported(o, i) and lib.linear_supported(o, i, transpose=True)
    got = lib.linear_rows(x.cuda(), w.cuda(), b.cuda())
    assert rel(got.cpu().double(), want) <= 2e-6
    got = lib.linear_rows(x.cuda(), w.cuda(), b.cuda(), relu=True)
    assert rel(got.cpu().double(), want.clamp_min(0)) <= 2e-6
    got = lib.linear_rows(x.cuda(), w.cuda(), None)
    assert rel(got.cpu().double(), x.double() @ w.double().t()) <= 2e-6
    dx = lib.linear_rows(dy.cuda(), w.cuda(), transpose=True)
    assert rel(dx.cpu().double(), dy.double() @ w.double()) <= 2e-6


@pytest.mark.parametrize("rows,e,with_res", [(4099, 60, True), (66000, 60, False), (3001, 120, True), (2049, 120, False)])
def test_layernorm_kernels(rows, e, with_res):
    from act3d_chained_diffuser_b200.autograd_ops import residual_layer_norm
    norm = torch.nn.LayerNorm(e)
    with torch.no_grad():
        norm.weight.copy_(1 + synth.normal("ln.w", (e,), 0.2))
        norm.bias.copy_(synth.normal("ln.b", (e,), 0.2))
    x = synth.normal("ln.x", (2, rows // 2 + 1, e), 1.5)[:, : rows // 2]
    res = synth.normal("ln.r", tuple(x.shape), 0.7) if with_res else None
    g = synth.normal("ln.g", tuple(x.shape), 1.0)
    ref = torch.nn.LayerNorm(e).double()
    ref.load_state_dict(norm.state_dict())
    xr = x.double().clone().requires_grad_(True)
    rr = res.double().clone().requires_grad_(True) if with_res else None
    want = ref(xr if rr is None else xr + rr)
    want.backward(g.double())
    norm = norm.cuda()
    xc = x.cuda().requires_grad_(True)
    rc = res.cuda().requires_grad_(True) if with_res else None
    got = residual_layer_norm(norm, xc, rc)
    got.backward(g.cuda())
    assert rel(got.detach().cpu().double(), want.detach()) <= 2e-6
    assert rel(xc.grad.cpu().double(), xr.grad) <= 1e-5
    if with_res:
        assert torch.equal(rc.grad, xc.grad)
    assert rel(norm.weight.grad.cpu().double(), ref.weight.grad) <= 1e-5
    assert rel(norm.bias.grad.cpu().double(), ref.bias.grad) <= 1e-5
