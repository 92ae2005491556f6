"""ChainedDiffuser trajectory planner on the B200 kernels -- drop-in for the reference's
model/trajectory_optimization/diffusion_model.py (+ diffusion_head.py, model/utils/encoder.py):
same ctor kwargs, forward / compute_trajectory signatures and state_dict keys.

What runs where
  PyTorch (host code): frozen backbone + FPN (cuDNN), 512->E instruction / gripper token encoders
      and the adaLN modulation tables (tiny library GEMMs, once per call), quaternion <-> 6D, RNG.
  libact3d_b200.so: point pyramid, token gather, vision->language attention (cd_ctx_lang), the
      rotary K/V cache of the context for all 8 cross-attention layers (a3d_ctx_kv), and per
      denoising step cd_step_begin + 8 x (cd_cross, cd_post) including the DDPM posterior update.
Unlike the reference, which re-runs the backbone, the FPN and the vision-language attention on
every one of the 100 steps (diffusion_head.py:222-224), everything that does not depend on the
trajectory or the timestep is computed once (bit-identical hoist).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn
from torchvision.ops import FeaturePyramidNetwork

from . import lib, train_layers
from .autograd_ops import gather_tokens
from .ddpm import PosteriorTable
from .packing import PackCache, pack_ada_layer, pack_kv_set, pack_lang_layer, pack_mlp, pack_traj_encoder
from .params import ParallelStackParams, mlp
from .rotations import matrix_to_ortho6d, matrix_to_quat, normalise_quat, ortho6d_to_matrix, quat_to_matrix
from .trunk import EvalTrunk, build_backbone, normalize_images


def _repeat(n, tie, factory):
    if tie:
        one = factory()
        return nn.ModuleList(one for _ in range(n))
    return nn.ModuleList(factory() for _ in range(n))


def sinusoidal(t, dim):
    """[sin(t w_k) | cos(t w_k)], w_k = 1e4^(-k/(dim/2-1))   (position_encodings.py:13-20)."""
    half = dim // 2
    w = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * w[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


class DiffusionHead(nn.Module):
    """Denoiser.  Owns the parameters under the reference's key names (diffusion_head.py:12-198,
    encoder.py:14-79) and drives the kernels."""

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
        if feat_scales_to_use not in (1, 2, 3) or attn_rounds < 1:
            raise AssertionError("feat_scales_to_use must be 1, 2 or 3 and attn_rounds >= 1")
        if feat_scales_to_use > 1 and not use_goal:
            # without a goal the reference attends to the WHOLE fine map (65536 points per camera) at scales > 0
            # (diffusion_head.py:253-259); only the local-refinement form (find_traj_nn) is built
            raise NotImplementedError("feat_scales_to_use > 1 needs use_goal=True (local refinement around the trajectory)")
        if use_sigma:
            raise NotImplementedError("DiffusionPlanner (B200): use_sigma=True is not built (accepted: False, the value of every "
                                      "reference entry point)")
        if embedding_dim != 120 or num_attn_heads != 8:
            raise NotImplementedError(f"DiffusionPlanner (B200): embedding_dim={embedding_dim}, num_attn_heads={num_attn_heads} is not "
                                      "built: the sm_100a denoiser kernels are specialised for embedding_dim=120 with 8 heads "
                                      "(head_dim 15), the configuration of scripts/train_trajectory.sh (see INTEGRATION.md, "
                                      "'Supported configurations')")
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
        self._packs = PackCache()
        self.fold_trunk = True
        self._eval_trunk = EvalTrunk()
        self.parallel_heads = False         # position / rotation heads on two streams: measured SLOWER at C3 (16.9k vs 25.0k steps/s)
        self._side_stream_obj = None

    @property
    def _side_stream(self):
        if self._side_stream_obj is None:
            self._side_stream_obj = torch.cuda.Stream()
        return self._side_stream_obj

    # ------------------------------------------------------------------ packed weights / tables
    @property
    def num_offsets(self):
        """(attention round, feature scale) pairs, each with its own stacks unless weight_tying (diffusion_head.py:56-198)."""
        return self.attn_rounds * self.feat_scales

    def _ada_layers(self, off=0):
        return (list(self.traj_attention[off].layers) + list(self.pos_attention[off].layers)
                + list(self.rot_attention[off].layers))

    def _weights(self, num_timesteps, device, off=0):
        params = [p for n_, p in self.named_parameters() if not n_.startswith(("backbone.", "feature_pyramid."))]
        e, h = self.embedding_dim, self.num_attn_heads

        def build():
            layers = self._ada_layers(off)
            ada = [pack_ada_layer(l, e, h) for l in layers]
            kv = [pack_kv_set(l.cross_12, e, h) for l in layers]
            # adaLN modulation of every (timestep, layer): Linear(SiLU(time_emb))   (layers.py:282-290, encoder.py:199)
            t_emb = F.silu(sinusoidal(torch.arange(num_timesteps, device=device), e).float())
            rows = []
            for l in layers:
                per = []
                for nm in ("adaln_12", "adaln_1", "adaln_ff1"):
                    lin = getattr(l, nm).modulation[1]
                    mod = F.linear(t_emb, lin.weight.detach().float(), lin.bias.detach().float())     # (T, 2E)
                    per.append(F.pad(mod.view(num_timesteps, 2, e), (0, 16 * h - e)))                 # (T, 2, EP)
                rows.append(torch.stack(per, dim=1))                                                  # (T, 3, 2, EP)
            out = dict(
                ada_w=[a[0] for a in ada], ada_v=[a[1] for a in ada],
                wkv=torch.stack([k[0] for k in kv]).contiguous(), bkv=torch.stack([k[1] for k in kv]).contiguous(),
                ada=torch.stack(rows, dim=1).contiguous(),                                            # (T, nl, 3, 2, EP)
                traj_enc=pack_traj_encoder(self.traj_encoder, e),
                pos_reg=pack_mlp(self.pos_regressor[off], e), rot_reg=pack_mlp(self.rot_regressor[off], e),
                lang=pack_lang_layer(self.traj_lang_attention[off].layers[0], e, h),
            )
            if self.use_instruction:
                vl = [pack_lang_layer(l, e, h) for l in self.vl_attention[off].layers]
                out["vl_w"] = torch.cat([x[0] for x in vl]).contiguous()
                out["vl_v"] = torch.cat([x[1] for x in vl]).contiguous()
            return out
        return self._packs.get(("planner", num_timesteps, str(device), off), params, build)

    # ------------------------------------------------------------------ step-invariant context
    def encode_context(self, visible_rgb, visible_pcd, instruction, curr_gripper, goal_gripper, num_timesteps,
                       static=None, length=None):
        """Everything of DiffusionHead.forward that does not depend on the trajectory or the timestep
        (diffusion_head.py:221-247, 289-323).  Returns one context record per (round, scale) offset in
        ``ctx["offs"]``: for scale 0 the finished K/V cache of the whole coarse map; for scales > 0 (local
        refinement, needs ``length``) the buffers that ``refresh_local`` fills from the current trajectory.
        ``static``: dict of persistent buffers (kv, lang_k, lang_v) to write into, so that a captured CUDA graph
        of the single-offset sampling loop can be replayed on new inputs."""
        lib.load()
        e, h = self.embedding_dim, self.num_attn_heads
        b, ncam = visible_rgb.shape[:2]
        dev = visible_rgb.device
        rgb = visible_rgb.reshape(b * ncam, *visible_rgb.shape[2:])
        needed = tuple(dict.fromkeys(self.feature_map_pyramid[s] for s in range(self.feat_scales)))
        if self.training or not self.fold_trunk or isinstance(self.backbone, torch.nn.Identity):
            fpn, fpn_bias = self.feature_pyramid(self.backbone(normalize_images(self.normalize, rgb))), {}
        else:
            fpn, fpn_bias = self._eval_trunk(self.normalize, self.backbone, self.feature_pyramid, rgb, needed=needed,
                                             defer_bias=True)
        pcd = visible_pcd.reshape(b * ncam, *visible_pcd.shape[2:]).contiguous().float()
        levels, pts_cache = [], {}
        for s in range(self.feat_scales):
            f = self.downscaling_factor_pyramid[s]
            if f not in pts_cache:
                pts_cache[f] = lib.pcd_pyramid(pcd, f).view(b, -1, 3)
            name = self.feature_map_pyramid[s]
            levels.append(dict(fm=fpn[name].float(), bias=fpn_bias.get(name), pts=pts_cache[f]))
        n_goal = int(self.use_goal)
        cur_tok = self.curr_gripper_encoder(curr_gripper.float()) + self.curr_gripper_embed.weight[0]
        goal_tok = (self.goal_gripper_encoder(goal_gripper.float()) + self.goal_gripper_embed.weight[0]) if self.use_goal else None
        instr = F.linear(instruction.float(), self.instruction_encoder.weight, self.instruction_encoder.bias) \
            if self.use_instruction else None

        def instr_kv(attn):
            wi, bi = attn.in_proj_weight, attn.in_proj_bias
            return (F.linear(instr, wi[e:2 * e], bi[e:2 * e]).contiguous(),
                    F.linear(instr, wi[2 * e:], bi[2 * e:]).contiguous())

        ctx = dict(batch=b, ncam=ncam, levels=levels, offs=[], cur_tok=cur_tok, goal_tok=goal_tok,
                   cur_xyz=curr_gripper[:, :3].float(), goal_xyz=goal_gripper[:, :3].float() if self.use_goal else None)
        for off in range(self.num_offsets):
            scale = off % self.feat_scales
            w = self._weights(num_timesteps, dev, off)
            nl = len(w["ada_w"])
            oc = dict(w=w, scale=scale, batch=b)
            if self.use_instruction:
                kvs = [instr_kv(l.cross_12) for l in self.vl_attention[off].layers]
                oc["vl_k"], oc["vl_v"] = torch.stack([k for k, _ in kvs]), torch.stack([v for _, v in kvs])
                lk, lv = instr_kv(self.traj_lang_attention[off].layers[0].cross_12)
                if static is not None and off == 0:
                    if static.get("lang_k") is not None:
                        static["lang_k"].copy_(lk)
                        static["lang_v"].copy_(lv)
                        lk, lv = static["lang_k"], static["lang_v"]
                    else:
                        static["lang_k"], static["lang_v"] = lk, lv
                oc["lang_k"], oc["lang_v"] = lk, lv
            else:
                oc["lang_k"] = oc["lang_v"] = None
            lvl = levels[scale]
            if scale == 0:
                nvis = lvl["pts"].shape[1]
            else:
                if length is None:
                    raise ValueError("encode_context: local refinement scales need the trajectory length")
                nvis = (64 if scale == 1 else 16) * length            # find_traj_nn, diffusion_head.py:253-259
            rows = nvis + 1 + n_goal
            oc.update(nvis=nvis, nk=rows, set_bytes=lib.kv_bytes(1, b, rows, h),
                      tok=torch.empty(b, rows, e, device=dev), pos=torch.empty(b, rows, 3, device=dev))
            if scale == 0:
                lib.gather_tokens(lvl["fm"], lvl["pts"], None, b, ncam, oc["tok"], oc["pos"], bias=lvl["bias"])
                self._finish_context(ctx, oc, static.get("kv") if (static is not None and off == 0) else None)
                if static is not None and off == 0:
                    static["kv"] = oc["kv"]
            else:
                oc["kv"] = torch.empty(lib.kv_bytes(nl, b, rows, h), device=dev, dtype=torch.uint8)
            ctx["offs"].append(oc)
        # single-offset accessors used by the captured sampling loop
        first = ctx["offs"][0]
        ctx.update(kv=first["kv"], nk=first["nk"], set_bytes=first["set_bytes"], w=first["w"],
                   lang_k=first["lang_k"], lang_v=first["lang_v"])
        return ctx

    def _finish_context(self, ctx, oc, kv_out=None):
        """Visual tokens already gathered into oc["tok"][:, :nvis]: vision->language attention, gripper / goal
        tokens, rotary K/V cache of the 8 cross-attention layers of this offset (diffusion_head.py:305-323)."""
        h = self.num_attn_heads
        w, nvis = oc["w"], oc["nvis"]
        if self.use_instruction:
            lib.cd_ctx_lang(oc["tok"], nvis, oc["vl_k"], oc["vl_v"], w["vl_w"], w["vl_v"], oc["vl_k"].shape[0])
        oc["tok"][:, nvis] = ctx["cur_tok"]
        oc["pos"][:, nvis] = ctx["cur_xyz"]
        if self.use_goal:
            oc["tok"][:, nvis + 1] = ctx["goal_tok"]
            oc["pos"][:, nvis + 1] = ctx["goal_xyz"]
        nl = len(w["ada_w"])
        oc["kv"] = lib.ctx_kv(oc["tok"], oc["pos"], oc["nk"], h, w["wkv"], w["bkv"], [1] * nl,
                              out=kv_out if kv_out is not None else oc.get("kv"))

    def refresh_local(self, ctx, off, traj_xyz):
        """Local refinement context of offset ``off`` (scale > 0): the 64*L / 16*L fine points nearest the current
        trajectory estimate (find_traj_nn, utils.py:38-48 -> a3d_traj_topk), their features, and the K/V cache."""
        oc = ctx["offs"][off]
        lvl = ctx["levels"][oc["scale"]]
        idx = lib.traj_topk(traj_xyz.contiguous().float(), lvl["pts"], oc["nvis"])
        lib.gather_tokens(lvl["fm"], lvl["pts"], idx, ctx["batch"], ctx["ncam"], oc["tok"], oc["pos"], bias=lvl["bias"])
        self._finish_context(ctx, oc)
        return idx

    def denoise_offsets(self, ctx, trajectory, trajectory_mask, t_idx, work):
        """All (round, scale) refinements of one denoiser evaluation: list of (B, L, 9) trajectories
        (diffusion_head.py:249-277).  Token features and rotary positions always come from the INPUT trajectory;
        each offset refines the previous offset's estimate."""
        outs, base = [], trajectory
        for off, oc in enumerate(ctx["offs"]):
            if oc["scale"] > 0:
                self.refresh_local(ctx, off, base[..., :3])
            pos_upd, rot = self.denoise(oc, trajectory, trajectory_mask, t_idx, work)
            base = torch.cat((base[..., :3] + pos_upd, rot), -1)
            outs.append(base)
        return outs

    # ------------------------------------------------------------------ one denoiser evaluation
    def denoise(self, ctx, trajectory, trajectory_mask, t_idx, work, update=None):
        """One forward of the denoiser on (B, L, 9); returns (pos_upd (B,L,3), rot (B,L,6)) or, with
        ``update`` (DDPM step arguments), writes the next trajectory in place of returning.

        The position and rotation heads (2 adaLN layers each, diffusion_head.py:344-356) both start from the
        output of the shared trajectory stack and do not depend on each other, and a cd_post launch is only one
        CTA per sample (32 of 148 SMs at C3): with ``parallel_heads`` the rotation branch runs on a side stream
        next to the position branch (fork after the last shared cross-attention, join before the final launch,
        which needs the position update for the DDPM step).  Bit-identical, but measured slower on B200 (the
        226 KB-smem cd_post CTAs need whole SMs and wait behind the other branch's 1024-CTA cd_cross), so it
        is off by default."""
        w = ctx["w"]                               # ctx: the record of ONE offset (or the single-offset context)
        b, length, _ = trajectory.shape
        n_traj = len(self.traj_attention[0].layers)
        nl = len(w["ada_w"])
        ada = w["ada"]
        xb, att, qb = work["x"], work["att"], work["q"]
        enc1, enc2, enc2_b = w["traj_enc"]
        lang_w, lang_v = w["lang"] if self.use_instruction else (None, None)
        mask = work["mask_u8"]
        kv, sb, nk = ctx["kv"], ctx["set_bytes"], ctx["nk"]

        def post(li, x_in, att_buf, x_out, **kw):
            lib.cd_post(trajectory, mask, work["wp_pe"], t_idx, ada, nl, li, x_in, att_buf, w["ada_w"][li], w["ada_v"][li],
                        x_out, **kw)

        def next_q(li, q_out):
            return dict(next_wq=w["ada_w"][li], next_bq=w["ada_v"][li], next_ada_layer=li, q_out=q_out)

        lib.cd_step_begin(trajectory, work["wp_pe"], t_idx, ada, nl, enc1, enc2, enc2_b, lang_w, lang_v, ctx["lang_k"],
                          ctx["lang_v"], xb[0], w["ada_w"][0], w["ada_v"][0], 0, qb)
        # ---- shared trajectory stack
        for li in range(n_traj - 1):
            lib.cd_cross(qb, kv, li * sb, b, nk, att)
            post(li, xb[li], att, xb[li + 1], **next_q(li + 1, qb))
        last = n_traj - 1
        p0, p1, r0, r1 = n_traj, n_traj + 1, n_traj + 2, n_traj + 3        # the two position / rotation layers
        pos_reg = dict(reg_w=w["pos_reg"][0], reg_v=w["pos_reg"][1], reg_out=work["pos_upd"], reg_dim=3)
        rot_reg = dict(reg_w=w["rot_reg"][0], reg_v=w["rot_reg"][1], reg_out=work["rot_out"], reg_dim=6, update=update)
        lib.cd_cross(qb, kv, last * sb, b, nk, att)
        if not self.parallel_heads:
            post(last, xb[last], att, xb[p0], **next_q(p0, qb))
            lib.cd_cross(qb, kv, p0 * sb, b, nk, att)
            post(p0, xb[p0], att, xb[p0 + 1], **next_q(p1, qb))
            lib.cd_cross(qb, kv, p1 * sb, b, nk, att)
            post(p1, xb[p1], att, xb[p1 + 1], next_src=xb[p0], **pos_reg, **next_q(r0, qb))
            lib.cd_cross(qb, kv, r0 * sb, b, nk, att)
            post(r0, xb[p0], att, xb[r0 + 1], **next_q(r1, qb))
            lib.cd_cross(qb, kv, r1 * sb, b, nk, att)
            post(r1, xb[r1], att, xb[r1 + 1], **rot_reg)
            return work["pos_upd"], work["rot_out"]

        main = torch.cuda.current_stream()
        side = self._side_stream
        att_p, att_r, q_r, x_r = work["att_pos"], work["att_rot"], work["q_rot"], work["x_rot"]
        side.wait_stream(main)
        # last shared layer, evaluated on both streams: each copy also emits the rotary Q of its own branch's first layer
        post(last, xb[last], att, xb[p0], **next_q(p0, qb))
        with torch.cuda.stream(side):
            post(last, xb[last], att, x_r, **next_q(r0, q_r))
            lib.cd_cross(q_r, kv, r0 * sb, b, nk, att_r)
            post(r0, x_r, att_r, xb[r0 + 1], **next_q(r1, q_r))
            lib.cd_cross(q_r, kv, r1 * sb, b, nk, att_r)
        lib.cd_cross(qb, kv, p0 * sb, b, nk, att_p)
        post(p0, xb[p0], att_p, xb[p0 + 1], **next_q(p1, qb))
        lib.cd_cross(qb, kv, p1 * sb, b, nk, att_p)
        post(p1, xb[p1], att_p, xb[p1 + 1], **pos_reg)
        side.wait_stream(main)                 # the DDPM update needs the position head and rewrites the trajectory
        with torch.cuda.stream(side):
            post(r1, xb[r1], att_r, xb[r1 + 1], **rot_reg)
        main.wait_stream(side)
        return work["pos_upd"], work["rot_out"]

    def make_work(self, b, length, trajectory_mask, device):
        e = self.embedding_dim
        nl = len(self._ada_layers())
        mask_u8 = None
        if trajectory_mask is not None and bool(trajectory_mask.any()):
            mask_u8 = trajectory_mask.to(device=device, dtype=torch.uint8).contiguous()
        return dict(
            x=torch.empty(nl + 1, b, 64, e, device=device),
            att=torch.empty(lib.cd_cross_part_floats(b), device=device),      # partial cross-attention results (cd_cross -> cd_post)
            att_pos=torch.empty(lib.cd_cross_part_floats(b), device=device), att_rot=torch.empty(lib.cd_cross_part_floats(b), device=device),
            q=torch.empty(b, 8, 64, 16, device=device, dtype=torch.float16),
            q_rot=torch.empty(b, 8, 64, 16, device=device, dtype=torch.float16), x_rot=torch.empty(b, 64, e, device=device),
            pos_upd=torch.empty(b, length, 3, device=device), rot_out=torch.empty(b, length, 6, device=device),
            wp_pe=sinusoidal(torch.arange(length, device=device), e).float().contiguous(), mask_u8=mask_u8)

    # ------------------------------------------------------------------ reference-compatible forward
    def forward(self, trajectory, trajectory_mask, timestep, visible_rgb, visible_pcd, curr_gripper, goal_gripper,
                instruction):
        """Same contract as diffusion_head.py:200-277: returns a list with one (B, L, 9) tensor."""
        if not trajectory.is_cuda:
            raise RuntimeError("DiffusionHead (B200) runs on CUDA tensors only: there is no CPU fallback path")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(trajectory, trajectory_mask, timestep, visible_rgb, visible_pcd, curr_gripper,
                                       goal_gripper, instruction)
        n_t = int(timestep.max().item()) + 1
        b, length, _ = trajectory.shape
        ctx = self.encode_context(visible_rgb, visible_pcd, instruction, curr_gripper, goal_gripper, max(n_t, 100),
                                  length=length)
        work = self.make_work(b, length, trajectory_mask, trajectory.device)
        traj = trajectory.float().contiguous()
        return self.denoise_offsets(ctx, traj, trajectory_mask, timestep.to(torch.int32).contiguous(), work)


    # ------------------------------------------------------------------ differentiable forward (training)
    def _forward_train(self, trajectory, trajectory_mask, timestep, visible_rgb, visible_pcd, curr_gripper,
                       goal_gripper, instruction):
        """One denoiser evaluation with an autograd graph to every trainable parameter
        (diffusion_head.py:200-363, encoder.py:81-203); returns the list of refinements (one per offset).
        Attention cores (with the reference's 0.1 dropout on the attention weights in train mode), rotary
        embedding and token gather are the kernels of csrc/a3d_train.cu; projections, adaLN, LayerNorm, FFN and
        the regressors are torch.nn ops."""
        lib.load()
        e = self.embedding_dim
        b, ncam = visible_rgb.shape[:2]
        dev = trajectory.device
        training = self.training
        rgb = visible_rgb.reshape(b * ncam, *visible_rgb.shape[2:]).float()
        fpn = self.feature_pyramid(self.backbone(normalize_images(self.normalize, rgb)))
        pcd = visible_pcd.reshape(b * ncam, *visible_pcd.shape[2:]).contiguous().float()
        pts_cache = {}
        for s_ in range(self.feat_scales):
            f = self.downscaling_factor_pyramid[s_]
            if f not in pts_cache:
                pts_cache[f] = lib.pcd_pyramid(pcd, f).view(b, -1, 3)
        instr = self.instruction_encoder(instruction.float()) if self.use_instruction else None
        cur = (self.curr_gripper_encoder(curr_gripper.float()) + self.curr_gripper_embed.weight).unsqueeze(1)
        cur_xyz = curr_gripper[:, None, :3].float()
        if self.use_goal:
            goal = (self.goal_gripper_encoder(goal_gripper.float()) + self.goal_gripper_embed.weight).unsqueeze(1)
            goal_xyz = goal_gripper[:, None, :3].float()

        traj = trajectory.float()
        x0 = self.traj_encoder(traj)                                                  # once, from the input trajectory
        traj_pos = traj[..., :3].detach().contiguous()
        t_emb = sinusoidal(timestep.to(dev), e).float()                               # encoder.py:199
        wp_pe = sinusoidal(torch.arange(traj.shape[1], device=dev), e).float()[None]  # diffusion_head.py:326-328
        mask = trajectory_mask if trajectory_mask is not None and bool(trajectory_mask.any()) else None
        outs, base = [], traj
        for off in range(self.num_offsets):
            scale = off % self.feat_scales
            fm = fpn[self.feature_map_pyramid[scale]].float()
            pts = pts_cache[self.downscaling_factor_pyramid[scale]]
            idx = None
            if scale > 0:                                                             # diffusion_head.py:253-259
                with torch.no_grad():
                    idx = lib.traj_topk(base[..., :3].detach().contiguous(), pts, (64 if scale == 1 else 16) * traj.shape[1])
            ctx, ctx_pos = gather_tokens(fm, pts, idx, b, ncam)
            if self.use_instruction:
                ctx = train_layers.parallel_stack(self.vl_attention[off], ctx, None, instr, training=training)
            ctx = torch.cat([ctx, cur], dim=1)
            ctx_pos = torch.cat([ctx_pos, cur_xyz], dim=1)
            if self.use_goal:
                ctx = torch.cat([ctx, goal], dim=1)
                ctx_pos = torch.cat([ctx_pos, goal_xyz], dim=1)
            x = x0
            if self.use_instruction:                                                  # diffusion_head.py:330-336
                x = train_layers.parallel_stack(self.traj_lang_attention[off], x, mask, instr, sem_pos=wp_pe,
                                                training=training)
            common = dict(x_mask=mask, ctx=ctx, x_pos=traj_pos, ctx_pos=ctx_pos, sem_pos=wp_pe, t_emb=t_emb,
                          training=training)
            x = train_layers.parallel_stack(self.traj_attention[off], x, **common)
            pos_f = train_layers.parallel_stack(self.pos_attention[off], x, **common)
            rot_f = train_layers.parallel_stack(self.rot_attention[off], x, **common)
            upd = torch.cat((self.pos_regressor[off](pos_f), self.rot_regressor[off](rot_f)), -1)
            base = torch.cat((base[..., :3] + upd[..., :3], upd[..., 3:]), -1)        # diffusion_head.py:271-274
            outs.append(base)
        return outs


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
        self.position_noise_scheduler = PosteriorTable("scaled_linear", diffusion_timesteps)
        self.rotation_noise_scheduler = PosteriorTable("squaredcos_cap_v2", diffusion_timesteps)
        self.n_steps = diffusion_timesteps
        self.gripper_loc_bounds = torch.tensor(gripper_loc_bounds)
        self._noise_fn = None          # test hook: callable(shape) -> CPU/GPU tensor, called in the reference's order
        self._timestep_fn = None       # test hook: callable(batch) -> long tensor of training timesteps
        self.rng_compat = False        # True: draw Gaussian noise with the reference's torch.randn call sequence
        self.use_cuda_graph = True     # launch-per-layer path only: capture the 100-step loop once per shape and replay it
        self.persistent_loop = True    # single-offset sampling: the whole loop in ONE persistent cluster kernel (cd_loop.cu)
        self._samplers = {}            # per problem shape: noise / trajectory buffers, K/V cache, captured graph
        self.max_sampler_shapes = 4    # LRU bound on the above (a C3-sized entry holds ~0.5 GB); clear_samplers() frees all

    # ------------------------------------------------------------------ frame conversions (torch, elementwise)
    def normalize_pos(self, pos):
        lo = self.gripper_loc_bounds[0].float().to(pos.device)
        hi = self.gripper_loc_bounds[1].float().to(pos.device)
        return (pos - lo) / (hi - lo) * 2.0 - 1.0

    def unnormalize_pos(self, pos):
        lo = self.gripper_loc_bounds[0].float().to(pos.device)
        hi = self.gripper_loc_bounds[1].float().to(pos.device)
        return (pos + 1.0) / 2.0 * (hi - lo) + lo

    def convert_rot(self, signal):
        """[xyz, quaternion (taken real-first, as the reference does), rest] -> [xyz, 6D, rest]
        (diffusion_model.py:197-213)."""
        quat = normalise_quat(signal[..., 3:7])
        rot = quat_to_matrix(quat)
        lead = rot.shape[:-2]
        r6 = matrix_to_ortho6d(rot.reshape(-1, 3, 3)).reshape(*lead, 6)
        return torch.cat([signal[..., :3], r6, signal[..., 7:]], dim=-1)

    def unconvert_rot(self, signal):
        """[xyz, 6D, rest] -> [xyz, quaternion, rest]   (diffusion_model.py:215-230)."""
        lead = signal.shape[:-1]
        quat = matrix_to_quat(ortho6d_to_matrix(signal[..., 3:9].reshape(-1, 6))).reshape(*lead, 4)
        return torch.cat([signal[..., :3], quat, signal[..., 9:]], dim=-1)

    def _randn(self, shape, device):
        if self._noise_fn is not None:
            return self._noise_fn(tuple(shape)).to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(shape, device=device)

    # ------------------------------------------------------------------ sampling
    def clear_samplers(self):
        """Drop every cached sampling state (persistent buffers and captured CUDA graphs)."""
        self._samplers.clear()

    def _sampler_state(self, b, length, rows_key, dev):
        """Persistent buffers (+ the captured CUDA graph) of the sampling loop for one problem shape."""
        key = (b, length, rows_key, str(dev))
        st = self._samplers.pop(key, None)
        if st is not None:
            self._samplers[key] = st                  # most recently used last
        else:
            while len(self._samplers) >= self.max_sampler_shapes:     # least recently used shape goes (buffers + graph)
                self._samplers.pop(next(iter(self._samplers)))
            n = self.n_steps
            st = dict(traj=torch.empty(b, length, 9, device=dev), cond=torch.empty(b, length, 9, device=dev),
                      cmask=torch.empty(b, length, 9, device=dev, dtype=torch.uint8),
                      noise_pos=torch.empty(n, b, length, 3, device=dev), noise_rot=torch.empty(n, b, length, 6, device=dev),
                      static={}, graph=None, graph_sig=None)
            self._samplers[key] = st
        return st

    def _run_steps(self, ctx, st, work, trajectory_mask, timesteps, t_all):
        head = self.prediction_head
        pc, rc = self.position_noise_scheduler.coef, self.rotation_noise_scheduler.coef
        for k, t in enumerate(timesteps):
            last = k == len(timesteps) - 1
            upd = dict(last_step=int(last), traj_out=st["traj"], pos_upd=work["pos_upd"], cond_data=st["cond"],
                       cond_mask=st["cmask"], coef=[pc[t, 0], pc[t, 1], pc[t, 2], rc[t, 0], rc[t, 1], rc[t, 2]],
                       noise_pos=st["noise_pos"][k], noise_rot=st["noise_rot"][k])
            head.denoise(ctx, st["traj"], trajectory_mask, t_all[k], work, update=upd)

    def _run_steps_multi(self, ctx, st, work, trajectory_mask, timesteps, t_all):
        """Sampling loop with several refinements per step (diffusion_model.py:98-117 over diffusion_head.py:249-277)."""
        head = self.prediction_head
        pc, rc = self.position_noise_scheduler.coef, self.rotation_noise_scheduler.coef
        cmask = st["cmask"].bool()
        for k, t in enumerate(timesteps):
            traj = st["traj"]
            out = head.denoise_offsets(ctx, traj, trajectory_mask, t_all[k], work)[-1]
            out = torch.where(cmask, st["cond"], out)
            if k == len(timesteps) - 1:
                nxt = out
            else:
                pos = pc[t, 0] * out[..., :3].clamp(-1, 1) + pc[t, 1] * traj[..., :3] + pc[t, 2] * st["noise_pos"][k]
                rot = rc[t, 0] * out[..., 3:9].clamp(-1, 1) + rc[t, 1] * traj[..., 3:9] + rc[t, 2] * st["noise_rot"][k]
                nxt = torch.cat((pos, rot), -1)
            st["traj"].copy_(nxt)

    @torch.no_grad()
    def conditional_sample(self, condition_data, condition_mask, fixed_inputs):
        """100-step DDPM ancestral sampling with inpainting of the conditioned waypoints
        (diffusion_model.py:86-119).  The whole loop (17 kernels per step) is captured once per problem shape in
        a CUDA graph and replayed.  Gaussian noise: with ``rng_compat`` (or the test hook ``_noise_fn``) the draws
        are made in the reference's call order -- (B,L,9) once, then per step (B,L,3) and (B,L,6) for t > 0;
        otherwise the per-step tensors are filled with two bulk normal_() calls (same distribution)."""
        trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper = fixed_inputs
        head = self.prediction_head
        dev = condition_data.device
        b, length, _ = condition_data.shape
        self.position_noise_scheduler.set_timesteps(self.n_steps)
        self.rotation_noise_scheduler.set_timesteps(self.n_steps)
        timesteps = self.position_noise_scheduler.timesteps
        has_mask = trajectory_mask is not None and bool(trajectory_mask.any())
        st = self._sampler_state(b, length, (rgb_obs.shape[1], has_mask), dev)
        multi = head.num_offsets > 1
        ctx = head.encode_context(rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper, self.n_steps,
                                  static=None if multi else st["static"], length=length)
        if "work" not in st:
            st["work"] = head.make_work(b, length, trajectory_mask, dev)
            st["t_all"] = torch.tensor(timesteps, device=dev, dtype=torch.int32)[:, None].repeat(1, b).contiguous()
        work = st["work"]
        if has_mask:
            work["mask_u8"].copy_(trajectory_mask.to(torch.uint8))
        st["cond"].copy_(condition_data.float())
        st["cmask"].copy_(condition_mask.to(torch.uint8))
        compat = self.rng_compat or self._noise_fn is not None
        st["traj"].copy_(self._randn(st["cond"].shape, dev) + st["cond"])
        if compat:
            for k in range(len(timesteps) - 1):
                st["noise_pos"][k].copy_(self._randn((b, length, 3), dev))
                st["noise_rot"][k].copy_(self._randn((b, length, 6), dev))
        else:
            st["noise_pos"].normal_()
            st["noise_rot"].normal_()

        if multi:
            # coarse-to-fine refinement (feat_scales_to_use > 1 / attn_rounds > 1): the local context is rebuilt from
            # the running estimate inside every step; launched eagerly (no graph), DDPM update as elementwise torch ops
            self._run_steps_multi(ctx, st, work, trajectory_mask if has_mask else None, timesteps, st["t_all"])
            return st["traj"].clone()
        if self.persistent_loop:
            # the whole loop in one persistent cluster kernel (csrc/cd_loop.cu): trajectory state in shared memory
            head_w = ctx["w"]
            if "coef" not in st:
                pc, rc = self.position_noise_scheduler.coef, self.rotation_noise_scheduler.coef
                st["coef"] = torch.cat([pc, rc], dim=1).float().contiguous().to(dev)
                st["steps_dev"] = torch.tensor(timesteps, device=dev, dtype=torch.int32)
            lang = head_w["lang"] if self.prediction_head.use_instruction else (None, None)
            enc1, enc2, enc2_b = head_w["traj_enc"]
            lib.cd_denoise_loop(st["traj"], st["cond"], st["cmask"], work["mask_u8"] if has_mask else None, work["wp_pe"],
                                st["steps_dev"], head_w["ada"], len(head.traj_attention[0].layers), st["coef"],
                                st["noise_pos"], st["noise_rot"], enc1, enc2, enc2_b, lang[0], lang[1], ctx["lang_k"],
                                ctx["lang_v"], head_w["ada_w"], head_w["ada_v"], head_w["pos_reg"], head_w["rot_reg"],
                                ctx["kv"], ctx["set_bytes"], ctx["nk"])
            return st["traj"].clone()
        if not self.use_cuda_graph:
            self._run_steps(ctx, st, work, trajectory_mask if has_mask else None, timesteps, st["t_all"])
            return st["traj"].clone()
        sig = (ctx["kv"].data_ptr(), id(ctx["w"]), ctx["nk"], bool(getattr(head, "parallel_heads", False)))
        if st["graph"] is None or st["graph_sig"] != sig:
            # one eager evaluation first: sets kernel attributes / warms caches outside the capture
            scratch = st["traj"].clone()
            head.denoise(ctx, scratch, trajectory_mask if has_mask else None, st["t_all"][0], work)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_steps(ctx, st, work, trajectory_mask if has_mask else None, timesteps, st["t_all"])
            # the graph reads the K/V cache and the packed weights; the FPN maps / tokens behind them are not kept
            st["graph"], st["graph_sig"] = g, sig
            st["ctx_keepalive"] = {k: v for k, v in ctx.items() if k not in ("levels", "offs")}
        st["graph"].replay()
        return st["traj"].clone()

    @torch.no_grad()
    def compute_trajectory(self, trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper):
        """(B, L) mask, (B, ncam, 3, H, W) rgb / world-frame pcd, (B, 53, 512), (B, 7) poses -> (B, L, 7)
        (diffusion_model.py:121-185)."""
        if not rgb_obs.is_cuda:
            raise RuntimeError("DiffusionPlanner (B200) runs on CUDA tensors only: there is no CPU fallback path")
        dev = rgb_obs.device
        pcd_n = self.normalize_pos(pcd_obs.float().permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()
        cur = curr_gripper.float().clone()
        goal = goal_gripper.float().clone()
        cur[:, :3] = self.normalize_pos(cur[:, :3])
        goal[:, :3] = self.normalize_pos(goal[:, :3])
        cur, goal = self.convert_rot(cur), self.convert_rot(goal)

        b, d = cur.shape
        length = trajectory_mask.size(1)
        cond = torch.zeros(b, length, d, device=dev)
        cmask = torch.zeros_like(cond)
        cond[:, 0] = cur
        cmask[:, 0] = 1
        if self._use_goal_at_test:
            # last un-padded waypoint of every sample <- goal pose, and it and the padding are inpainted
            # (diffusion_model.py:163-167: cond[i][-n_pad-1] = goal, mask[i][-n_pad-1:] = 1), without a host sync
            last = length - 1 - trajectory_mask.sum(1).long()
            cond[torch.arange(b, device=dev), last] = goal
            cmask[torch.arange(length, device=dev)[None, :] >= last[:, None]] = 1
        traj = self.conditional_sample(cond, cmask.bool(), (trajectory_mask, rgb_obs, pcd_n, instruction, cur, goal))
        traj = self.unconvert_rot(traj)
        traj[:, :, :3] = self.unnormalize_pos(traj[:, :, :3])
        return traj

    def forward(self, gt_trajectory, trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper,
                run_inference=False):
        if run_inference:
            return self.compute_trajectory(trajectory_mask, rgb_obs, pcd_obs, instruction, curr_gripper, goal_gripper)
        # ---- training objective (diffusion_model.py:253-324); differentiable when grad is enabled
        dev = rgb_obs.device
        gt = gt_trajectory.float().clone()
        gt[:, :, :3] = self.normalize_pos(gt[:, :, :3])
        pcd_n = self.normalize_pos(pcd_obs.float().permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3).contiguous()
        cur, goal = curr_gripper.float().clone(), goal_gripper.float().clone()
        cur[:, :3] = self.normalize_pos(cur[:, :3])
        goal[:, :3] = self.normalize_pos(goal[:, :3])
        gt, cur, goal = self.convert_rot(gt), self.convert_rot(cur), self.convert_rot(goal)
        noise = self._randn(gt.shape, dev)
        if self._timestep_fn is not None:
            t = self._timestep_fn(len(noise)).to(dev).long()
        else:
            t = torch.randint(0, self.n_steps, (len(noise),), device=dev).long()
        noisy = torch.cat((self.position_noise_scheduler.add_noise(gt[..., :3], noise[..., :3], t),
                           self.rotation_noise_scheduler.add_noise(gt[..., 3:9], noise[..., 3:9], t)), -1)
        total = 0
        for pred in self.prediction_head(noisy, trajectory_mask, t, rgb_obs, pcd_n, cur, goal, instruction):
            total = total + (100 * F.l1_loss(pred[..., :3], gt[..., :3], reduction="mean")      # diffusion_model.py:315-324
                             + 10 * F.l1_loss(pred[..., 3:9], gt[..., 3:9], reduction="mean"))
        return total
