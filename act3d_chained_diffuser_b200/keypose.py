"""Act3D keypose detector on the B200 kernels -- drop-in for the reference's
model/keypose_optimization/act3d.py (same ctor kwargs, forward signature, output dict and
state_dict keys).

Host code is PyTorch (backbone/FPN through cuDNN, tiny token encoders, the action head); the
coarse-to-fine loop itself -- point pyramid, local top-k, gather, K/V cache with 3-D rotary, the
fused ghost-point / query cross-attention stacks, mask logits, top-ghost pick and the ghost
sampler -- runs in libact3d_b200.so with no host synchronisation between levels.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torchvision.ops import FeaturePyramidNetwork

from . import lib, train_layers
from .autograd_ops import gather_tokens
from .packing import PackCache, pack_kv_set, pack_xattn_layer
from .params import XAttnStackParams, mlp
from .rotations import normalise_quat, ortho6d_to_matrix
from .trunk import EvalTrunk, build_backbone, normalize_images


def _shared_or_separate(n, tie, factory):
    if tie:
        one = factory()
        return nn.ModuleList(one for _ in range(n))
    return nn.ModuleList(factory() for _ in range(n))


class Act3D(nn.Module):

    def __init__(self, backbone="clip", image_size=(256, 256), embedding_dim=60, num_attn_heads=4,
                 num_ghost_point_cross_attn_layers=2, num_query_cross_attn_layers=2, num_vis_ins_attn_layers=2,
                 rotation_parametrization="quat_from_query", gripper_loc_bounds=None, num_ghost_points=1000,
                 num_ghost_points_val=10000, weight_tying=True, gp_emb_tying=True, ins_pos_emb=False,
                 num_sampling_level=3, fine_sampling_ball_diameter=0.16, regress_position_offset=False,
                 use_instruction=False):
        super().__init__()
        if backbone not in ("resnet", "clip"):
            raise AssertionError(backbone)
        image_size = tuple(image_size)
        if image_size != (256, 256):
            # the reference asserts {(128,128),(256,256)} but its 128x128 branch is broken
            raise AssertionError("image_size must be (256, 256)")
        if rotation_parametrization not in ("quat_from_top_ghost", "quat_from_query", "6D_from_top_ghost", "6D_from_query"):
            raise AssertionError(rotation_parametrization)
        if num_sampling_level not in (1, 2, 3, 4):
            raise AssertionError(num_sampling_level)
        if embedding_dim != 60 or num_attn_heads != 4:
            raise NotImplementedError(f"Act3D (B200): embedding_dim={embedding_dim}, num_attn_heads={num_attn_heads} is not built: the sm_100a "
                                      "kernels are specialised for embedding_dim=60 with 4 heads (head_dim 15), the configuration every "
                                      "reference script trains and evaluates (see INTEGRATION.md, 'Supported configurations')")
        if ins_pos_emb:
            raise NotImplementedError("Act3D (B200): ins_pos_emb=1 is not built (accepted: ins_pos_emb=0, the value of every reference "
                                      "script; see INTEGRATION.md, 'Supported configurations')")

        self.image_size = image_size
        self.rotation_parametrization = rotation_parametrization
        self.num_ghost_points = num_ghost_points // num_sampling_level
        self.num_ghost_points_val = num_ghost_points_val // num_sampling_level
        self.num_sampling_level = num_sampling_level
        d = fine_sampling_ball_diameter
        self.sampling_ball_diameter_pyramid = [None, d, d / 4.0, d / 16.0]
        self.gripper_loc_bounds = np.array(gripper_loc_bounds)
        self.regress_position_offset = regress_position_offset
        self.weight_tying, self.gp_emb_tying, self.ins_pos_emb = weight_tying, gp_emb_tying, ins_pos_emb
        self.use_instruction = use_instruction
        self.embedding_dim, self.num_attn_heads = embedding_dim, num_attn_heads

        self.backbone, self.normalize = build_backbone(backbone)
        for p in self.backbone.parameters():
            p.requires_grad = False
        self.feature_pyramid = FeaturePyramidNetwork([64, 256, 512, 1024, 2048], embedding_dim)
        self.feature_map_pyramid = ["res3", "res1", "res1", "res1"]
        self.downscaling_factor_pyramid = [8, 2, 2, 2]

        e, h, n = embedding_dim, num_attn_heads, num_sampling_level
        self.ghost_points_embed_pyramid = _shared_or_separate(n, gp_emb_tying, lambda: nn.Embedding(1, e))
        self.curr_gripper_embed = nn.Embedding(1, e)
        self.query_embed = nn.Embedding(1, e)
        self.ghost_point_cross_attn_pyramid = _shared_or_separate(
            n, weight_tying, lambda: XAttnStackParams(e, h, num_ghost_point_cross_attn_layers))
        if use_instruction:
            self.vis_ins_attn_pyramid = _shared_or_separate(
                n, weight_tying, lambda: XAttnStackParams(e, h, num_vis_ins_attn_layers))
        self.query_cross_attn_pyramid = _shared_or_separate(
            n, weight_tying, lambda: XAttnStackParams(e, h, num_query_cross_attn_layers))
        if regress_position_offset:
            self.ghost_point_offset_predictor = mlp(e, e, 3)
        self.rotation_dim = 4 if "quat" in rotation_parametrization else 6
        self.gripper_state_predictor = mlp(e, e, self.rotation_dim + 1)
        if use_instruction:
            self.instruction_encoder = nn.Linear(512, e)

        # not part of the state_dict
        self._packs = PackCache()
        self._sampler_seed = 0x5EED
        self._sampler_calls = 0
        self._teacher_positions = None      # test hook: list of (B,1,3) fed to the next level
        self._last_topk = None              # debug/test hook: top-k indices per level of the last call
        self._profile_events = None         # bench hook: list collecting (tag, start, end) CUDA events
        self.fold_trunk = True              # eval: BN-folded channels-last copy of the frozen backbone (trunk.EvalTrunk)
        self.overlap_query = True           # run the 1-token query stack on a side stream next to the ghost stack
        self.use_cuda_graph = False         # inference on CUDA inputs: capture the whole forward once per input signature and
                                            # replay it (no host launch cost: matters at small batches / online evaluation);
                                            # the big feature / point pyramids of the returned dict then alias static buffers
                                            # that the next call overwrites
        self._graphs = {}
        self.upload_chunks = 1              # use_cuda_graph with HOST inputs: pieces the image upload / trunk are pipelined in
                                            # (measured at C2: 1 piece 12.73 ms, 2: 12.96, 4: 14.10, 8: 15.25 -- the trunk loses more
                                            # on small batches than the overlap wins; profiles/r2_e2e_chunks.jsonl)
        self.train_channels_last = True     # training forward: backbone + FPN on channels-last activations
        self._trunk_nhwc = False
        self._graph_counter = None          # device call counter of the ghost sampler while a graph is being captured
        self._graph_counter_buf = None
        self._side_stream_obj = None
        self._eval_trunk = EvalTrunk()

    @property
    def _side_stream(self):
        if self._side_stream_obj is None:
            self._side_stream_obj = torch.cuda.Stream()
        return self._side_stream_obj

    # ------------------------------------------------------------------ packed weights
    def _stack_pack(self, tag, stack):
        params = list(stack.parameters())
        e, h = self.embedding_dim, self.num_attn_heads

        def build():
            layers = [pack_xattn_layer(stack.attn_layers[l].multihead_attn, stack.attn_layers[l].norm,
                                       stack.ffw_layers[l], e, h) for l in range(stack.num_layers)]
            kv = [pack_kv_set(stack.attn_layers[l].multihead_attn, e, h) for l in range(stack.num_layers)]
            return dict(w=torch.cat([x[0] for x in layers]).contiguous(), v=torch.cat([x[1] for x in layers]).contiguous(),
                        wkv=torch.stack([k[0] for k in kv]), bkv=torch.stack([k[1] for k in kv]))
        return self._packs.get(tag, params, build)

    # ------------------------------------------------------------------ sampling
    def seed_ghost_sampler(self, seed):
        """Seed of the device-side Philox ghost sampler (the reference draws from numpy's global RNG)."""
        self._sampler_seed, self._sampler_calls = int(seed), 0
        if self._graph_counter_buf is not None:
            self._graph_counter_buf.zero_()

    def ensure_sampler_counter(self, device):
        """Allocate the device-side call counter.  Must happen BEFORE a capture starts: allocated inside one, its
        zero-fill would be part of the graph and every replay would draw the same ghost points."""
        if self._graph_counter_buf is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("Act3D.ensure_sampler_counter(device) must be called before the CUDA-graph capture starts")
            self._graph_counter_buf = torch.zeros(1, dtype=torch.int64, device=device)

    @contextlib.contextmanager
    def device_sampler_counter(self, device):
        """While a CUDA graph of a forward is being captured the sampler's call counter must live on the device:
        inside this scope every draw is keyed on (seed, level + device counter), and leaving the scope appends the
        kernel that advances the counter -- so each replay of the graph draws fresh ghost points."""
        self.ensure_sampler_counter(device)
        self._graph_counter = self._graph_counter_buf
        try:
            yield
            lib.counter_add(self._graph_counter_buf, self.num_sampling_level)
        finally:
            self._graph_counter = None

    def _sample_ghost_points(self, total_timesteps, device, level, anchor=None):
        """(B, Ng, 3) uniform ghost points; same contract as act3d.py:394-440 but sampled on the
        device (no anchor.cpu() round trip).  Ng switches on self.training like the reference."""
        n = self.num_ghost_points if self.training else self.num_ghost_points_val
        ctr = self._graph_counter            # inside a CUDA-graph capture: the call counter lives on the device
        if ctr is None:
            self._sampler_calls += 1
            sid = self._sampler_calls
        else:
            sid = level + 1                  # + the device counter, advanced by num_sampling_level per replay
        if level == 0:
            return lib.sample_ghost(None, 0.0, self.gripper_loc_bounds, total_timesteps, n,
                                    self._sampler_seed, sid, device, counter=ctr)
        anc = anchor[:, 0].detach().contiguous().float()
        return lib.sample_ghost(anc, self.sampling_ball_diameter_pyramid[level] / 2, self.gripper_loc_bounds,
                                total_timesteps, n, self._sampler_seed, sid, device, counter=ctr)

    # ------------------------------------------------------------------ visual trunk
    def _trunk(self, visible_rgb):
        """backbone + FPN (PyTorch / cuDNN): (B, ncam, 3, H, W) -> ({level name: feature map}, {level name: deferred bias})."""
        rgb = visible_rgb.reshape(-1, *visible_rgb.shape[2:])
        if self.training or not self.fold_trunk or not isinstance(self.backbone, torch.nn.Module) \
                or isinstance(self.backbone, torch.nn.Identity):
            return self.feature_pyramid(self.backbone(normalize_images(self.normalize, rgb))), {}
        return self._eval_trunk(self.normalize, self.backbone, self.feature_pyramid, rgb,
                                needed=self.feature_map_pyramid[:self.num_sampling_level], defer_bias=True)

    def _compute_visual_features(self, visible_rgb, visible_pcd, num_cameras, staged=None, trunk_out=None):
        """backbone + FPN (PyTorch) and the point pyramid (kernel).  Unlike act3d.py:359-392 no
        rotary table is built here: angles are evaluated inside the K/V kernel for the tokens that
        are actually attended to.  ``trunk_out``: result of an earlier ``_trunk`` call on the same images."""
        b = visible_rgb.shape[0]
        feats, feat_bias = trunk_out if trunk_out is not None else self._trunk(visible_rgb)
        if staged is not None:      # point clouds were uploaded on the copy stream while the backbone ran
            torch.cuda.current_stream().wait_stream(self._side_stream)
            visible_pcd = staged[0]
        pcd = visible_pcd.reshape(b * num_cameras, *visible_pcd.shape[2:]).contiguous().float()
        feats_pyr, pcd_pyr, cache = [], [], {}
        for i in range(self.num_sampling_level):
            f = self.downscaling_factor_pyramid[i]
            if f not in cache:
                cache[f] = lib.pcd_pyramid(pcd, f).view(b, -1, 3)
            fm = feats[self.feature_map_pyramid[i]].float()       # NCHW or channels-last: the gather reads either in place
            feats_pyr.append(fm.unflatten(0, (b, num_cameras)))
            pcd_pyr.append(cache[f])
        # output-convolution biases that the trunk left to the token gather (None: already applied)
        self._feat_bias = [feat_bias.get(self.feature_map_pyramid[i]) for i in range(self.num_sampling_level)]
        return feats_pyr, pcd_pyr

    # ------------------------------------------------------------------ forward
    def forward(self, visible_rgb, visible_pcd, instruction, curr_gripper, gt_action=None):
        """
        visible_rgb (B, ncam, 3, H, W) in [0,1]; visible_pcd (B, ncam, 3, H, W) world xyz;
        instruction (B, 53, 512); curr_gripper (B, 8); gt_action (B, 8) or None.
        Returns the reference's output dict (act3d.py:340-357).
        """
        staged = None
        if (not visible_rgb.is_cuda and self.use_cuda_graph and not torch.is_grad_enabled() and self._graph_ok()
                and next(self.parameters()).is_cuda):
            return self._forward_graphed(visible_rgb, visible_pcd, instruction, curr_gripper, gt_action)
        if not visible_rgb.is_cuda:
            # Host inputs (e.g. straight from a DataLoader with pin_memory): upload here, images first so that the
            # backbone starts while the point clouds / instruction are still crossing PCIe on the copy stream.
            # The computation itself has no CPU path.
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("Act3D (B200) runs on a CUDA device only: there is no CPU fallback path")
            visible_rgb = visible_rgb.to(dev, non_blocking=True)
            copy = self._side_stream
            copy.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy):
                staged = [t.to(dev, non_blocking=True) if t is not None else None
                          for t in (visible_pcd, instruction, curr_gripper, gt_action)]
        lib.load()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # differentiable path: attention core / rotary / gather kernels with their backward (csrc/a3d_train.cu),
            # projections and norms as torch.nn ops; the fused inference kernels below are forward-only
            if staged is not None:
                torch.cuda.current_stream().wait_stream(self._side_stream)
                visible_pcd, instruction, curr_gripper, gt_action = staged
            return self._forward_train(visible_rgb, visible_pcd, instruction, curr_gripper, gt_action)
        if self.use_cuda_graph and staged is None and self._graph_ok():
            return self._forward_graphed(visible_rgb, visible_pcd, instruction, curr_gripper, gt_action)
        return self._forward_infer(visible_rgb, visible_pcd, instruction, curr_gripper, gt_action, staged)

    # ------------------------------------------------------------------ CUDA-graph replay of the inference forward
    def _graph_ok(self):
        return (self._teacher_positions is None and self._profile_events is None
                and "_sample_ghost_points" not in self.__dict__ and not torch.cuda.is_current_stream_capturing())

    def _forward_graphed(self, visible_rgb, visible_pcd, instruction, curr_gripper, gt_action):
        """The forward has no host synchronisation (sampler, top-k, argmax all stay on the device), so the ~140 launches
        of one call are captured once per (input shapes, parameter versions, sampler seed) and replayed.  Ghost points
        stay fresh: the sampler's call counter lives on the device and is advanced inside the graph.
        Two graphs share one memory pool: the trunk (needs the images only) and everything after it -- with HOST inputs
        the images are uploaded first, the trunk graph starts, and the point clouds / instruction / gripper cross PCIe
        on the copy stream underneath it, exactly like the eager staged path."""
        inputs = [visible_rgb, visible_pcd, instruction, curr_gripper, gt_action]
        dev = next(self.parameters()).device
        host = not visible_rgb.is_cuda
        bsz = visible_rgb.shape[0]
        # host inputs: the images cross PCIe in `chunks` pieces and the trunk runs piece by piece behind them
        chunks = max(1, min(int(self.upload_chunks), bsz)) if host else 1
        key = (tuple(tuple(t.shape) + (str(t.dtype),) if t is not None else None for t in inputs), self.training,
               str(dev), self._sampler_seed, self.overlap_query, chunks,
               tuple(p._version for p in self.parameters()), tuple(bf._version for bf in self.buffers()))
        entry = self._graphs.get(key)
        if entry is None:
            static_in = [t.detach().to(dev, copy=True) if t is not None else None for t in inputs]
            bounds = [(bsz * c) // chunks for c in range(chunks + 1)]
            pieces = [static_in[0][lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:])]
            main = torch.cuda.current_stream(dev)
            warm = torch.cuda.Stream(device=dev)
            warm.wait_stream(main)
            with torch.cuda.stream(warm):              # cuDNN autotuning, weight packing, kernel attributes: outside the capture
                for _ in range(2):
                    self._forward_infer(*static_in, None)
                    if chunks > 1:
                        for piece in pieces:
                            self._trunk(piece)
            main.wait_stream(warm)
            torch.cuda.synchronize(dev)
            self.ensure_sampler_counter(dev)
            g_trunks, trunk_outs, pool = [], [], None
            for piece in pieces:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    trunk_outs.append(self._trunk(piece))
                pool = g.pool()
                g_trunks.append(g)
            g_rest = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_rest, pool=pool), self.device_sampler_counter(dev):
                if chunks > 1:                         # per-piece feature maps -> one batch (a 0.1 ms copy at C2)
                    feats = {k: torch.cat([t[0][k] for t in trunk_outs]) for k in trunk_outs[0][0]}
                    trunk_out = (feats, trunk_outs[0][1])
                else:
                    trunk_out = trunk_outs[0]
                out = self._forward_infer(*static_in, None, trunk_out)
            if len(self._graphs) >= 4:                 # bounded: a graph pins its activations (hundreds of MB at batch 16)
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._graphs[key] = (g_trunks, g_rest, static_in, out, trunk_outs, pieces, bounds)
        g_trunks, g_rest, static_in, out, _, pieces, bounds = entry
        main = torch.cuda.current_stream(dev)
        if not host:
            for dst, src in zip(static_in, inputs):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
            g_trunks[0].replay()
        else:
            copy = self._side_stream
            copy.wait_stream(main)                     # the previous replay may still read the static buffers
            landed = []
            with torch.cuda.stream(copy):              # one stream = one PCIe order: image pieces first, then the rest
                for piece, lo, hi in zip(pieces, bounds[:-1], bounds[1:]):
                    piece.copy_(visible_rgb[lo:hi], non_blocking=True)
                    landed.append(copy.record_event())
                for dst, src in zip(static_in[1:], inputs[1:]):
                    if src is not None:
                        dst.copy_(src, non_blocking=True)
            for g, ev in zip(g_trunks, landed):
                main.wait_event(ev)
                g.replay()
            main.wait_stream(copy)
        g_rest.replay()
        keep = ("visible_rgb_features_pyramid", "visible_pcd_pyramid")      # large: returned as views of the static buffers

        def own(v):
            if torch.is_tensor(v):
                return v.clone()
            if isinstance(v, (list, tuple)):
                return [own(x) for x in v]
            return v
        return {k: (v if k in keep else own(v)) for k, v in out.items()}

    def _forward_infer(self, visible_rgb, visible_pcd, instruction, curr_gripper, gt_action=None, staged=None,
                       trunk_out=None):
        e, h = self.embedding_dim, self.num_attn_heads
        b, ncam, _, height, width = visible_rgb.shape
        dev = visible_rgb.device
        gt_position = gt_action[:, :3].unsqueeze(1).detach().float() if gt_action is not None else None
        grip_xyz = curr_gripper[:, :3].float()

        feats_pyr, pcd_pyr = self._compute_visual_features(visible_rgb, visible_pcd, ncam, staged, trunk_out)
        if staged is not None:
            visible_pcd, instruction, curr_gripper, gt_action = staged
            gt_position = gt_action[:, :3].unsqueeze(1).detach().float() if gt_action is not None else None
            grip_xyz = curr_gripper[:, :3].float()

        if self.use_instruction:
            instr = F.linear(instruction.float(), self.instruction_encoder.weight, self.instruction_encoder.bias)
            instr = instr.contiguous()                                       # (B, 53, E)
            n_instr = instr.shape[1]
            instr_zero_pos = torch.zeros(b, n_instr, 3, device=dev)
        else:
            instr, n_instr = None, 0

        n_vis = 32 * 32 * ncam
        rows = n_vis + 1 + n_instr
        position_pyramid, masks_pyramid, ghost_pyramid, carried = [], [], [], []
        topk_log = []
        query_feat = None
        ghost_feats = None
        lg = self.ghost_point_cross_attn_pyramid[0].num_layers
        lq = self.query_cross_attn_pyramid[0].num_layers

        for i in range(self.num_sampling_level):
            anchor = None if i == 0 else (gt_position if gt_position is not None else carried[-1])
            ghost = self._sample_ghost_points(b, dev, level=i, anchor=anchor).contiguous().float()
            ng = ghost.shape[1]

            # ---- context tokens: coarse grid (level 0) or the 1024*ncam nearest fine points
            tok = torch.empty(b, rows, e, device=dev)
            pos = torch.empty(b, rows, 3, device=dev)
            fm = feats_pyr[i].flatten(0, 1)
            if i == 0:
                idx = None
            else:
                idx = lib.local_topk(carried[-1][:, 0].contiguous(), pcd_pyr[i], n_vis)
            topk_log.append(idx)
            lib.gather_tokens(fm, pcd_pyr[i], idx, b, ncam, tok, pos, bias=self._feat_bias[i])
            tok[:, n_vis] = self.curr_gripper_embed.weight[0]
            pos[:, n_vis] = grip_xyz

            if self.use_instruction:
                # context tokens (incl. gripper) cross-attend to the instruction, no rotary (act3d.py:261-270)
                vi = self.vis_ins_attn_pyramid[i]
                pk = self._stack_pack(("vis", id(vi)), vi)
                kv_i = lib.ctx_kv(instr, instr_zero_pos, n_instr, h, pk["wkv"], pk["bkv"], [0] * vi.num_layers)
                lib.xattn_stack(tok, rows * e, e, None, b, n_vis + 1, n_instr, e, h, e, vi.num_layers, kv_i, 0,
                                lib.kv_bytes(1, b, n_instr, h), pk["w"], pk["v"], feat_out=tok, feat_rows=rows)
                tok[:, n_vis + 1:] = instr
                pos[:, n_vis + 1:] = 0.0

            # ---- K/V cache of this context for the ghost layers (rotary) and the query layers
            gs, qs = self.ghost_point_cross_attn_pyramid[i], self.query_cross_attn_pyramid[i]
            pg, pq = self._stack_pack(("ghost", id(gs)), gs), self._stack_pack(("query", id(qs)), qs)
            wkv = torch.cat([pg["wkv"], pq["wkv"]])
            bkv = torch.cat([pg["bkv"], pq["bkv"]])
            q_rot = 0 if i == 0 else 1            # the query is not localised at level 0 (act3d.py:287-294)
            kv = lib.ctx_kv(tok, pos, rows, h, wkv, bkv, [1] * lg + [q_rot] * lq)
            set_bytes = lib.kv_bytes(1, b, rows, h)

            # ---- query token (1 per sample; both layer outputs feed the mask logits).  It is a latency-bound
            #      16-CTA launch, so it runs on a side stream concurrently with the ghost-point stack.
            q_all = torch.empty(lq, b, 1, e, device=dev)
            if i == 0:
                q_x0, q_sb = self.query_embed.weight.detach().float().contiguous(), 0
                q_pos = None
            else:
                q_x0, q_sb = query_feat.contiguous(), e
                q_pos = carried[-1].contiguous()
            main = torch.cuda.current_stream()
            side = self._side_stream if self.overlap_query else None
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    lib.xattn_stack(q_x0, q_sb, 0, q_pos, b, 1, rows, e, h, e, lq, kv, lg * set_bytes, set_bytes,
                                    pq["w"], pq["v"], feat_out=q_all, feat_rows=1, feat_all_layers=True)
            else:
                lib.xattn_stack(q_x0, q_sb, 0, q_pos, b, 1, rows, e, h, e, lq, kv, lg * set_bytes, set_bytes,
                                pq["w"], pq["v"], feat_out=q_all, feat_rows=1, feat_all_layers=True)
            query_feat = q_all[-1, :, 0]                                     # (B, E)

            # ---- ghost points: fused attention stack; mask logits against both query layers
            last_level = i == self.num_sampling_level - 1
            want_feats = last_level and (self.regress_position_offset or "top_ghost" in self.rotation_parametrization)
            logits = torch.empty(lq, b, ng, device=dev)
            g_x0 = self.ghost_points_embed_pyramid[i].weight.detach().float().contiguous()
            prof = self._profile_events
            if prof is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            if side is not None:
                ghost_feats = torch.empty(1, b, ng, e, device=dev)
                lib.xattn_stack(g_x0, 0, 0, ghost, b, ng, rows, e, h, e, lg, kv, 0, set_bytes, pg["w"], pg["v"],
                                feat_out=ghost_feats, feat_rows=ng)
                if prof is not None:
                    ev1.record()
                main.wait_stream(side)
                lib.mask_logits(ghost_feats[0], q_all.view(lq, b, e), logits)
            else:
                ghost_feats = torch.empty(1, b, ng, e, device=dev) if want_feats else None
                lib.xattn_stack(g_x0, 0, 0, ghost, b, ng, rows, e, h, e, lg, kv, 0, set_bytes, pg["w"], pg["v"],
                                feat_out=ghost_feats, feat_rows=ng, qvec=q_all.view(lq, b, e), logits=logits)
                if prof is not None:
                    ev1.record()
            if prof is not None:
                prof.append(("ghost_xattn", ev0, ev1))

            top_idx, top_pos = lib.argmax_pick(logits[-1], ghost)
            position_i = top_pos.unsqueeze(1)
            ghost_pyramid.append(ghost.transpose(1, 2))
            position_pyramid.append(position_i)
            masks_pyramid.append([logits[j] for j in range(lq)])
            carried.append(self._teacher_positions[i].to(dev) if self._teacher_positions is not None else position_i)

        self._last_topk = topk_log

        # ---- offsets + action head (act3d.py:323-337, 507-535): tiny, stays PyTorch
        top_idx_l = top_idx.long()
        ar = torch.arange(b, device=dev)
        offsets = None
        position = top_pos
        if self.regress_position_offset:
            offsets = self.ghost_point_offset_predictor(ghost_feats[0]).permute(0, 2, 1)    # (B, 3, Ng)
            position = position + offsets[ar, :, top_idx_l]
        if "top_ghost" in self.rotation_parametrization:
            feats = ghost_feats[0][ar, top_idx_l]
        else:
            feats = query_feat
        pred = self.gripper_state_predictor(feats)
        if "quat" in self.rotation_parametrization:
            rotation = normalise_quat(pred[:, :self.rotation_dim])
        else:
            rotation = ortho6d_to_matrix(pred[:, :self.rotation_dim])
        gripper = torch.sigmoid(pred[:, self.rotation_dim:])

        return {
            "position": position, "rotation": rotation, "gripper": gripper,
            "position_pyramid": position_pyramid,
            "visible_rgb_mask_pyramid": [None] * self.num_sampling_level,
            "ghost_pcd_masks_pyramid": masks_pyramid,
            "ghost_pcd_pyramid": ghost_pyramid,
            "fine_ghost_pcd_offsets": offsets if self.regress_position_offset else None,
            "visible_rgb_features_pyramid": feats_pyr,
            "visible_pcd_pyramid": pcd_pyr,
            "query_features": query_feat.unsqueeze(0),
            "instruction_features": instr.transpose(0, 1) if instr is not None else None,
            "instruction_dummy_pos": self._identity_rope(b, n_instr, dev) if instr is not None else None,
        }

    # ------------------------------------------------------------------ differentiable forward (training)
    def _forward_train(self, visible_rgb, visible_pcd, instruction, curr_gripper, gt_action=None):
        """Same computation and output dict as forward() with an autograd graph to every trainable
        parameter the reference trains: FPN, embeddings, instruction encoder, the
        three attention stacks and the heads.  Index selection (top-k, argmax) and ghost sampling run in
        the same kernels as inference and carry no gradient (act3d.py:233-314)."""
        e, h = self.embedding_dim, self.num_attn_heads
        b, ncam = visible_rgb.shape[:2]
        dev = visible_rgb.device
        gt_position = gt_action[:, :3].unsqueeze(1).detach().float() if gt_action is not None else None
        grip_xyz = curr_gripper[:, :3].float()

        rgb = visible_rgb.reshape(b * ncam, *visible_rgb.shape[2:]).float()
        x = normalize_images(self.normalize, rgb)
        if self.train_channels_last and isinstance(self.backbone, nn.Module) and not isinstance(self.backbone, nn.Identity):
            # NHWC activations end to end: cuDNN's tensor-core convolutions and batch-norm kernels are NHWC-native, so
            # this drops a layout conversion around every convolution (1.2 ms of a 30 ms step).  Parameters keep their
            # shapes and state_dict; the gather kernel and its backward read either layout.
            if not self._trunk_nhwc:
                self.backbone.to(memory_format=torch.channels_last)
                self.feature_pyramid.to(memory_format=torch.channels_last)
                self._trunk_nhwc = True
            x = x.contiguous(memory_format=torch.channels_last)
        feats = self.feature_pyramid(self.backbone(x))
        pcd = visible_pcd.reshape(b * ncam, *visible_pcd.shape[2:]).contiguous().float()
        feats_pyr, pcd_pyr, cache = [], [], {}
        for i in range(self.num_sampling_level):
            f = self.downscaling_factor_pyramid[i]
            if f not in cache:
                cache[f] = lib.pcd_pyramid(pcd, f).view(b, -1, 3)
            feats_pyr.append(feats[self.feature_map_pyramid[i]].float())
            pcd_pyr.append(cache[f])

        if self.use_instruction:
            instr = self.instruction_encoder(instruction.float())                    # (B, 53, E)
            n_instr = instr.shape[1]
            instr_pos = torch.zeros(b, n_instr, 3, device=dev)
        else:
            instr, n_instr = None, 0
        grip_tok = self.curr_gripper_embed.weight.unsqueeze(0).expand(b, 1, e)
        n_vis = 32 * 32 * ncam

        position_pyramid, masks_pyramid, ghost_pyramid, carried, topk_log = [], [], [], [], []
        query_feat, ghost_feats = None, None
        for i in range(self.num_sampling_level):
            anchor = None if i == 0 else (gt_position if gt_position is not None else carried[-1])
            with torch.no_grad():
                ghost = self._sample_ghost_points(b, dev, level=i, anchor=anchor).contiguous().float()
                idx = None if i == 0 else lib.local_topk(carried[-1][:, 0].contiguous(), pcd_pyr[i], n_vis)
            topk_log.append(idx)
            tok, pos = gather_tokens(feats_pyr[i], pcd_pyr[i], idx, b, ncam)           # (B, n_vis, E), (B, n_vis, 3)
            ctx = torch.cat([tok, grip_tok], dim=1)
            ctx_pos = torch.cat([pos, grip_xyz.unsqueeze(1)], dim=1)
            if self.use_instruction:                                                  # act3d.py:261-270
                ctx = train_layers.xattn_stack(self.vis_ins_attn_pyramid[i], ctx, instr)[-1]
                ctx = torch.cat([ctx, instr], dim=1)
                ctx_pos = torch.cat([ctx_pos, instr_pos], dim=1)

            ng = ghost.shape[1]
            g0 = self.ghost_points_embed_pyramid[i].weight.unsqueeze(0).expand(b, ng, e)
            ghost_feats = train_layers.xattn_stack(self.ghost_point_cross_attn_pyramid[i], g0, ctx, ghost, ctx_pos)[-1]

            if i == 0:                                                                # act3d.py:281-301
                q0 = self.query_embed.weight.unsqueeze(0).expand(b, 1, e)
                q_layers = train_layers.xattn_stack(self.query_cross_attn_pyramid[i], q0, ctx)
            else:
                q_layers = train_layers.xattn_stack(self.query_cross_attn_pyramid[i], query_feat.unsqueeze(1), ctx,
                                                    carried[-1].contiguous(), ctx_pos)
            query_feat = q_layers[-1][:, 0]
            masks = [torch.einsum("be,bne->bn", ql[:, 0], ghost_feats) for ql in q_layers]   # act3d.py:493-494
            with torch.no_grad():
                top_idx, top_pos = lib.argmax_pick(masks[-1].detach().contiguous(), ghost)
            position_i = top_pos.unsqueeze(1)
            ghost_pyramid.append(ghost.transpose(1, 2))
            position_pyramid.append(position_i)
            masks_pyramid.append(masks)
            carried.append(self._teacher_positions[i].to(dev) if self._teacher_positions is not None else position_i)
        self._last_topk = topk_log

        top_idx_l = top_idx.long()
        ar = torch.arange(b, device=dev)
        offsets, position = None, top_pos
        if self.regress_position_offset:
            offsets = self.ghost_point_offset_predictor(ghost_feats).permute(0, 2, 1)        # (B, 3, Ng)
            position = position + offsets[ar, :, top_idx_l]
        feats_head = ghost_feats[ar, top_idx_l] if "top_ghost" in self.rotation_parametrization else query_feat
        pred = self.gripper_state_predictor(feats_head)
        if "quat" in self.rotation_parametrization:
            rotation = normalise_quat(pred[:, :self.rotation_dim])
        else:
            rotation = ortho6d_to_matrix(pred[:, :self.rotation_dim])
        gripper = torch.sigmoid(pred[:, self.rotation_dim:])
        return {
            "position": position, "rotation": rotation, "gripper": gripper,
            "position_pyramid": position_pyramid,
            "visible_rgb_mask_pyramid": [None] * self.num_sampling_level,
            "ghost_pcd_masks_pyramid": masks_pyramid,
            "ghost_pcd_pyramid": ghost_pyramid,
            "fine_ghost_pcd_offsets": offsets if self.regress_position_offset else None,
            "visible_rgb_features_pyramid": [f.unflatten(0, (b, ncam)) for f in feats_pyr],
            "visible_pcd_pyramid": pcd_pyr,
            "query_features": query_feat.unsqueeze(0),
            "instruction_features": instr.transpose(0, 1) if instr is not None else None,
            "instruction_dummy_pos": self._identity_rope(b, n_instr, dev) if instr is not None else None,
        }

    def _identity_rope(self, b, n, dev):
        """Rotary table of zero positions = (cos 1, sin 0) (act3d.py:212-213); kept for output parity."""
        t = torch.zeros(b, n, self.embedding_dim, 2, device=dev)
        t[..., 0] = 1.0
        return t
