"""torch.autograd bindings of the training kernels (csrc/a3d_train.cu).

Training keeps the reference's parameterisation (ordinary nn.Parameters, stock DDP gradient
all-reduce, engine.py:121-124) and replaces the memory-heavy part of its autograd graph -- the
(B*H, Nq, Nk) attention scores, their softmax and both gradients, and the full-map rotary tables
(multihead_custom_attention.py:391-415, position_encodings.py:58-97) -- by kernels that recompute
probabilities tile by tile.  Everything here is CUDA-only: there is no eager fallback.
"""
import torch

from . import lib


def _seed_from_torch():
    """Dropout seed drawn from torch's CPU generator: follows torch.manual_seed, costs no device sync."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


class _RopeApply(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, pos):
        x = x.contiguous()
        pos = pos.contiguous().float()
        ctx.save_for_backward(pos)
        return lib.rope_apply(x, pos, transpose=False)

    @staticmethod
    def backward(ctx, grad):
        (pos,) = ctx.saved_tensors
        return lib.rope_apply(grad.contiguous(), pos, transpose=True), None


def rope_apply(x, pos):
    """x (B, N, E) fp32, pos (B, N, 3): 3-D rotary embedding of the full E vector
    (position_encodings.py:31-34 over the table of :58-97; positions carry no gradient)."""
    return _RopeApply.apply(x, pos)


class _AttnCore(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, key_mask, heads, dropout_p, seed):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = lib.attn_fwd(q, k, v, key_mask, heads, dropout_p, seed)
        ctx.save_for_backward(q, k, v, o, lse, key_mask if key_mask is not None else torch.empty(0))
        ctx.heads, ctx.dropout_p, ctx.seed, ctx.has_mask = heads, dropout_p, seed, key_mask is not None
        return o

    @staticmethod
    def backward(ctx, grad):
        q, k, v, o, lse, key_mask = ctx.saved_tensors
        dq, dk, dv = lib.attn_bwd(q, k, v, key_mask if ctx.has_mask else None, o, grad.contiguous(), lse, ctx.heads,
                                  ctx.dropout_p, ctx.seed)
        return dq, dk, dv, None, None, None, None


def attention_core(q, k, v, heads, key_padding_mask=None, dropout_p=0.0):
    """softmax(q k^T [+ -inf on masked keys]) v per head without the score tensor.
    q (B, Nq, E) already scaled and rotated, k / v (B, Nk, E); head h = channels [15h, 15h+15)
    (multihead_custom_attention.py:355-359, 391-415).  key_padding_mask (B, Nk) bool, True = ignore."""
    mask = None
    if key_padding_mask is not None:
        mask = key_padding_mask.to(device=q.device, dtype=torch.uint8).contiguous()
    seed = _seed_from_torch() if dropout_p > 0.0 else 0
    return _AttnCore.apply(q, k, v, mask, heads, float(dropout_p), seed)


class _GatherTokens(torch.autograd.Function):

    @staticmethod
    def forward(ctx, feat, pcd, idx, batch, ncam, k):
        e = feat.shape[1]
        tok = torch.empty(batch, k, e, device=feat.device)
        pos = torch.empty(batch, k, 3, device=feat.device)
        if not (feat.is_contiguous() or feat.is_contiguous(memory_format=torch.channels_last)):
            feat = feat.contiguous()
        lib.gather_tokens(feat, pcd, idx, batch, ncam, tok, pos)
        ctx.save_for_backward(idx if idx is not None else torch.empty(0, dtype=torch.int32))
        ctx.meta = (batch, ncam, k, idx is not None, tuple(feat.shape), not feat.is_contiguous())
        ctx.mark_non_differentiable(pos)
        return tok, pos

    @staticmethod
    def backward(ctx, dtok, _dpos):
        (idx,) = ctx.saved_tensors
        batch, ncam, k, has_idx, shape, channels_last = ctx.meta
        dfeat = lib.gather_tokens_bwd(dtok.contiguous(), idx if has_idx else None, batch, ncam, k, shape, channels_last)
        return dfeat, None, None, None, None, None


def gather_tokens(feat, pcd, idx, batch, ncam):
    """feat (B*ncam, E, h, w) [requires grad], pcd (B, ncam*h*w, 3), idx (B, K) int32 or None ->
    tok (B, K, E) differentiable w.r.t. feat, pos (B, K, 3)   (act3d.py:236-254)."""
    k = idx.shape[1] if idx is not None else ncam * feat.shape[2] * feat.shape[3]
    return _GatherTokens.apply(feat, pcd, idx, batch, ncam, k)


class _Linear(torch.autograd.Function):
    """y = x W^T + b [ReLU] for the many-row token tensors of the training path.  Forward and the data gradient are the
    tensor-core kernel a3d_linear_fwd (error-compensated fp16 pairs: fp32-class accuracy; autograd's GEMMs are fp32
    SIMT because TF32 is off, as in the reference); the weight / bias gradient is a3d_linear_wgrad (row-split
    reduction at HBM speed instead of a single-CTA SIMT GEMM + separate bias reduction)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = weight.contiguous()
        b = bias.contiguous() if bias is not None else None
        if lib.linear_supported(w.shape[0], w.shape[1]):
            y = lib.linear_rows(x2, w, b, relu=relu)
        else:
            y = torch.nn.functional.linear(x2, w, b)
            if relu:
                y = torch.relu_(y)
        ctx.save_for_backward(x2, w, y if relu else torch.empty(0))
        ctx.has_bias, ctx.relu, ctx.x_shape = bias is not None, relu, x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, grad):
        x2, weight, y = ctx.saved_tensors
        g2 = grad.reshape(-1, grad.shape[-1])
        g2 = (g2 * (y > 0)) if ctx.relu else g2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            if lib.linear_supported(weight.shape[0], weight.shape[1], transpose=True):
                dx = lib.linear_rows(g2, weight, transpose=True).view(ctx.x_shape)
            else:
                dx = (g2 @ weight).view(ctx.x_shape)
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if weight.numel() <= 128 * 128:      # one to four 64 x 64 output tiles: the row-split kernel
                dw, db = lib.linear_wgrad(g2, x2, want_bias=ctx.has_bias)
            else:                                # wide FFN matrices give a library GEMM enough output tiles
                dw, db = g2.t() @ x2, (g2.sum(0) if ctx.has_bias else None)
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False):
    """torch.nn.functional.linear (+ optional fused ReLU) for the training path; rows = every leading dimension of x."""
    if x.numel() // x.shape[-1] < 2048:          # few rows: autograd's own GEMM is as good
        y = torch.nn.functional.linear(x, weight, bias)
        return torch.relu(y) if relu else y
    return _Linear.apply(x, weight, bias, relu)


class _ResidualLayerNorm(torch.autograd.Function):
    """LayerNorm(x + res) * g + b in one pass (a3d_layernorm_fwd), backward in one pass + a fixed-order column sum
    (a3d_layernorm_bwd); `res` may be None."""

    @staticmethod
    def forward(ctx, x, res, weight, bias, eps):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        r2 = res.reshape(-1, x.shape[-1]).contiguous() if res is not None else None
        y, z, mean, rstd = lib.layernorm_fwd(x2, r2, weight.contiguous(), bias.contiguous(), eps)
        ctx.save_for_backward(z, mean, rstd, weight)
        ctx.has_res = res is not None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, grad):
        z, mean, rstd, weight = ctx.saved_tensors
        g2 = grad.reshape(-1, grad.shape[-1]).contiguous()
        dz, dg, db = lib.layernorm_bwd(g2, z, mean, rstd, weight.contiguous())
        dz = dz.view(grad.shape)
        return dz, (dz if ctx.has_res else None), dg, db, None


def residual_layer_norm(norm, x, res=None):
    """``norm(x + res)`` for an nn.LayerNorm over the last dimension (layers.py:309, 331, 147, 182, 209)."""
    e = x.shape[-1]
    if e not in (60, 120) or x.numel() // e < 2048 or norm.weight is None:
        return norm(x if res is None else x + res)
    return _ResidualLayerNorm.apply(x, res, norm.weight, norm.bias, norm.eps)
