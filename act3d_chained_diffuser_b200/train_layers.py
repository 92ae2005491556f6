"""Differentiable (training-mode) layer stacks over the parameter containers of params.py.

Batch-first restatement of the reference's layer zoo for the training path: the attention core,
the rotary embedding and the token gather are the kernels of csrc/a3d_train.cu (autograd_ops.py);
the E x E projections, LayerNorms, adaLN modulations and FFNs are torch.nn.functional calls on the
modules' own nn.Parameters, so autograd / DDP / AdamW see exactly the reference's parameter set.

Reference: RelativeCrossAttentionLayer/Module + FeedforwardLayer (model/utils/layers.py:293-351),
ParallelAttentionLayer / ParallelAttention / AdaLN (layers.py:7-290),
MultiheadCustomAttention (model/utils/multihead_custom_attention.py:157-462).
"""
import torch.nn.functional as F

from .autograd_ops import attention_core, linear, residual_layer_norm, rope_apply


def mha(attn, heads, query, key, value, q_pos=None, k_pos=None, key_padding_mask=None, dropout_p=0.0):
    """query (B, Nq, E), key / value (B, Nk, E); *_pos (B, N, 3) or None (no rotary).
    Projection slices W[0:E], W[E:2E], W[2E:3E] whatever the q/k/v aliasing
    (multihead_custom_attention.py:247-303); q scaled by head_dim^-1/2 before the rotation (:325, :348-353)."""
    e = query.shape[-1]
    w, b = attn.in_proj_weight, attn.in_proj_bias
    q = linear(query, w[:e], b[:e]) * (float(e // heads) ** -0.5)
    if key is value:        # one (2E x E) projection like the reference's kv_same branch (:268-275): half the GEMM launches
        k, v = linear(key, w[e:], b[e:]).chunk(2, dim=-1)
    else:
        k = linear(key, w[e:2 * e], b[e:2 * e])
        v = linear(value, w[2 * e:], b[2 * e:])
    if q_pos is not None:
        q = rope_apply(q, q_pos)
        k = rope_apply(k, k_pos)
    o = attention_core(q, k, v, heads, key_padding_mask, dropout_p)
    return linear(o, attn.out_proj.weight, attn.out_proj.bias)


def xattn_stack(stack, x, ctx, q_pos=None, k_pos=None):
    """Act3D stack: per layer x <- LN(x + MHA(x, ctx, ctx)); x <- LN(x + W2 relu(W1 x)).
    Rotary only when the query has a position (layers.py:300-310, 328-332, 345-351; dropout 0).
    Returns the list of per-layer outputs, each (B, Nq, E)."""
    outs = []
    for l in range(stack.num_layers):
        al, fl = stack.attn_layers[l], stack.ffw_layers[l]
        rot = q_pos is not None
        att = mha(al.multihead_attn, stack.num_heads, x, ctx, ctx, q_pos if rot else None, k_pos if rot else None)
        x = residual_layer_norm(al.norm, x, att)
        x = residual_layer_norm(fl.norm, x, linear(linear(x, fl.linear1.weight, fl.linear1.bias, relu=True),
                                                   fl.linear2.weight, fl.linear2.bias))
        outs.append(x)
    return outs


def ada_ln(mod, x, t_emb):
    """x (B, N, C) * (1 + scale) + shift, [scale | shift] = Linear(SiLU(t_emb))   (layers.py:273-290)."""
    scale, shift = mod.modulation(t_emb).chunk(2, dim=-1)
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def parallel_stack(stack, x, x_mask, ctx, x_pos=None, ctx_pos=None, sem_pos=None, t_emb=None, training=False,
                   dropout=0.1):
    """ChainedDiffuser's post-norm stack acting on seq1 (cross-attention to the read-only context
    [+ self-attention] [+ FFN]); layers.py:115-218 with pre_norm=False, cross_attention2 =
    self_attention2 = False.  Dropout (attention weights, residual branches, FFN) is active only
    when ``training`` (layers.py:146, 181; multihead_custom_attention.py:413)."""
    heads = stack.n_heads
    p_att = dropout if training else 0.0
    for layer in stack.layers:
        rot = stack.rotary_pe
        ada = stack.use_adaln and t_emb is not None
        # ---- cross attention seq1 -> seq2 (layers.py:124-147)
        q1 = x if sem_pos is None else x + sem_pos
        if ada:
            q1 = ada_ln(layer.adaln_12, q1, t_emb)
        att = mha(layer.cross_12, heads, q1, ctx, ctx, x_pos if rot else None, ctx_pos if rot else None,
                  dropout_p=p_att)
        x = residual_layer_norm(layer.norm_12, x, F.dropout(att, dropout, training))
        # ---- self attention on seq1 (layers.py:165-182)
        if stack.self_attention:
            qk = x if sem_pos is None else x + sem_pos
            vv = x
            if ada:
                qk = ada_ln(layer.adaln_1, qk, t_emb)
                vv = ada_ln(layer.adaln_1, vv, t_emb)
            att = mha(layer.sa1, heads, qk, qk, vv, x_pos if rot else None, x_pos if rot else None,
                      key_padding_mask=x_mask, dropout_p=p_att)
            x = residual_layer_norm(layer.norm_1, x, F.dropout(att, dropout, training))
        # ---- FFN-1 (layers.py:205-209); ffn_12 holds its own Dropout modules
        if stack.apply_ffn:
            y = ada_ln(layer.adaln_ff1, x, t_emb) if ada else x
            ffn = layer.ffn_12                      # Linear, ReLU, Dropout, Linear, Dropout (layers.py:76-82)
            hid = F.dropout(linear(y, ffn[0].weight, ffn[0].bias, relu=True), ffn[2].p, training)
            x = residual_layer_norm(layer.norm_122, y, F.dropout(linear(hid, ffn[3].weight, ffn[3].bias), ffn[4].p, training))
    return x
